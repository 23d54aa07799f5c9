#!/bin/bash
# GPU session 4: GN / GELU tests, then ncu --set full of one mixer layer's forward GEMMs and one backward GELU' GEMM pair.
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider --tb=line -k "groupnorm or gelu or colsum" 2>&1 | cut -c1-300 | tail -5
timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_c4.md > gpurun_out/prof_step_c4.log 2>&1; echo "== prof_step rc=$?"; grep -E "groupnorm|^# " gpurun_out/prof_step_c4.log | head -12 | cut -c1-160
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 42 -c 4 -o gpurun_out/prof_mixer_fwd_r2 \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_a.log 2>&1
echo "== ncu mixer fwd rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 454 -c 2 -o gpurun_out/prof_mixer_bwd_r2 \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_b.log 2>&1
echo "== ncu mixer bwd rc=$?"
ls -la gpurun_out/*.ncu-rep
