#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 42 -c 4 -o gpurun_out/prof_mixer_fwd_r3 \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_a.log 2>&1
echo "== ncu mixer fwd rc=$?"
ls -la gpurun_out/*.ncu-rep
