#!/bin/bash
# GPU session 2: halo-conv descriptor variants, e2e gradient diagnostic, per-launch step breakdown, full test suite.
mkdir -p gpurun_out
for v in 0 1; do
  FFVC_HALO_BASEOFF=$v timeout -k 10 200 python -m pytest tests/test_gemm_gpu.py -q -m gpu -k halo -p no:cacheprovider --tb=line 2>&1 | cut -c1-300 > gpurun_out/halo_v$v.log
  echo "== halo baseoff=$v: $(tail -1 gpurun_out/halo_v$v.log)"
done
timeout -k 10 300 python tools/diag_e2e.py > gpurun_out/diag_e2e.log 2>&1; echo "== diag rc=$?"; grep -v Warning gpurun_out/diag_e2e.log | cut -c1-250 | tail -20
for f in test_gemm_gpu test_ops_gpu test_models_gpu; do
  timeout -k 10 300 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-400 gpurun_out/$f.full > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
done
timeout -k 10 300 python tools/prof_step.py > gpurun_out/prof_step.log 2>&1; echo "== prof_step rc=$?"; head -30 gpurun_out/prof_step.log | cut -c1-200
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 8 -c 2 -o gpurun_out/prof_halo \
   python bench.py --no-graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/prof_halo.log 2>&1
echo "== ncu full halo rc=$?"
