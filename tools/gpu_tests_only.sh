#!/bin/bash
mkdir -p gpurun_out
for f in test_gemm_gpu test_ops_gpu test_models_gpu; do
  timeout -k 10 400 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-400 gpurun_out/$f.full > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
  if [ $rc -ne 0 ]; then grep -E "^E  |^tests/|Error" gpurun_out/$f.log | head -12; fi
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "== $f TIMED OUT: aborting"; exit 1; fi
done
timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_c10.md > gpurun_out/prof_step_c10.log 2>&1; echo "== prof_step rc=$?"; grep -E "^# |layernorm|conv3x3_halo" gpurun_out/prof_step_c10.log | cut -c1-160
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
