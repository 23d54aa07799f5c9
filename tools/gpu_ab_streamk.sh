#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gemm_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x 2>&1 | tail -3 | cut -c1-300
for sk in 1 0; do
  FFVC_STREAM_K=$sk timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_sk$sk.md > gpurun_out/prof_step_sk$sk.log 2>&1; echo "== prof_step stream_k=$sk rc=$?"; grep -E "^# |atomic" gpurun_out/prof_step_sk$sk.log | cut -c1-160
done
