#!/bin/bash
# Final GPU session of the round: full parity suite, smoke, bench (with CPU baseline), ncu launch list of one step, step breakdown.
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/pytest_gpu.full 2>&1; rc=$?
cut -c1-400 gpurun_out/pytest_gpu.full | tail -60 > gpurun_out/pytest_gpu.log; rm -f gpurun_out/pytest_gpu.full
echo "== pytest -m gpu (rc=$rc): $(tail -1 gpurun_out/pytest_gpu.log)"
timeout -k 10 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout -k 10 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; cut -c1-3500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "== bench reference rc=$?"; cut -c1-600 gpurun_out/bench_ref.json
timeout -k 10 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --csv --log-file gpurun_out/launches_final.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
echo "== ncu rc=$? lines=$(wc -l < gpurun_out/launches_final.csv) $(tail -1 gpurun_out/one_step.log)"
timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_final.md > gpurun_out/prof_step.log 2>&1; echo "== prof_step rc=$?"
