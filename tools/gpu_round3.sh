#!/bin/bash
# GPU session (round 1, re-entry): parity tests incl. the switchable kernel variants, A/B micro-benchmarks,
# bench A/B (switches off / on), ncu launch list of ONE step with DRAM bytes.  Everything logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
ON="ln_fwd_v2=1,ln_bwd_v2=1,pool_v2=1"
for f in test_ops_gpu test_models_gpu test_gemm_gpu; do
  timeout -k 10 420 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-600 gpurun_out/$f.full | tail -150 > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "== $f TIMED OUT: aborting"; exit 1; fi
done
timeout -k 10 240 python tools/ab_kernels.py > gpurun_out/ab_kernels.json 2> gpurun_out/ab_kernels.err; echo "== ab_kernels rc=$?"; cat gpurun_out/ab_kernels.json | tr -d '\n' | cut -c1-3000; echo; tail -3 gpurun_out/ab_kernels.err
FFVC_OPTS=$ON timeout -k 10 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_on.json 2> gpurun_out/bench_on.err; echo "== bench ON rc=$?"; cut -c1-700 gpurun_out/bench_on.json; tail -3 gpurun_out/bench_on.err
FFVC_OPTS=$ON timeout -k 10 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --csv --log-file gpurun_out/launches_on.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
echo "== ncu rc=$? lines=$(wc -l < gpurun_out/launches_on.csv) $(tail -1 gpurun_out/one_step.log)"
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_off.json 2> gpurun_out/bench_off.err; echo "== bench OFF rc=$?"; cut -c1-400 gpurun_out/bench_off.json; tail -3 gpurun_out/bench_off.err
FFVC_OPTS=$ON timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_on.md > gpurun_out/prof_step.log 2>&1; echo "== prof_step rc=$?"
