#!/bin/bash
# GPU session: parity tests, A/B micro-benchmarks (LayerNorm pipe kernels with workspace reduction, GroupNorm rings),
# bench A/B over the switches, ncu launch list of ONE step with DRAM bytes.
mkdir -p gpurun_out
ON="ln_fwd_v2=2,ln_bwd_v2=1,pool_v2=1"
for f in test_ops_gpu test_models_gpu test_gemm_gpu; do
  timeout -k 10 420 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-600 gpurun_out/$f.full | tail -150 > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "== $f TIMED OUT: aborting"; exit 1; fi
done
timeout -k 10 300 python tools/ab_kernels.py > gpurun_out/ab_kernels.json 2> gpurun_out/ab_kernels.err; echo "== ab_kernels rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/ab_kernels.json"))
    for k, v in d.items():
        print("%-58s %9.1f us %7.0f" % (k, v["us"], list(v.values())[1]))
except Exception as e:
    print("ab_kernels parse failed", e)
PY
tail -3 gpurun_out/ab_kernels.err
run_bench() {  # name, FFVC_OPTS, FFVC_GN_FUSED_MIN
  FFVC_OPTS="$2" FFVC_GN_FUSED_MIN="$3" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "== bench $1 rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms', d['clocks'])" 2>&1 | tail -1)"; tail -2 gpurun_out/bench_$1.err
}
run_bench off "" ""
run_bench on "$ON" ""
run_bench on_gn8m "$ON,gn_ring=1" 8000000
run_bench on_gn2m "$ON,gn_ring=1" 2000000
FFVC_OPTS=$ON timeout -k 10 400 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
   --clock-control none --csv --log-file gpurun_out/launches_on.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
echo "== ncu rc=$? lines=$(wc -l < gpurun_out/launches_on.csv) $(tail -1 gpurun_out/one_step.log)"
FFVC_OPTS=$ON timeout -k 10 300 python tools/prof_step.py --out gpurun_out/step_breakdown_on.md > gpurun_out/prof_step.log 2>&1; echo "== prof_step rc=$?"
