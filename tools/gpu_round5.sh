#!/bin/bash
# GPU session: GEMM per-shape tuning, then bench with and without the tuned table.
mkdir -p gpurun_out
timeout -k 10 330 python tools/tune_gemm.py --out gpurun_out/gemm_tuned.json > gpurun_out/tune.log 2>&1; echo "== tune rc=$?"; tail -32 gpurun_out/tune.log | cut -c1-200
if [ -s gpurun_out/gemm_tuned.json ]; then cp gpurun_out/gemm_tuned.json feed_forward_vqgan_clip_b200/gemm_tuned.json; fi
run_bench() {  # name, FFVC_GEMM_TUNED
  FFVC_GEMM_TUNED="$2" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "== bench $1 rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms', d['clocks'], 'gemm TF', round(d['roofline']['achieved'],1), 'alg bytes/launch', d['roofline'].get('algorithmic_bytes_per_launch'))" 2>&1 | tail -1)"; tail -2 gpurun_out/bench_$1.err
}
run_bench untuned 0
run_bench tuned 1
timeout -k 10 200 python -m pytest tests/test_models_gpu.py -q -m gpu -p no:cacheprovider --tb=short -x > gpurun_out/models_tuned.log 2>&1; echo "== models tests with table rc=$? $(tail -1 gpurun_out/models_tuned.log)"
