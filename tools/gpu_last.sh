#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 60 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider --tb=line -k "layernorm or mixer" > gpurun_out/ln.log 2>&1; echo "== LN tests rc=$? $(tail -1 gpurun_out/ln.log)"
timeout -k 5 70 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "== bench rc=$? $(python -c "import json; d=json.load(open('gpurun_out/bench_last.json')); print(round(d['value'],1), round(d['ms_per_step'],2), d['last_loss'])" 2>&1 | tail -1)"
