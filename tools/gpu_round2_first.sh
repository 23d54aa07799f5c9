#!/bin/bash
# First GPU session of round 2 (one B200):  bash tools/gpu_round2_first.sh
# 1. the whole GPU suite, including the tests written after round 1's last GPU session (tests/test_zz_*_gpu.py)
# 2. smoke + bench (N = 1) + reference arm
# For the data-parallel A/Bs use:  gpurun --gpus 2 -- 'bash tools/gpu_multi_ab2.sh 2'
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_tests.log; echo "== tests rc=$? $(tail -1 gpurun_out/r2_tests.log)"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "== smoke rc=$? $(tail -1 gpurun_out/r2_smoke.log)"
timeout -k 10 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "== bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','last_loss','clocks')}, d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline',{}).get('value'))"
timeout -k 10 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "== reference arm rc=$?"
