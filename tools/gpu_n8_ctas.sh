#!/bin/bash
# N = 8: NCCL CTA cap (NVLS all-reduce needs few SMs) against the default, config #2
N=8
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
    bench.py --gpus $N --steps 15 --warmup 4 "$@" > gpurun_out/r02_n8b_$tag.json 2> gpurun_out/r02_n8b_$tag.err
  echo "$tag rc=$? $(python - <<P
import json
try:
    l=[x for x in open('gpurun_out/r02_n8b_$tag.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('FAILED', e)
P
)"; }
timeout 200 python bench.py --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/r02_n8b_single.json 2>/dev/null
echo "single $(python -c "
import json
l=[x for x in open('gpurun_out/r02_n8b_single.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]);print(round(d['value'],1), round(d['ms_per_step'],2))")"
run default
run ctas8 --nccl-max-ctas 8
run ctas16 --nccl-max-ctas 16
