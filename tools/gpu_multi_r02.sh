#!/bin/bash
# N-GPU A/B of the data-parallel knobs (run under gpurun --gpus N): bench lines into gpurun_out/r02_n${N}_*.json
N=${1:-2}
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 15 --warmup 4 "$@" > gpurun_out/r02_n${N}_$tag.json 2> gpurun_out/r02_n${N}_$tag.err
  echo "$tag rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r02_n${N}_$tag.json'));print(round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
  tail -c 300 gpurun_out/r02_n${N}_$tag.err | grep -i "teardown\|error" 
}
timeout 200 python bench.py --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/r02_n${N}_single.json 2>/dev/null
echo "single $(python -c "import json;d=json.load(open('gpurun_out/r02_n${N}_single.json'));print(round(d['value'],1), round(d['ms_per_step'],2))")"
run default
run tail --tail-overlap
run comm8 --comm-sms 8
run comm8_tail --comm-sms 8 --tail-overlap
run comm16_tail --comm-sms 16 --tail-overlap
run ctas8_tail --nccl-max-ctas 8 --tail-overlap
run b4_comm8_tail --comm-sms 8 --tail-overlap --bucket-layers 4
