#!/bin/bash
# 8-GPU lines for every BASELINE config + the data-parallel knobs on config #2 (run under gpurun --gpus 8)
N=${1:-8}
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --steps 15 --warmup 4 "$@" > gpurun_out/r02_n${N}_$tag.json 2> gpurun_out/r02_n${N}_$tag.err
  echo "$tag rc=$? $(python - <<P
import json
try:
    l=[x for x in open('gpurun_out/r02_n${N}_$tag.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print(round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('FAILED', e)
P
)"
  grep -i "teardown\|Traceback\|Error" gpurun_out/r02_n${N}_$tag.err | head -3
}
for c in 2 3 4 5; do
  timeout 300 python bench.py --config $c --steps 15 --warmup 4 --no-cpu-baseline > gpurun_out/r02_n1_c$c.json 2>/dev/null
  echo "single config $c $(python -c "
import json
l=[x for x in open('gpurun_out/r02_n1_c$c.json').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]);print(round(d['value'],1), round(d['ms_per_step'],2))")"
done
run c2_default --config 2
run c2_noadam --config 2 --no-adam-overlap
run c2_noadam_tail --config 2 --no-adam-overlap --tail-overlap
run c3_default --config 3
run c4_default --config 4
run c5_default --config 5
