#!/bin/bash
# same-box A/B of bench.py under environment switches: tools/gpu_ab_bench.sh "TAG1:ENV1=..,ENV2=.." "TAG2:" ...   (2 rounds each, interleaved)
mkdir -p gpurun_out
for round in a b; do
  for spec in "$@"; do
    tag=${spec%%:*}; envs=${spec#*:}
    env $(echo $envs | tr ',' ' ') timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline $BENCH_ARGS > gpurun_out/ab_${tag}_$round.json 2> gpurun_out/ab_${tag}_$round.err
    python - <<P
import json
try:
    l=[x for x in open('gpurun_out/ab_${tag}_$round.json').read().splitlines() if x.startswith('{')]
    d=json.loads(l[-1]); print('$tag $round', round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms  e2e', round(d['e2e']['value'],1), 'clk', d['clocks']['sm_mhz'], 'loss', round(d['last_loss'],4))
except Exception as e:
    print('$tag $round FAILED', e); print(open('gpurun_out/ab_${tag}_$round.err').read()[-1500:])
P
  done
done
