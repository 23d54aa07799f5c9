#!/bin/bash
# N-GPU A/B of the data-parallel knobs (after the round-1 fix that moved `proj` into the last bucket):  bash tools/gpu_multi_ab2.sh N
#   bucket size 8 / 4 / 2, Adam overlapped with the last bucket, NCCL CTA cap
N=${1:-2}
mkdir -p gpurun_out
i=0
for cfg in "--bucket-layers 8" "--bucket-layers 4" "--bucket-layers 2" "--bucket-layers 4 --tail-overlap" "--bucket-layers 4 --nccl-max-ctas 8" "--bucket-layers 0"; do
  i=$((i+1)); tag=$(echo $cfg | tr -d ' -' )
  NCCL_DEBUG=WARN timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$i \
     bench.py --gpus $N --steps 8 --warmup 3 $cfg > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
  echo "== N=$N $cfg rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_n${N}_$tag.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','last_loss')}, d['e2e']['value'])"; tail -2 gpurun_out/bench_n${N}_$tag.err | cut -c1-200
done
