#!/bin/bash
# N-GPU bench, bucketed/overlapped gradient all-reduce vs one all-reduce after backward:  bash tools/gpu_multi_ab.sh N
N=${1:-2}
mkdir -p gpurun_out
for bl in 8 0; do
  NCCL_DEBUG=WARN timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$bl \
     bench.py --gpus $N --steps 8 --warmup 3 --bucket-layers $bl > gpurun_out/bench_n${N}_b$bl.json 2> gpurun_out/bench_n${N}_b$bl.err
  echo "== bench N=$N bucket_layers=$bl rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n${N}_b$bl.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','last_loss','clocks')}, d['e2e']['value'])"; tail -3 gpurun_out/bench_n${N}_b$bl.err | cut -c1-300
done
