#!/bin/bash
# multi-GPU bench exactly as the driver launches it:  bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/nvsmi_multi.txt 2>&1
NCCL_DEBUG=WARN timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "== bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
   bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "== reference arm N=$N rc=$?"; tail -c 600 gpurun_out/bench_ref_n$N.json
