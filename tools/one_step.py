#!/usr/bin/env python
"""ncu target: ONE eager train step of the bench workload (bench.py --config C, default #2) between cudaProfilerStart/Stop, so that
`ncu --profile-from-start off` lists exactly the kernels of one step:

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none --csv --log-file gpurun_out/launches.csv python tools/one_step.py

The workload is built by bench.build_b200 (same nets, batch, seeds as `python bench.py`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--config", type=int, default=2)
    args = ap.parse_args()
    conf = bench.CONFIGS[args.config]
    args.batch = args.batch or conf["batch"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ts = bench.build_b200(dev, conf, 1, None)
    x = bench.synthetic_embeddings(args.batch, 1000).to(dev)
    ts.step(x)                       # allocator warm-up, lazy init, TMA descriptor caches
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    loss = ts.step(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("one step done, loss %.5f" % loss.item())


if __name__ == "__main__":
    main()
