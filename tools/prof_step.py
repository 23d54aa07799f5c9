#!/usr/bin/env python
"""Per-launch breakdown of one eager train step (bench.py --config C) with CUDA events around EVERY ffvc_* launch, aggregated by
(kernel, shape).  Runs on the GPU box:  python tools/prof_step.py [--config 2] [--batch 0] [--out gpurun_out/step_breakdown.md]
The numbers are warm-cache, in-step timings (unlike the serialised cold-cache ncu launch list)."""
import argparse
import collections
import os
import sys

os.environ.setdefault("FFVC_SIDE_STREAMS", "0")     # every launch on the timed stream

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from feed_forward_vqgan_clip_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/step_breakdown.md")
    ap.add_argument("--gn-pipeline", type=int, default=1)
    args = ap.parse_args()
    from feed_forward_vqgan_clip_b200 import _lib
    _lib.load().ffvc_groupnorm_set_pipeline(args.gn_pipeline)
    if os.environ.get("FFVC_STREAM_K") is not None:
        _lib.load().ffvc_gemm_set_stream_k(int(os.environ["FFVC_STREAM_K"]))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    conf = bench.CONFIGS[args.config]
    args.batch = args.batch or conf["batch"]
    ts = bench.build_b200(dev, conf, 1, None)
    x = bench.synthetic_embeddings(args.batch, 1000).to(dev)
    ts.step(x)
    ts.step(x)
    torch.cuda.synchronize()

    recs = []
    real_gemm, real_call = ops.gemm, ops.call

    def timed_gemm(a, b, out, M, N, K, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = real_gemm(a, b, out, M, N, K, **kw)
        e.record()
        modes = "AK AM AC".split()[kw.get("a_mode", 0)] + "," + "BK BM".split()[kw.get("b_mode", 0)]
        epi = "act%d mul%d%s%s%s%s%s" % (kw.get("act", 0), kw.get("mul_mode", 0), " bias%d" % kw.get("bias_mode", 1) if kw.get("bias") is not None else "",
                                         " res" if kw.get("res") is not None else "", " pre" if kw.get("pre_out") is not None else "",
                                         " f32" if out.dtype == torch.float32 else "", " atomic" if kw.get("atomic") else "")
        key = "gemm %dx%dx%d b%d seg%d sp%d %s %s" % (M, N, K, kw.get("batch", 1), kw.get("k_segs", 1), kw.get("splits", 1), modes, epi)
        recs.append((key, 2.0 * M * N * K * kw.get("batch", 1) * kw.get("k_segs", 1), s, e))
        return r

    def timed_call(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_call(name, *a)
        e.record()
        key, fl = name, 0.0
        if name.startswith("conv3x3_halo"):
            n, h, w, cin, cout = a[3:8]
            res = a[9] if name == "conv3x3_halo_gnbwd" else a[10]
            key = "%s n%d %dx%d %d->%d%s" % (name, n, h, w, cin, cout, " res" if res is not None else "")
            fl = 2.0 * n * h * w * cout * 9 * cin
        else:
            ints = [str(v) for v in a if isinstance(v, int) and not isinstance(v, bool)][:5]
            key = name + " " + ",".join(ints)
        recs.append((key, fl, s, e))

    ops.gemm = timed_gemm
    for m in list(sys.modules.values()):
        if m is not None and getattr(m, "__name__", "").startswith("feed_forward_vqgan_clip_b200") and getattr(m, "call", None) is real_call:
            m.call = timed_call
    s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_all.record()
    ts.step(x)
    e_all.record()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for key, fl, s, e in recs:
        t = s.elapsed_time(e)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += fl
    tot = sum(a[1] for a in agg.values())
    lines = ["# eager step breakdown (CUDA events per launch), config #%d, batch %d: step %.2f ms, sum of launches %.2f ms, %d launches"
             % (args.config, args.batch, s_all.elapsed_time(e_all), tot, len(recs)), "", "| launch | n | ms total | us each | TFLOP/s |", "|---|---:|---:|---:|---:|"]
    for key, (n, t, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.3f | %.1f | %s |" % (key, n, t, 1e3 * t / n, ("%.0f" % (fl / t / 1e9)) if fl else ""))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
