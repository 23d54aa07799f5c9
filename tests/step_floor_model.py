#!/usr/bin/env python
"""Roofline floor of the whole train step from a LAUNCH TRACE of the real host path (no GPU needed).

The engines run on the CPU with every `ffvc_*` launch going to tests/abi_model.py (a torch statement of the C ABI) and a recorder
in front of it: for every launch the algorithmic FLOPs (2*M*N*K of GEMMs / convs, attention matmuls) and the algorithmic bytes
(every tensor operand once: what the launch must read and write at least) are taken from its arguments.  Two batch sizes give
the per-prompt and the batch-independent (weights, optimizer) parts, extrapolated to BASELINE config #2's 64 prompts.  Each launch's
floor is max(FLOPs / tensor peak, bytes / HBM bandwidth) with the peaks of MEASURED_PEAKS.json; the sum over the step is the time
below which no schedule of THESE launches can go — to be compared with the measured step (profiles/r01_step_breakdown_final.md).

    python tests/step_floor_model.py [--md profiles/r01_step_floor_model.md]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def trace_step(B, cutn=8):
    import abi_model
    import oracle.clip_vit as oclip
    import oracle.mixer as omix
    import oracle.vqgan as ovq
    from feed_forward_vqgan_clip_b200 import clip_vit, cutouts, mixer, ops, train_step, vqgan
    rec = []

    def nbytes(t):
        return t.numel() * t.element_size() if torch.is_tensor(t) else 0

    def gemm(a, b, out, M, N, K, **kw):
        nb, segs = kw.get("batch", 1), kw.get("k_segs", 1)
        a_n = (nb if kw.get("a_role", 0) == 1 else 1) * (segs if kw.get("a_role", 0) == 2 else 1)
        b_n = (nb if kw.get("b_role", 0) == 1 else 1) * (segs if kw.get("b_role", 0) == 2 else 1)
        a_bytes = 2.0 * M * K / 9 if kw.get("a_mode", 0) == 2 else 2.0 * M * K * a_n
        o_sz = 4.0 if out.dtype == torch.float32 else 2.0
        extra = sum(2.0 for k_ in ("pre_out", "aux", "res") if kw.get(k_) is not None)
        by = a_bytes + 2.0 * N * K * b_n + M * N * nb * (o_sz * (2 if kw.get("atomic") else 1) + extra)
        modes = "AK AM AC".split()[kw.get("a_mode", 0)] + "," + "BK BM".split()[kw.get("b_mode", 0)]
        epi = "act%d mul%d%s%s%s%s%s" % (kw.get("act", 0), kw.get("mul_mode", 0), " bias%d" % kw.get("bias_mode", 1) if kw.get("bias") is not None else "",
                                         " res" if kw.get("res") is not None else "", " pre" if kw.get("pre_out") is not None else "",
                                         " f32" if out.dtype == torch.float32 else "", " atomic" if kw.get("atomic") else "")
        rec.append(("gemm", 2.0 * M * N * K * nb * segs, by, ("gemm %dx%dx%d b%d seg%d " + modes + " " + epi, (M, N, K, nb, segs))))
        return abi_model.gemm_raw(a, b, out, M, N, K, **kw)

    def call(name, *a):
        fl = 0.0
        if name.startswith("conv3x3_halo"):
            n_, h_, w_, cin_, cout_ = a[3:8]
            fl = 2.0 * n_ * h_ * w_ * cout_ * 9 * cin_
        elif name in ("mha_small_fwd", "mha_small_bwd"):
            N_, T_, H_, dh_ = a[-5:-1] if name == "mha_small_fwd" else a[-5:-1]
            fl = (4.0 if name == "mha_small_fwd" else 10.0) * N_ * H_ * T_ * T_ * dh_
        ws = {"groupnorm_stats": (1,), "groupnorm_bwd": (6,), "layernorm_bwd_sums": (12,)}.get(name, ())
        by = sum(nbytes(t) for i, t in enumerate(a) if i not in ws)
        if name in ("adam_step", "adam_step_ema"):                      # p, m, v (and ema) are read AND written
            by += sum(nbytes(t) for t in (a[0], a[2], a[3])) + (nbytes(a[5]) if name == "adam_step_ema" else 0)
        if name.startswith("conv3x3_halo"):
            res = a[9] if name == "conv3x3_halo_gnbwd" else a[10]
            key = (name + " n%d %dx%d %d->%d" + (" res" if res is not None else ""), tuple(a[3:8]))
        else:
            ints = tuple(v for v in a if isinstance(v, int) and not isinstance(v, bool))[:5]
            key = (name + " " + ",".join(["%d"] * len(ints)), ints)
        rec.append((("tcgen05 conv " if fl and name.startswith("conv") else "") + name, fl, float(by), key))
        return abi_model.call(name, *a)

    ops.gemm_raw, ops.gemm, ops.call = abi_model.gemm_raw, gemm, call
    ops.require_cuda = lambda dev, what: None
    for mod in (mixer, vqgan, cutouts, train_step, clip_vit):
        mod.call = call
    net = mixer.Mixer(input_dim=512, image_size=16, channels=256, patch_size=1, dim=1024, depth=32)
    net.load_state_dict(omix.init_mixer_state_dict(512, 16, 256, 1024, 32, seed=0))
    vq = vqgan.VQModel()
    vq.load_state_dict(ovq.init_vqgan_state_dict(seed=1))
    clip = clip_vit.CLIP()
    clip.visual.load_state_dict(oclip.init_clip_state_dict(seed=2))
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 512, generator=g) * 0.45
    ts = train_step.TrainStep(net, vq.eval().requires_grad_(False), clip.eval().requires_grad_(False), cutn=cutn, lr=1e-3)
    prm = cutouts.sample_params(cutn * B, 224, g)
    ts.step(x, None, prm)                                                # first step: one-off work (bf16 shadow cast, codebook split)
    del rec[:]
    ts.step(x, None, prm)                                                # steady state: what a CUDA-graph replay launches
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default=None)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
        else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    bw, tf = peaks["hbm_gbs"] * 1e9, peaks["bf16_tflops_sustained"] * 1e12
    r1, r2 = trace_step(1), trace_step(2)
    assert [r[0] for r in r1] == [r[0] for r in r2], "the launch sequence must not depend on the batch size"
    B = args.batch
    fam = collections.OrderedDict()
    tot = [0.0, 0.0, 0.0, 0]
    shapes = collections.OrderedDict()
    for (name, f1, b1, (fmt, i1)), (_, f2, b2, (_, i2)) in zip(r1, r2):
        fl, by = f1 + (f2 - f1) * (B - 1), b1 + (b2 - b1) * (B - 1)     # linear in the batch: value(B) = v1 + (v2 - v1) (B - 1)
        floor = max(fl / tf, by / bw)
        try:
            skey = fmt % tuple(u + (v - u) * (B - 1) for u, v in zip(i1, i2))
        except TypeError:
            skey = name
        se = shapes.setdefault(skey, [0, 0.0, 0.0, 0.0])
        se[0] += 1
        se[1] += fl
        se[2] += by
        se[3] += floor
        key = "tcgen05 GEMM / conv" if (name == "gemm" or name.startswith("tcgen05")) else name
        e = fam.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
        e[0] += 1
        e[1] += fl
        e[2] += by
        e[3] += floor
        e[4] += fl / tf
        tot[0] += fl
        tot[1] += by
        tot[2] += floor
        tot[3] += 1
    # measured per-family times of the same launches: the committed CUDA-event breakdown of one eager step
    import re
    measured, mshape, mpath = {}, {}, os.path.join(ROOT, "profiles", "r01_step_breakdown_final.md")
    if os.path.exists(mpath):
        for ln in open(mpath):
            c = [x.strip() for x in ln.split("|")]
            if len(c) > 4 and c[2].isdigit():
                k2 = re.sub(r" sp\d+ ", " ", c[1])                      # the split-K factor is a launch knob, not a shape
                mshape[k2] = mshape.get(k2, 0.0) + float(c[3])
                nm = c[1].split()[0]
                nm = "tcgen05 GEMM / conv" if (nm == "gemm" or nm.startswith("conv3x3_halo")) else nm
                measured[nm] = measured.get(nm, 0.0) + float(c[3])
    lines = ["# Roofline floor of the config #2 train step (B = %d) from a launch trace of the host path" % B, "",
             "`python tests/step_floor_model.py` — engines on the CPU, every launch recorded in front of `tests/abi_model.py`; algorithmic",
             "FLOPs and bytes per launch from its arguments, traced at B = 1 and 2 and extended linearly to B = %d.  Peaks: HBM %.0f GB/s," % (B, peaks["hbm_gbs"]),
             "bf16 %.0f TFLOP/s sustained (`MEASURED_PEAKS.json`).  floor = sum over launches of max(FLOPs / peak, bytes / bandwidth)." % peaks["bf16_tflops_sustained"],
             "measured = CUDA-event time of the same launches in one eager step on a B200 (`profiles/r01_step_breakdown_final.md`).", "",
             "| launch family | launches | TFLOP | GB | floor ms | of which tensor ms | measured ms | floor / measured |",
             "|---|---:|---:|---:|---:|---:|---:|---:|"]
    msum = 0.0
    for k, (n, fl, by, fo, tm) in sorted(fam.items(), key=lambda kv: -kv[1][3]):
        m = measured.get(k)
        msum += m or 0.0
        lines.append("| %s | %d | %.2f | %.2f | %.3f | %.3f | %s | %s |" % (k, n, fl / 1e12, by / 1e9, fo * 1e3, tm * 1e3,
                                                                   "%.3f" % m if m else "", "%.2f" % (fo * 1e3 / m) if m else ""))
    lines.append("| **total** | %d | %.2f | %.2f | **%.2f** | | %.2f | %.2f |" % (tot[3], tot[0] / 1e12, tot[1] / 1e9, tot[2] * 1e3, msum,
                                                                         tot[2] * 1e3 / msum if msum else 0))
    lines += ["", "Step floor %.1f ms = %.0f prompts/s at B = %d; pure tensor time at the sustained peak %.1f ms; pure HBM time of all launches %.1f ms." %
              (tot[2] * 1e3, B / tot[2], B, tot[0] / tf * 1e3, tot[1] / bw * 1e3)]
    lines += ["", "## Per launch shape (the 40 largest measured times; key as in `profiles/r01_step_breakdown_final.md` without the split factor)", "",
              "| launch | n | GFLOP each | MB each | floor us each | measured us each | floor / measured | ms above the floor (all n) |",
              "|---|---:|---:|---:|---:|---:|---:|---:|"]
    rows = []
    for k, (n, fl, by, fo) in shapes.items():
        m = mshape.get(k)
        if m:
            rows.append((m - fo * 1e3, k, n, fl, by, fo, m))
    for gap, k, n, fl, by, fo, m in sorted(rows, key=lambda r: -r[6])[:40]:
        lines.append("| %s | %d | %.1f | %.1f | %.1f | %.1f | %.2f | %.2f |" % (k, n, fl / n / 1e9, by / n / 1e6, fo / n * 1e6, m / n * 1e3,
                                                                        fo * 1e3 / m, gap))
    unmatched = [k for k in shapes if k not in mshape]
    lines += ["", "%d of %d launch shapes matched a measured row (%.1f of %.1f ms measured)." %
              (len(rows), len(shapes), sum(r[6] for r in rows), sum(mshape.values()))]
    out = "\n".join(lines) + "\n"
    print(out)
    if args.md:
        open(args.md, "w").write(out)


if __name__ == "__main__":
    main()
