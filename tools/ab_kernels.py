#!/usr/bin/env python
"""A/B micro-benchmark of the switchable kernel variants (ffvc_set_option) at the bench workload's shapes, on the GPU box:

    python tools/ab_kernels.py > gpurun_out/ab_kernels.json

Each launch works on a different one of NSETS buffer sets (together larger than the 126 MB L2), timed with CUDA events
over REPS launches after warm-up; reported: microseconds per launch and the implied HBM GB/s over the algorithmic bytes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from feed_forward_vqgan_clip_b200 import _lib  # noqa: E402
from feed_forward_vqgan_clip_b200.ops import call  # noqa: E402

DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32
NSETS, REPS = 6, 30


def timeit(fn):
    for i in range(NSETS):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(REPS):
        fn(i % NSETS)
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / REPS          # us per launch


def setopt(**kw):
    lib = _lib.load()
    for k, v in kw.items():
        assert lib.ffvc_set_option(k.encode(), int(v)) >= 0


def main():
    torch.cuda.set_device(0)
    out = {}
    # ---------------- LayerNorm, mixer shape (config #2: 64 prompts x 256 tokens, D = 1024)
    rows, D, T = 16384, 1024, 256
    sets = []
    for i in range(NSETS):
        g = torch.Generator(device=DEV).manual_seed(i)
        sets.append(dict(x=torch.randn(rows, D, device=DEV, generator=g).to(BF), dy=torch.randn(rows, D, device=DEV, generator=g).to(BF),
                         add=torch.randn(rows, D, device=DEV, generator=g).to(BF), y=torch.empty(rows, D, device=DEV, dtype=BF),
                         mean=torch.zeros(rows, device=DEV), rstd=torch.ones(rows, device=DEV)))
    gamma, beta = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
    dg, db, cs = (torch.zeros(D, device=DEV) for _ in range(3))
    rs = torch.zeros(T, device=DEV)
    tensor_bytes = rows * D * 2
    ws = torch.empty(int(_lib.load().ffvc_layernorm_bwd_ws_bytes(D, T)) // 4, device=DEV)

    def ln_fwd(i):
        s = sets[i]
        call("layernorm_fwd", s["x"], gamma, beta, s["y"], s["mean"], s["rstd"], rows, D, 1e-5)

    def ln_bwd_col(i):       # LN1 backward of a mixer layer: wgrad + residual add + column sums of dx
        s = sets[i]
        call("layernorm_bwd_sums", s["dy"], s["x"], gamma, s["mean"], s["rstd"], s["add"], s["y"], dg, db, cs, None, 0, ws, rows, D)

    def ln_bwd_row(i):       # LN2 backward: wgrad + residual add + per-token row sums of dx
        s = sets[i]
        call("layernorm_bwd_sums", s["dy"], s["x"], gamma, s["mean"], s["rstd"], s["add"], s["y"], dg, db, None, rs, T, ws, rows, D)

    def ln_bwd_plain(i):
        s = sets[i]
        call("layernorm_bwd", s["dy"], s["x"], gamma, s["mean"], s["rstd"], s["add"], s["y"], dg, db, rows, D)

    for i in range(NSETS):
        ln_fwd(i)
    for v2 in (0, 1, 2):
        setopt(ln_fwd_v2=v2, ln_bwd_v2=v2)
        t = timeit(ln_fwd)
        out["layernorm_fwd 16384x1024 v2=%d" % v2] = {"us": t, "GBps": 2 * tensor_bytes / t / 1e3}
        t = timeit(ln_bwd_col)
        out["layernorm_bwd+colsum 16384x1024 v2=%d" % v2] = {"us": t, "GBps": 4 * tensor_bytes / t / 1e3}
        t = timeit(ln_bwd_row)
        out["layernorm_bwd+rowsum 16384x1024 v2=%d" % v2] = {"us": t, "GBps": 4 * tensor_bytes / t / 1e3}
    setopt(ln_fwd_v2=0, ln_bwd_v2=0)
    t = timeit(ln_bwd_plain)
    out["layernorm_bwd (no sums) 16384x1024 v1"] = {"us": t, "GBps": 4 * tensor_bytes / t / 1e3}
    # CLIP shape (no wgrad): 25600 x 768
    rows2, D2 = 25600, 768
    sets2 = []
    for i in range(NSETS):
        g = torch.Generator(device=DEV).manual_seed(100 + i)
        sets2.append(dict(x=torch.randn(rows2, D2, device=DEV, generator=g).to(BF), dy=torch.randn(rows2, D2, device=DEV, generator=g).to(BF),
                          add=torch.randn(rows2, D2, device=DEV, generator=g).to(BF), y=torch.empty(rows2, D2, device=DEV, dtype=BF),
                          mean=torch.zeros(rows2, device=DEV), rstd=torch.ones(rows2, device=DEV)))
    gamma2, beta2 = torch.ones(D2, device=DEV), torch.zeros(D2, device=DEV)

    def ln2_fwd(i):
        s = sets2[i]
        call("layernorm_fwd", s["x"], gamma2, beta2, s["y"], s["mean"], s["rstd"], rows2, D2, 1e-5)

    def ln2_bwd(i):
        s = sets2[i]
        call("layernorm_bwd_sums", s["dy"], s["x"], gamma2, s["mean"], s["rstd"], s["add"], s["y"], None, None, None, None, 0, ws, rows2, D2)

    for i in range(NSETS):
        ln2_fwd(i)
    for v2 in (0, 1, 2):
        setopt(ln_fwd_v2=v2, ln_bwd_v2=v2)
        t = timeit(ln2_fwd)
        out["layernorm_fwd 25600x768 v2=%d" % v2] = {"us": t, "GBps": 2 * rows2 * D2 * 2 / t / 1e3}
        t = timeit(ln2_bwd)
        out["layernorm_bwd (frozen) 25600x768 v2=%d" % v2] = {"us": t, "GBps": 4 * rows2 * D2 * 2 / t / 1e3}
    setopt(ln_fwd_v2=0, ln_bwd_v2=0)
    del sets, sets2
    # ---------------- GroupNorm(32)+swish at the decoder's shapes: two-pass kernels vs the single-kernel L2-resident forms
    for (N, HW, C) in ((64, 65536, 128), (64, 16384, 128), (64, 16384, 256), (64, 4096, 256)):
        g = torch.Generator(device=DEV).manual_seed(7)
        x = torch.randn(N, HW, C, device=DEV, generator=g).to(BF)
        dy = torch.randn(N, HW, C, device=DEV, generator=g).to(BF)
        add = torch.randn(N, HW, C, device=DEV, generator=g).to(BF)
        y = torch.empty_like(x)
        gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        gws = torch.empty(max((int(_lib.load().ffvc_groupnorm_ws_bytes(N, 32)) + 7) // 8, int(_lib.load().ffvc_groupnorm_ws_doubles(N, HW, 32))),
                          device=DEV, dtype=torch.float64)
        mean, rstd = torch.empty(N * 32, device=DEV), torch.empty(N * 32, device=DEV)
        tb = N * HW * C * 2

        def two_fwd(i):
            call("groupnorm_stats", x, gws, mean, rstd, N, HW, C, 32, 1e-6)
            call("groupnorm_apply", x, mean, rstd, gam, bet, y, N, HW, C, 32, 1)

        def two_bwd(i):
            call("groupnorm_bwd", dy, x, mean, rstd, gam, bet, gws, add, y, N, HW, C, 32, 1)

        def fused_fwd(i):
            call("groupnorm_fused_fwd", x, gam, bet, y, mean, rstd, gws, N, HW, C, 32, 1, 1e-6)

        def fused_bwd(i):
            call("groupnorm_fused_bwd", dy, x, mean, rstd, gam, bet, gws, add, y, N, HW, C, 32, 1)

        tag = "%dx%dx%d" % (N, HW, C)
        t = timeit(two_fwd)
        out["groupnorm fwd two-pass %s" % tag] = {"us": t, "GBps": 3 * tb / t / 1e3}
        t = timeit(two_bwd)
        out["groupnorm bwd two-pass %s" % tag] = {"us": t, "GBps": 6 * tb / t / 1e3}
        for ring in (0, 1):
            setopt(gn_ring=ring)
            t = timeit(fused_fwd)
            out["groupnorm fwd fused ring=%d %s" % (ring, tag)] = {"us": t, "GBps_algorithmic_2pass": 3 * tb / t / 1e3}
            t = timeit(fused_bwd)
            out["groupnorm bwd fused ring=%d %s" % (ring, tag)] = {"us": t, "GBps_algorithmic_2pass": 6 * tb / t / 1e3}
        setopt(gn_ring=0)
        del x, dy, add, y
    # ---------------- cutout pool backward (64 images 256x256 -> 224x224)
    B, H, P = 64, 256, 224
    ps = [dict(x=torch.rand(B, H, H, 3, device=DEV), dy=torch.randn(B, P, P, 3, device=DEV), dx=torch.empty(B, H, H, 3, device=DEV))
          for _ in range(NSETS)]

    def pool(i):
        s = ps[i]
        call("cutout_pool_bwd", s["x"], s["dy"], s["dx"], B, H, H, P, 0)

    for v2 in (0, 1):
        setopt(pool_v2=v2)
        t = timeit(pool)
        out["cutout_pool_bwd 64x256x256 v2=%d" % v2] = {"us": t, "GBps": (2 * B * H * H * 3 * 4 + B * P * P * 3 * 4) / t / 1e3}
    setopt(pool_v2=0)
    del ps
    # ---------------- fused Adam (16-byte form) on the 319.7 M-parameter arena, with and without the EMA arena
    n = 319696128
    p, g_, m, v = (torch.zeros(n, device=DEV) for _ in range(4))
    g_.normal_()
    sh = torch.empty(n, device=DEV, dtype=BF)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1, 1, 1, 0, 0, 0, 0, 1, 1e-3, 0, 0, 0], device=DEV, dtype=F32)

    def adam(i):
        call("adam_tick", hyper)
        call("adam_step", p, g_, m, v, sh, n, hyper)

    t = timeit(adam)
    out["adam_tick+adam_step 319.7M (26 B/param)"] = {"us": t, "GBps": 26.0 * n / t / 1e3}
    ema = torch.zeros(n, device=DEV)
    hyper[15] = 0.995

    def adam_ema(i):
        call("adam_tick", hyper)
        call("adam_step_ema", p, g_, m, v, sh, ema, n, hyper)

    t = timeit(adam_ema)
    out["adam_tick+adam_step_ema 319.7M (34 B/param)"] = {"us": t, "GBps": 34.0 * n / t / 1e3}

    def sumsq(i):
        call("sumsq", g_, hyper[10:11], n)

    t = timeit(sumsq)
    out["sumsq (clip_grad_norm) 319.7M (4 B/param)"] = {"us": t, "GBps": 4.0 * n / t / 1e3}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
