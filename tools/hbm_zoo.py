#!/usr/bin/env python
"""One launch of every HBM-bound kernel of the train step at config #2's shapes (64 prompts) — the ncu --set full target for the
non-GEMM kernels:  ncu --set full --clock-control none -k regex:. -o gpurun_out/r02_ncu_zoo python tools/hbm_zoo.py
(a warm-up round runs first; ncu's --launch-skip drops it:  -s <number printed by --count>)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from feed_forward_vqgan_clip_b200 import _lib, ops  # noqa: E402
from feed_forward_vqgan_clip_b200.cutouts import CLIP_MEAN, CLIP_STD, CutoutEngine, sample_params  # noqa: E402
from feed_forward_vqgan_clip_b200.ops import call  # noqa: E402

DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32


def rnd(*s, dt=BF, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*s, device=DEV, generator=g).to(dt)


def main():
    torch.cuda.set_device(0)
    lib = _lib.load()
    B, T, D = 64, 256, 1024
    R = B * T
    x, dy, add = rnd(R, D, seed=1), rnd(R, D, seed=2), rnd(R, D, seed=3)
    y = torch.empty_like(x)
    gamma, beta = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
    mean, rstd = torch.empty(R, device=DEV), torch.empty(R, device=DEV)
    dg, db, cs, rs = (torch.zeros(n, device=DEV) for n in (D, D, D, T))
    lnws = torch.empty(int(lib.ffvc_layernorm_bwd_ws_bytes(D, T)) // 4, device=DEV)
    N, HW, Cc = 64, 65536, 128
    gx, gdy, gadd = rnd(N * HW, Cc, seed=4), rnd(N * HW, Cc, seed=5), rnd(N * HW, Cc, seed=6)
    gy = torch.empty_like(gx)
    gm, gr = torch.empty(N * 32, device=DEV), torch.empty(N * 32, device=DEV)
    gga, gbe = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    gws = torch.empty(int(lib.ffvc_groupnorm_ws_doubles(N, HW, 32)), device=DEV, dtype=torch.float64)
    n_par = 319696128
    p, g, m, v = (torch.zeros(n_par, device=DEV) for _ in range(4))
    shadow = torch.empty(n_par, device=DEV, dtype=BF)
    hyper = torch.tensor([1e-3, .9, .999, 1e-8, 0.1, 0.03, 1, 0, 1, 0, 0, 1, 1e-3, 0, 0, 0], device=DEV)
    big = rnd(R, 4 * D, seed=7)
    bsum, rsum = torch.zeros(4 * D, device=DEV), torch.zeros(4 * T, device=DEV)
    img = torch.rand(B, 256, 256, 3, device=DEV)
    cut = CutoutEngine(224, 8, 32, torch.device(DEV))
    prm = sample_params(8 * B, 224, torch.Generator().manual_seed(1), with_noise=False)
    prm = {k: (t.to(DEV) if torch.is_tensor(t) else t) for k, t in prm.items()}
    prm["erase"] = torch.tensor([int(t) for t in prm["erase"]], dtype=torch.int32, device=DEV)
    prm["facs"] = torch.rand(8 * B, device=DEV) * 0.1
    prm["noise_raw"] = torch.randn(8 * B, 3, 224, 224, device=DEV)
    qkv, dout = rnd(512, 50, 2304, seed=8), rnd(512, 50, 768, seed=9)
    att, dqkv = torch.empty(512, 50, 768, device=DEV, dtype=BF), torch.empty(512, 50, 2304, device=DEV, dtype=BF)
    taps = rnd(N * HW, 32, dt=F32, seed=10)
    raw, post = torch.empty(N * HW, 3, device=DEV), torch.empty(N * HW, 3, device=DEV)
    up_in = rnd(N, 128, 128, 128, seed=11)
    up_out = torch.empty(N, 256, 256, 128, device=DEV, dtype=BF)
    feats, dfe = rnd(16, 262144, 64, seed=12).abs(), torch.empty(16, 262144, 64, device=DEV, dtype=BF)
    dl = torch.zeros(1, device=DEV)

    def once():
        call("layernorm_fwd", x, gamma, beta, y, mean, rstd, R, D, 1e-5)
        call("layernorm_bwd_sums", dy, x, gamma, mean, rstd, add, y, dg, db, cs, None, 0, lnws, R, D)
        call("layernorm_bwd_sums", dy, x, gamma, mean, rstd, add, y, dg, db, None, rs, T, lnws, R, D)
        call("groupnorm_stats", gx, gws, gm, gr, N, HW, Cc, 32, 1e-6)
        call("groupnorm_apply", gx, gm, gr, gga, gbe, gy, N, HW, Cc, 32, 1)
        call("groupnorm_bwd", gdy, gx, gm, gr, gga, gbe, gws, gadd, gy, N, HW, Cc, 32, 1)
        call("adam_step", p, g, m, v, shadow, n_par, hyper)
        call("colsum", big, bsum, R, 4 * D)
        call("rowsum", big, rsum, B, 4 * T, D)
        patches, sv, _ = cut.forward(img, prm)
        cut.backward(sv, patches)
        call("mha_small_fwd", qkv, att, 512, 50, 12, 64, 0.125)
        call("mha_small_bwd", qkv, dout, dqkv, 512, 50, 12, 64, 0.125)
        call("conv_taps_gather", taps, None, raw, post, N, 256, 256, 3)
        call("upsample2x_fwd", up_in, up_out, N, 128, 128, 128)
        call("upsample2x_bwd", up_out, up_in, N, 128, 128, 128)
        call("diversity_tap", feats, dl, dfe, 2, 8, 262144, 64, -0.1)

    ops.reset_launch_count()
    once()
    torch.cuda.synchronize()
    n = ops.launch_count()
    if "--count" in sys.argv:
        print(n)
        return
    torch.cuda.cudart().cudaProfilerStart()
    once()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("zoo: %d ffvc launches per round" % n)


if __name__ == "__main__":
    main()
