"""Per-kernel parity of the HBM-bound ops (through the C ABI) against plain PyTorch fp32 on the same inputs.
bf16 tensors are compared against fp32 math on the bf16-rounded inputs; tolerance 2e-2 of the reference's max
magnitude unless stated (bf16 has 8 mantissa bits: one rounding of the output is 4e-3 relative)."""
import math

import pytest
import torch
import torch.nn.functional as F

from feed_forward_vqgan_clip_b200 import ops
from feed_forward_vqgan_clip_b200.ops import call

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32


def rnd(*shape, seed=0, dtype=BF, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def close(out, ref, tol=2e-2):
    out, ref = out.float(), ref.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, "max err %g vs scale %g" % (err, scale)


@pytest.mark.parametrize("v2", [0, 1, 2])
@pytest.mark.parametrize("rows,D,T", [(1000, 1024, 250), (77, 768, 11), (4096, 128, 256), (1030, 256, 103), (6, 512, 3),
                                      (16384, 1024, 256), (3, 1024, 1)])
def test_layernorm_fwd_bwd(rows, D, T, v2, ffvc_options):
    """v2 = 1 / 2: the column-owning kernels with 4 / 8 (fwd) and 4 / 2 (bwd) rows in flight (D in {256, 512, 768, 1024}; other D fall back to the warp-per-row form), including
    ragged row counts (rows % 4 != 0) and the fused bias-gradient sums of dx."""
    ffvc_options(ln_fwd_v2=v2, ln_bwd_v2=v2)
    x, dy = rnd(rows, D, seed=1), rnd(rows, D, seed=2)
    add = rnd(rows, D, seed=3)
    gamma = 1 + 0.1 * torch.randn(D, device=DEV)
    beta = 0.1 * torch.randn(D, device=DEV)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    call("layernorm_fwd", x, gamma, beta, y, mean, rstd, rows, D, 1e-5)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br, 1e-5)
    close(y, yr)
    assert torch.allclose(mean, x.float().mean(1), atol=1e-4)
    assert torch.allclose(rstd, (x.float().var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-3)
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    call("layernorm_bwd", dy, x, gamma, mean, rstd, add, dx, dg, db, rows, D)
    close(dx, xr.grad + add.float())
    close(dg, gr.grad, 1e-2)
    close(db, br.grad, 1e-2)
    dx2 = torch.empty_like(x)
    call("layernorm_bwd", dy, x, gamma, mean, rstd, None, dx2, None, None, rows, D)
    close(dx2, xr.grad)
    # the fused form: same dx / dgamma / dbeta, plus column sums and per-token row sums of the dx it stored
    from feed_forward_vqgan_clip_b200 import _lib
    ws = torch.full((int(_lib.load().ffvc_layernorm_bwd_ws_bytes(D, T)) // 4,), float("nan"), device=DEV)   # never read unwritten
    for with_w, with_add, with_col, with_row in [(1, 1, 1, 1), (1, 1, 1, 0), (1, 0, 0, 1), (0, 1, 1, 1), (0, 0, 0, 0)]:
        dx3 = torch.empty_like(x)
        dg3, db3 = 0.5 + torch.zeros(D, device=DEV), 0.25 + torch.zeros(D, device=DEV)     # accumulate semantics
        cs, rsum = 2.0 + torch.zeros(D, device=DEV), 3.0 + torch.zeros(T, device=DEV)
        call("layernorm_bwd_sums", dy, x, gamma, mean, rstd, add if with_add else None, dx3, dg3 if with_w else None,
             db3 if with_w else None, cs if with_col else None, rsum if with_row else None, T, ws, rows, D)
        close(dx3, xr.grad + (add.float() if with_add else 0))
        if with_w:
            close(dg3 - 0.5, gr.grad, 1e-2)
            close(db3 - 0.25, br.grad, 1e-2)
        ref_cs = dx3.float().sum(0)
        ref_rs = dx3.float().view(rows // T, T, D).sum((0, 2))
        tol = 1e-3 * dx3.float().abs().sum(0).max().item() + 1e-3
        if with_col:
            assert (cs - 2.0 - ref_cs).abs().max().item() <= tol
        if with_row:
            tol_r = 1e-3 * dx3.float().abs().view(rows // T, T, D).sum((0, 2)).max().item() + 1e-3
            assert (rsum - 3.0 - ref_rs).abs().max().item() <= tol_r


@pytest.mark.parametrize("N,HW,C,swish", [(2, 256, 512, 1), (3, 1024, 128, 1), (1, 4096, 64, 0), (2, 16384, 128, 1)])
def test_groupnorm_fwd_bwd(N, HW, C, swish):
    x, dy, add = rnd(N, HW, C, seed=1), rnd(N, HW, C, seed=2), rnd(N, HW, C, seed=3)
    gamma = 1 + 0.1 * torch.randn(C, device=DEV)
    beta = 0.1 * torch.randn(C, device=DEV)
    from feed_forward_vqgan_clip_b200 import _lib
    ws = torch.full((int(_lib.load().ffvc_groupnorm_ws_doubles(N, HW, 32)),), float("nan"), device=DEV, dtype=torch.float64)
    mean, rstd = torch.empty(N * 32, device=DEV), torch.empty(N * 32, device=DEV)
    call("groupnorm_stats", x, ws, mean, rstd, N, HW, C, 32, 1e-6)
    y = torch.empty_like(x)
    call("groupnorm_apply", x, mean, rstd, gamma, beta, y, N, HW, C, 32, swish)
    xr = x.float().permute(0, 2, 1).contiguous().requires_grad_(True)       # (N, C, HW)
    u = F.group_norm(xr, 32, gamma, beta, eps=1e-6)
    yr = u * torch.sigmoid(u) if swish else u
    close(y, yr.permute(0, 2, 1))
    yr.backward(dy.float().permute(0, 2, 1))
    dx = torch.empty_like(x)
    call("groupnorm_bwd", dy, x, mean, rstd, gamma, beta, ws, add, dx, N, HW, C, 32, swish)
    close(dx, xr.grad.permute(0, 2, 1) + add.float())
    # reproducible and batch-invariant statistics (fixed-order reductions, partial sums per sample shaped by HW only): the
    # same sample inside a batch of 2N samples gives bit-identical mean / rstd / dx, run after run
    x2, dy2, add2 = torch.cat([x.flip(0), x]), torch.cat([dy.flip(0), dy]), torch.cat([add.flip(0), add])
    ws2 = torch.full((int(_lib.load().ffvc_groupnorm_ws_doubles(2 * N, HW, 32)),), float("nan"), device=DEV, dtype=torch.float64)
    m2, r2 = torch.empty(2 * N * 32, device=DEV), torch.empty(2 * N * 32, device=DEV)
    call("groupnorm_stats", x2, ws2, m2, r2, 2 * N, HW, C, 32, 1e-6)
    assert torch.equal(m2[N * 32:], mean) and torch.equal(r2[N * 32:], rstd)
    dxx = torch.empty_like(x2)
    call("groupnorm_bwd", dy2, x2, m2, r2, gamma, beta, ws2, add2, dxx, 2 * N, HW, C, 32, swish)
    assert torch.equal(dxx[N:], dx)


@pytest.mark.parametrize("ring", [0, 1])
@pytest.mark.parametrize("pipeline", [1, 0])
@pytest.mark.parametrize("N,HW,C,swish", [(2, 256, 512, 1), (3, 1024, 128, 1), (1, 4096, 64, 0), (5, 16384, 128, 1), (2, 4096, 256, 1),
                                          (3, 100, 64, 1), (2, 65536, 128, 1)])
def test_groupnorm_fused_single_kernel_forms(N, HW, C, swish, pipeline, ring, ffvc_options):
    """the L2-resident single-kernel GroupNorm (persistent grid + per-sample arrival counters) against torch AND against
    the two-pass kernels; both schedules (phase A of sample n+1 before / after the wait for sample n); ring = 1: loads
    through per-thread cp.async rings (also a thread with more items than the ring is deep: 65536 pixels / 148 CTAs)"""
    from feed_forward_vqgan_clip_b200 import _lib
    ffvc_options(gn_ring=ring)
    x, dy, add = rnd(N, HW, C, seed=1), rnd(N, HW, C, seed=2), rnd(N, HW, C, seed=3)
    gamma = 1 + 0.1 * torch.randn(C, device=DEV)
    beta = 0.1 * torch.randn(C, device=DEV)
    _lib.load().ffvc_groupnorm_set_pipeline(pipeline)
    nbytes = _lib.load().ffvc_groupnorm_ws_bytes(N, 32)
    assert nbytes == N * 64 * 8 + N * 4
    ws = torch.empty((nbytes + 7) // 8, device=DEV, dtype=torch.float64)
    mean, rstd = torch.empty(N * 32, device=DEV), torch.empty(N * 32, device=DEV)
    y = torch.empty_like(x)
    call("groupnorm_fused_fwd", x, gamma, beta, y, mean, rstd, ws, N, HW, C, 32, swish, 1e-6)
    xr = x.float().permute(0, 2, 1).contiguous().requires_grad_(True)       # (N, C, HW)
    u = F.group_norm(xr, 32, gamma, beta, eps=1e-6)
    yr = u * torch.sigmoid(u) if swish else u
    close(y, yr.permute(0, 2, 1))
    mean2, rstd2, y2 = torch.empty_like(mean), torch.empty_like(rstd), torch.empty_like(x)
    ws_two = torch.empty(int(_lib.load().ffvc_groupnorm_ws_doubles(N, HW, 32)), device=DEV, dtype=torch.float64)
    call("groupnorm_stats", x, ws_two, mean2, rstd2, N, HW, C, 32, 1e-6)
    call("groupnorm_apply", x, mean2, rstd2, gamma, beta, y2, N, HW, C, 32, swish)
    assert torch.allclose(mean, mean2, atol=1e-5) and torch.allclose(rstd, rstd2, rtol=1e-4)
    assert (y.float() - y2.float()).abs().max().item() <= 2e-2 * y2.float().abs().max().item()
    yr.backward(dy.float().permute(0, 2, 1))
    for a in (add, None):
        dx = torch.empty_like(x)
        call("groupnorm_fused_bwd", dy, x, mean, rstd, gamma, beta, ws, a, dx, N, HW, C, 32, swish)
        close(dx, xr.grad.permute(0, 2, 1) + (a.float() if a is not None else 0))
    _lib.load().ffvc_groupnorm_set_pipeline(1)


def test_upsample_and_transpose():
    x = rnd(2, 8, 8, 64, seed=1)
    y = torch.empty(2, 16, 16, 64, device=DEV, dtype=BF)
    call("upsample2x_fwd", x, y, 2, 8, 8, 64)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    close(y, ref, 1e-6)
    dy = rnd(2, 16, 16, 64, seed=2)
    dx = torch.empty_like(x)
    call("upsample2x_bwd", dy, dx, 2, 8, 8, 64)
    close(dx, dy.float().view(2, 8, 2, 8, 2, 64).sum(dim=(2, 4)))
    a = rnd(3, 50, 70, seed=3)
    b = torch.empty(3, 70, 50, device=DEV, dtype=BF)
    call("transpose", a, b, 3, 50, 70, 0, 0)
    assert torch.equal(b, a.transpose(1, 2).contiguous())
    c = torch.empty(3, 70, 50, device=DEV, dtype=F32)
    call("transpose", a.float(), c, 3, 50, 70, 1, 1)
    assert torch.equal(c, a.float().transpose(1, 2).contiguous())


def test_softmax_fwd_bwd():
    rows, n = 512, 256
    s = rnd(rows, n, seed=1, dtype=F32, scale=3.0)
    p = torch.empty(rows, n, device=DEV, dtype=BF)
    call("softmax_fwd", s, p, rows, n, n)
    close(p, torch.softmax(s, -1))
    dp = rnd(rows, n, seed=2, dtype=F32)
    ds = torch.empty(rows, n, device=DEV, dtype=BF)
    call("softmax_bwd", p, dp, ds, rows, n, n, 0.5)
    pf = p.float()
    close(ds, 0.5 * pf * (dp - (pf * dp).sum(-1, keepdim=True)))


@pytest.mark.parametrize("rows,n", [(3000, 520), (16384, 1024), (4096, 4096), (64, 65536), (777, 8), (5, 6152), (1000, 30)])
def test_colsum_bias_grad(rows, n):
    dy = rnd(rows, n, seed=1)
    db = torch.full((n,), 0.5, device=DEV)                     # accumulates into the gradient arena
    call("colsum", dy, db, rows, n)
    close(db, dy.float().sum(0) + 0.5, 2e-3)


@pytest.mark.parametrize("B,J,D", [(4, 96, 128), (8, 256, 1024), (2, 33, 2056), (3, 17, 100)])
def test_rowsum_bias_grad(B, J, D):
    d3 = rnd(B, J, D, seed=2)
    dj = torch.zeros(J, device=DEV)
    call("rowsum", d3, dj, B, J, D)
    close(dj, d3.float().sum(dim=(0, 2)), 2e-3)


def test_bias_grads_and_casts():
    x = rnd(1000, seed=3, dtype=F32)
    y = torch.empty(1000, device=DEV, dtype=BF)
    call("cast_f32_bf16", x, y, 1000)
    assert torch.equal(y, x.to(BF))
    z = torch.empty(1000, device=DEV, dtype=F32)
    call("cast_bf16_f32", y, z, 1000)
    assert torch.equal(z, y.float())


@pytest.mark.parametrize("C,ncodes,P", [(256, 16384, 512), (64, 512, 300), (256, 16384, 4096)])
def test_vq_nearest_matches_reference_argmin(C, ncodes, P):
    g = torch.Generator().manual_seed(0)
    cb = torch.randn(ncodes, C, generator=g).to(DEV)
    z = (torch.randn(P, C, generator=g) * 1.5).to(DEV)
    lo, hi = float(cb.min()), float(cb.max())
    cbT = cb.t().contiguous()
    cn = torch.empty(ncodes, device=DEV)
    call("rownorm2", cb, cn, ncodes, C)
    idx = torch.empty(P, device=DEV, dtype=torch.int32)
    zq = torch.empty(P, C, device=DEV, dtype=BF)
    zq32 = torch.empty(P, C, device=DEV, dtype=F32)
    zc = torch.empty(P, C, device=DEV, dtype=F32)
    call("vq_nearest", z, cb, cbT, cn, idx, zq, zq32, zc, P, C, ncodes, lo, hi)
    zcl = z.clamp(lo, hi)
    assert torch.equal(zc, zcl)
    # the reference expression (main.py:135-136), evaluated in float64 to be the arbiter of near ties
    d = (zcl.double().pow(2).sum(-1, keepdim=True) + cb.double().pow(2).sum(1) - 2 * zcl.double() @ cb.double().T)
    ref = d.argmin(-1)
    agree = (idx.long() == ref).float().mean().item()
    assert agree >= 0.999, agree
    # where they differ the chosen code must be a numerical tie
    bad = (idx.long() != ref).nonzero().flatten()
    for r in bad.tolist():
        assert abs(d[r, idx[r]].item() - d[r, ref[r]].item()) < 1e-3
    assert torch.equal(zq32, cb[idx.long()])
    assert torch.equal(zq, cb[idx.long()].to(BF))
    # ---- the tensor-core search (bf16 hi/lo split GEMM, K = 3C, arg-min epilogue) must pick the same codes
    csplit = torch.empty(ncodes, 3 * C, device=DEV, dtype=BF)
    cn2 = torch.empty(ncodes, device=DEV)
    call("vq_prepare_codebook", cb, csplit, cn2, ncodes, C)
    assert torch.allclose(cn2, cn, rtol=1e-5)
    hi_ = cb.to(BF)
    assert torch.equal(csplit[:, :C], hi_) and torch.equal(csplit[:, 2 * C:], hi_)
    assert torch.equal(csplit[:, C:2 * C], (cb - hi_.float()).to(BF))
    zsplit = torch.empty(P, 3 * C, device=DEV, dtype=BF)
    keys = torch.empty(P, device=DEV, dtype=torch.int64)
    idx2 = torch.empty(P, device=DEV, dtype=torch.int32)
    zq2, zq322, zc2 = torch.empty_like(zq), torch.empty_like(zq32), torch.empty_like(zc)
    call("vq_nearest_tc", z, cb, csplit, cn2, zsplit, keys, idx2, zq2, zq322, zc2, P, C, ncodes, lo, hi)
    assert torch.equal(zc2, zcl)
    agree2 = (idx2.long() == ref).float().mean().item()
    assert agree2 >= 0.999, agree2
    for r in (idx2.long() != ref).nonzero().flatten().tolist():
        assert abs(d[r, idx2[r]].item() - d[r, ref[r]].item()) < 2e-3
    assert torch.equal(zq322, cb[idx2.long()])
    assert torch.equal(zq2, cb[idx2.long()].to(BF))


def test_clamp_bwd_truth_table():
    # SURVEY §8 a4 [probe]: x=-1,g=+1 -> 0; x=2,g=+1 -> 1 (passes); g=-1 passes at x=-1, blocked at x=2; inside passes
    x = torch.tensor([-1.0, 2.0, -1.0, 2.0, 0.5, 0.5], device=DEV)
    g = torch.tensor([1.0, 1.0, -1.0, -1.0, 1.0, -1.0], device=DEV)
    gx = torch.empty_like(x)
    call("clamp_bwd", g, x, gx, 6, 0.0, 1.0)
    assert gx.tolist() == [0.0, 1.0, -1.0, 0.0, 1.0, -1.0]


def test_image_post_and_conv_cin3():
    d = rnd(4096, 3, seed=1, dtype=F32, scale=1.5)
    xr = torch.empty_like(d)
    call("image_post_fwd", d, xr, d.numel())
    assert torch.allclose(xr, ((d + 1) / 2).clamp(0, 1))
    g = rnd(4096, 3, seed=2, dtype=F32)
    gd = torch.empty_like(d)
    call("image_post_bwd", g, d, gd, d.numel())
    u = (d + 1) / 2
    ref = torch.where(g * (u - u.clamp(0, 1)) >= 0, 0.5 * g, torch.zeros_like(g))
    assert torch.allclose(gd, ref)
    # dgrad of a 128 -> 3 conv == conv of the 3-channel gradient with flipped, transposed filters
    N, H, W, CO = 2, 32, 32, 128
    w = rnd(3, CO, 3, 3, seed=3, dtype=F32, scale=0.1)        # forward conv_out weight [3][128][3][3]
    gy = rnd(N, H, W, 3, seed=4, dtype=F32)
    wT = w.flip(2, 3).permute(1, 2, 3, 0).reshape(CO, 27).contiguous()
    out = torch.empty(N, H, W, CO, device=DEV, dtype=BF)
    call("conv3x3_cin3", gy, wT, out, N, H, W, CO)
    ref = F.conv_transpose2d(gy.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    close(out, ref)
    # the same dgrad as im2col (K = 27 padded to 32) + tcgen05 GEMM — the form the decoder uses
    col = torch.empty(N * H * W, 32, device=DEV, dtype=BF)
    call("im2col3x3_cin3", gy, col, N, H, W)
    wp = torch.zeros(CO, 32, device=DEV)
    wp[:, :27] = wT
    out2 = torch.empty(N * H * W, CO, device=DEV, dtype=BF)
    ops.gemm(col, wp.to(BF), out2, N * H * W, CO, 32)
    close(out2.view(N, H, W, CO), ref)


def test_adam_matches_torch():
    n = 10007
    p0 = rnd(n, seed=1, dtype=F32)
    p = p0.clone()
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    shadow = torch.empty(n, device=DEV, dtype=BF)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 1, 1, 0.5, 0, 0, 0, 0, 0, 0, 0, 0, 0], device=DEV, dtype=F32)
    for step in range(1, 4):
        g = rnd(n, seed=10 + step, dtype=F32)
        ref.grad = 0.5 * g
        opt.step()
        call("adam_tick", hyper)
        call("adam_step", p, g, m, v, shadow, n, hyper)
        assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(shadow, p.to(BF))


@pytest.mark.parametrize("n", [10008, 10007])
def test_optimizer_block_clip_cosine_ema(n):
    """the reference's optimizer block (main.py:591,693,702-705,833-837,843-844): clip_grad_norm_ -> Adam -> cosine
    scheduler -> torch_ema update, through FusedAdam's device-side scalars.  n % 4 == 0 takes the 16-byte kernel."""
    from types import SimpleNamespace
    from feed_forward_vqgan_clip_b200.train_step import FusedAdam
    T_MAX, CLIP, DECAY, WORLD = 6, 0.7, 0.9, 2
    p0 = rnd(n, seed=1, dtype=F32)
    par = torch.nn.Parameter(p0.clone())
    eng = SimpleNamespace(arena=par.data, grad=torch.zeros(n, device=DEV), shadow=torch.empty(n, device=DEV, dtype=BF),
                          total=n, dev=torch.device(DEV), params=[par], ext_shadow_fresh=False, _shadow_version=None)
    opt = FusedAdam(eng, lr=2e-3)
    opt.set_grad_scale(1.0 / WORLD)
    opt.set_clip_grad_norm(CLIP)
    opt.set_cosine(T_MAX)
    opt.enable_ema(DECAY)
    ref = p0.clone().cpu().requires_grad_(True)
    ropt = torch.optim.Adam([ref], lr=2e-3)
    rsch = torch.optim.lr_scheduler.CosineAnnealingLR(ropt, T_max=T_MAX, eta_min=0)
    ema, num_updates = ref.detach().clone(), 0
    for step in range(1, 6):
        g = rnd(n, seed=10 + step, dtype=F32) * (0.02 if step % 2 else 0.001)     # clipped on odd steps only
        eng.grad.copy_(g)
        opt.apply()
        ref.grad = (g / WORLD).cpu()
        torch.nn.utils.clip_grad_norm_([ref], CLIP)
        ropt.step()
        rsch.step()
        num_updates += 1                                            # torch_ema.ExponentialMovingAverage.update
        d = min(DECAY, (1 + num_updates) / (10 + num_updates))
        ema.sub_((1.0 - d) * (ema - ref.detach()))
        assert torch.allclose(par.data.cpu(), ref.detach(), rtol=2e-5, atol=2e-6), step
        assert torch.allclose(opt.ema.cpu(), ema, rtol=2e-5, atol=2e-6), step
    assert torch.equal(eng.shadow, par.data.to(BF))
    # checkpoint round trip in torch.optim.Adam's own format (opt.th, main.py:593-596,911)
    sd = opt.state_dict()
    rsd = ropt.state_dict()
    assert torch.allclose(sd["state"][0]["exp_avg"].cpu(), rsd["state"][0]["exp_avg"], rtol=1e-4, atol=1e-7)
    assert torch.allclose(sd["state"][0]["exp_avg_sq"].cpu(), rsd["state"][0]["exp_avg_sq"], rtol=1e-4, atol=1e-9)
    assert int(sd["state"][0]["step"]) == int(rsd["state"][0]["step"]) == 5
    fresh = torch.optim.Adam([torch.nn.Parameter(p0.clone())], lr=2e-3)
    fresh.load_state_dict(sd)                                       # torch accepts our dict
    opt2 = FusedAdam(eng, lr=2e-3)
    opt2.load_state_dict(rsd)                                       # and we accept torch's
    assert torch.allclose(opt2.m.cpu(), rsd["state"][0]["exp_avg"]) and float(opt2.hyper[8]) == 5.0


@pytest.mark.parametrize("N,T,Hh", [(6, 50, 12), (3, 64, 2), (2, 17, 3), (5, 1, 1), (2, 33, 4)])
def test_mha_small_fwd_bwd(N, T, Hh):
    """tensor-core (mma.sync) attention for short sequences; every 16-row tile count and ragged tails"""
    dh = 64
    W = Hh * dh
    qkv = rnd(N, T, 3 * W, seed=1, scale=0.7)
    out = torch.empty(N, T, W, device=DEV, dtype=BF)
    call("mha_small_fwd", qkv, out, N, T, Hh, dh, dh ** -0.5)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.split(W, dim=-1)
    q = q.reshape(N, T, Hh, dh).transpose(1, 2)
    k = k.reshape(N, T, Hh, dh).transpose(1, 2)
    v = v.reshape(N, T, Hh, dh).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, -1)
    o = (a @ v).transpose(1, 2).reshape(N, T, W)
    close(out, o)
    do = rnd(N, T, W, seed=2)
    o.backward(do.float())
    dqkv = torch.empty_like(qkv)
    call("mha_small_bwd", qkv, do, dqkv, N, T, Hh, dh, dh ** -0.5)
    close(dqkv, x.grad)


def test_spherical_loss():
    import oracle.loss as ol
    N, B, D = 24, 3, 512
    emb = rnd(N, D, seed=1, dtype=F32)
    tgt = rnd(B, D, seed=2, dtype=F32, scale=0.45)
    loss = torch.zeros(1, device=DEV)
    demb = torch.empty(N, D, device=DEV)
    call("spherical_loss", emb, tgt, loss, demb, None, N, B, D, 1.0)
    e = emb.cpu().requires_grad_(True)
    ref = ol.spherical_dist_loss(e, tgt.cpu(), N // B)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    close(demb.cpu(), e.grad, 1e-3)


@pytest.mark.parametrize("pool_v2", [0, 1])
def test_cutouts_fwd_bwd_vs_oracle(pool_v2, ffvc_options):
    import oracle.cutouts as oc
    from feed_forward_vqgan_clip_b200.cutouts import CutoutEngine, sample_params, params_to_device
    ffvc_options(pool_v2=pool_v2)
    B, H, cutn, P = 2, 256, 4, 224
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, H, H, generator=g)
    prm = sample_params(cutn * B, P, g)
    # make sure every augmentation is exercised at least once
    assert (prm["affine_inv"] != torch.eye(3)).any() and (prm["persp_inv"] != torch.eye(3)).any()
    xr = x.clone().requires_grad_(True)
    ref = oc.make_cutouts(xr, cutn, prm, P, normalize=True)
    eng = CutoutEngine(P, cutn, 32, torch.device(DEV))
    img = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    patches, saved, out = eng.forward(img, params_to_device(prm, DEV), want_image=True)
    # a handful of pixels sit exactly on a hue-sector / floor boundary and may flip: compare robustly
    diff = (out.cpu() - ref.detach()).abs()
    assert (diff > 2e-3).float().mean().item() < 1e-4, diff.max()
    N = cutn * B
    pm = patches.float().cpu().view(N, 7, 7, 3, 32, 32).permute(0, 3, 1, 4, 2, 5).reshape(N, 3, P, P)
    assert (pm - ref.detach()).abs().max().item() < 0.05      # bf16 rounding of values up to ~3
    gy = torch.randn(N, 3, P, P, generator=g)
    ref.backward(gy)
    dp = gy.view(N, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5).reshape(N, 49, 3072).contiguous().to(DEV).to(BF)
    dimg = eng.backward(saved, dp).cpu().permute(0, 3, 1, 2)
    err = (dimg - xr.grad).abs()
    scale = xr.grad.abs().max().item()
    assert (err > 3e-2 * scale).float().mean().item() < 1e-3, (err.max().item(), scale)


@pytest.mark.parametrize("B,H,P", [(2, 256, 224), (3, 32, 224), (1, 512, 224), (2, 100, 37)])
def test_cutout_pool_bwd_row_kernel_matches_flat_kernel_and_autograd(B, H, P, ffvc_options):
    g = torch.Generator().manual_seed(11)
    x = torch.rand(B, H, H, 3, generator=g).to(DEV)
    dy = torch.randn(B, P, P, 3, generator=g).to(DEV)
    outs = []
    for v2 in (0, 1):
        ffvc_options(pool_v2=v2)
        dx = torch.full((B, H, H, 3), 7.0, device=DEV)
        call("cutout_pool_bwd", x, dy, dx, B, H, H, P, 0)
        outs.append(dx)
    assert torch.allclose(outs[0], outs[1], rtol=1e-6, atol=1e-7)       # same arithmetic per element
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)          # main.py:218 on NCHW
    yr = (F.adaptive_avg_pool2d(xr, P) + F.adaptive_max_pool2d(xr, P)) / 2
    yr.backward(dy.permute(0, 3, 1, 2))
    close(outs[1].permute(0, 3, 1, 2), xr.grad, 1e-5)


def test_tv_loss_fwd_bwd_vs_reference_expression():
    import oracle.loss as ol
    B, H, W = 2, 24, 40
    g = torch.Generator().manual_seed(7)
    img = torch.rand(B, 3, H, W, generator=g)
    xr = img.clone().requires_grad_(True)
    ref = 0.3 * ol.tv_loss(xr)                               # main.py:423-428
    ref.backward()
    nhwc = img.permute(0, 2, 3, 1).contiguous().to(DEV)
    loss = torch.zeros(1, device=DEV)
    dimg = torch.ones(B, H, W, 3, device=DEV)                # accumulates on top of an existing gradient
    call("tv_loss", nhwc, loss, dimg, B, H, W, 3, 0.3)
    assert abs(loss.item() - ref.item()) < 1e-6
    assert torch.allclose(dimg.cpu() - 1.0, xr.grad.permute(0, 2, 3, 1), atol=1e-7)
    x = rnd(1000, seed=8, dtype=F32)
    y = rnd(1000, seed=9, dtype=F32)
    y0 = y.clone()
    call("axpy_f32", x, y, 0.25, 1000)
    assert torch.allclose(y, y0 + 0.25 * x)


def test_sln_and_vitgan_attention_kernels():
    R, D = 96, 128
    n, w, ds = rnd(R, D, seed=1), rnd(R, D, seed=2), rnd(R, D, seed=3)
    gamma, beta = torch.tensor([0.7], device=DEV), torch.tensor([-0.4], device=DEV)
    s = torch.empty_like(n)
    call("sln_mod_fwd", n, w, gamma, beta, s, R * D)
    nf, wf = n.float().requires_grad_(True), w.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    sr = gr * wf * nf + br * wf                               # vitgan.py:21
    close(s, sr)
    sr.backward(ds.float())
    dn = torch.empty_like(n)
    dw = torch.zeros(R, D, device=DEV)
    dg, db = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
    call("sln_mod_bwd", ds, n, w, gamma, beta, dn, dw, dg, db, R * D)
    close(dn, nf.grad)
    close(dw, wf.grad, 1e-3)
    assert abs(dg.item() - gr.grad.item()) < 1e-2 * abs(gr.grad.item()) + 1e-2
    assert abs(db.item() - br.grad.item()) < 1e-2 * abs(br.grad.item()) + 1e-2
    # attention with the '(d k h)' interleaved layout (vitgan.py:82)
    B, T, H, dh = 3, 16, 6, 21
    Wd = H * dh
    Q3, Wp = (3 * Wd + 7) // 8 * 8, (Wd + 7) // 8 * 8
    qkv = torch.zeros(B, T, Q3, device=DEV, dtype=BF)
    qkv[..., :3 * Wd] = rnd(B, T, 3 * Wd, seed=4)
    out = torch.zeros(B, T, Wp, device=DEV, dtype=BF)
    probs = torch.empty(B * H, T, T, device=DEV)
    scale = 128 ** -0.5
    call("vitgan_attn_fwd", qkv, out, probs, B, T, H, dh, Q3, Wp, scale)
    x = qkv[..., :3 * Wd].float().requires_grad_(True)
    q, k, v = x.view(B, T, dh, 3, H).permute(3, 0, 4, 1, 2)
    a = torch.softmax(torch.einsum("bhid,bhjd->bhij", q, k) * scale, -1)
    r = torch.einsum("bhij,bhjd->bhid", a, v).permute(0, 2, 1, 3).reshape(B, T, Wd)
    close(out[..., :Wd], r)
    do = torch.zeros(B, T, Wp, device=DEV, dtype=BF)
    do[..., :Wd] = rnd(B, T, Wd, seed=5)
    r.backward(do[..., :Wd].float())
    dqkv = torch.zeros(B, T, Q3, device=DEV, dtype=BF)
    call("vitgan_attn_bwd", qkv, probs, do, dqkv, B, T, H, dh, Q3, Wp, scale)
    close(dqkv[..., :3 * Wd], x.grad)
    src = rnd(40, 126, seed=6, dtype=F32)
    dst = torch.empty(40, 128, device=DEV, dtype=BF)
    call("cast_f32_bf16_pitched", src, dst, 40, 126, 128)
    assert torch.equal(dst[:, :126], src.to(BF)) and float(dst[:, 126:].abs().max()) == 0.0


@pytest.mark.parametrize("R,C,HW", [(2, 64, 50), (4, 512, 9), (5, 128, 33), (16, 512, 7), (7, 40, 21), (2, 128, 41), (3, 256, 12),
                                    (2, 512, 10), (4, 64, 33), (3, 128, 2500)])
def test_diversity_tap_kernels_vs_reference_expression(R, C, HW):
    """main.py:779-787 on one tap: R <= 4 takes the in-register kernel, R > 4 (mode 'all': R = batch, B = 1; or repeat > 4)
    the streaming kernel; value and gradient w.r.t. the features"""
    import oracle.lpips as ol
    for B in (1, 3):
        f = rnd(R * B, HW, C, seed=R + B).abs() + 0.05                      # post-ReLU features
        fr = f.float().cpu().permute(0, 2, 1).reshape(R * B, C, HW, 1).clone().requires_grad_(True)
        a = ol.normalize_tensor(fr)
        div = ((a.view(R, 1, B, C, HW, 1) - a.view(1, R, B, C, HW, 1)) ** 2).sum(dim=3).mean()
        (-0.6 * div).backward()
        if B == 1:                                                          # 'all' is the same expression with B = 1
            div_all = ((a.view(R, 1, C, HW, 1) - a.view(1, R, C, HW, 1)) ** 2).sum(dim=2).mean()
            assert abs(div_all.item() - div.item()) < 1e-6
        loss = torch.zeros(1, device=DEV)
        d = torch.empty_like(f)
        call("diversity_tap", f, loss, d, R, B, HW, C, -0.6)
        assert abs(loss.item() + 0.6 * div.item()) < 2e-3 * abs(0.6 * div.item()) + 1e-6
        close(d.float().cpu().view(R * B, HW, C), fr.grad.view(R * B, C, HW).permute(0, 2, 1), 2e-2)


def test_lpips_diversity_engine_vs_oracle():
    """VGG16 taps + normalize_tensor + pairwise differences (main.py:776-782): value and gradient w.r.t. the image"""
    import oracle.lpips as ol
    from feed_forward_vqgan_clip_b200.lpips import LpipsVGG16
    from feed_forward_vqgan_clip_b200.cutouts import CLIP_MEAN, CLIP_STD
    sd = {k: (v.to(BF).float() if v.dim() == 4 else v) for k, v in ol.init_vgg_state_dict(seed=3).items()}
    net = LpipsVGG16()
    net.load_state_dict(sd)
    eng = net.to(DEV).engine()
    R, bs, H = 2, 1, 256
    g = torch.Generator().manual_seed(4)
    xr = torch.rand(R * bs, 3, H, H, generator=g)
    mean, std = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1), torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    xo = xr.clone().requires_grad_(True)
    div = ol.diversity(sd, xo, R, bs, mean, std)
    (-0.7 * div).backward()
    img = xr.permute(0, 2, 3, 1).contiguous().to(DEV)
    dimg = torch.zeros_like(img)
    loss = torch.zeros(1, device=DEV)
    eng.forward_backward(img, R, bs, 0.7, dimg, loss)
    assert abs(loss.item() - (-0.7 * div.item())) < 3e-2 * abs(0.7 * div.item()), (loss.item(), div.item())
    ref = xo.grad.permute(0, 2, 3, 1)
    mine = dimg.cpu()
    cs = float((mine * ref).sum() / (mine.norm() * ref.norm() + 1e-30))
    assert cs > 0.98, cs


def test_maxpool_and_relu_epilogue():
    x = rnd(2, 8, 8, 64, seed=1)
    y = torch.empty(2, 4, 4, 64, device=DEV, dtype=BF)
    call("maxpool2x2_fwd", x, y, 2, 8, 8, 64)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(y.float(), yr.permute(0, 2, 3, 1))
    dy = rnd(2, 4, 4, 64, seed=2)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    dx = torch.empty_like(x)
    call("maxpool2x2_bwd", x, dy, dx, 2, 8, 8, 64)
    assert torch.equal(dx.float(), xr.grad.permute(0, 2, 3, 1))
    a, b = rnd(256, 64, seed=3), rnd(128, 64, seed=4)
    out = torch.empty(256, 128, device=DEV, dtype=BF)
    ops.gemm(a, b, out, 256, 128, 64, act=ops.ACT_RELU)
    close(out, torch.relu(a.float() @ b.float().t()))
    aux = rnd(256, 128, seed=5)
    out2 = torch.empty(256, 128, device=DEV, dtype=BF)
    ops.gemm(a, b, out2, 256, 128, 64, aux=aux, mul_mode=ops.ACT_RELU)
    close(out2, (a.float() @ b.float().t()) * (aux.float() > 0))


def test_spherical_loss_with_input_loss_term_and_normalize_rows():
    """ffvc_spherical_loss2 (main.py:801-824: target term + input_loss_coef * source term) and ffvc_normalize_rows (main.py:734-735)
    against the reference's expressions"""
    g = torch.Generator().manual_seed(4)
    N, B, D = 24, 6, 512
    e = torch.randn(N, D, generator=g).to(DEV)
    t1, t2 = (torch.randn(B, D, generator=g) * 3).to(DEV), (torch.randn(B, D, generator=g) * 0.2).to(DEV)
    loss, de = torch.zeros(1, device=DEV), torch.empty(N, D, device=DEV)
    call("spherical_loss2", e, t1, t2, loss, de, None, N, B, D, 1.0, 0.37)
    er = e.clone().requires_grad_(True)
    en = F.normalize(er, dim=1)
    ref = sum(c * (F.normalize(t.repeat(N // B, 1), dim=-1).sub(en).norm(dim=-1).div(2).arcsin().pow(2).mul(2)).mean()
              for t, c in ((t1, 1.0), (t2, 0.37)))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert torch.allclose(de, er.grad, atol=1e-7, rtol=1e-3)
    x = (torch.randn(37, 96, generator=g) * 5).to(DEV)
    x[3] = 0                                             # F.normalize's eps path
    y = torch.empty_like(x)
    call("normalize_rows", x, y, 37, 96)
    assert torch.allclose(y, F.normalize(x, dim=1), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("n,h,w,c,co", [(2, 32, 32, 64, 3), (1, 24, 40, 128, 3), (3, 8, 16, 64, 1)])
def test_conv_out_as_one_gemm_plus_tap_gather(n, h, w, c, co):
    """the decoder's conv_out (3x3, pad 1, <= 3 output channels): ffvc_gemm into 9 * co tap columns + ffvc_conv_taps_gather against
    F.conv2d, and the fused image post-processing against clamp((y + 1) / 2, 0, 1) (main.py:142)"""
    g = torch.Generator().manual_seed(6)
    a = torch.randn(n, h, w, c, generator=g).to(BF)
    wt = (torch.randn(co, c, 3, 3, generator=g) * (9 * c) ** -0.5).to(BF)
    bias = torch.randn(co, generator=g)
    wv = torch.zeros(32, c)
    wv[:9 * co] = wt.float().permute(2, 3, 0, 1).reshape(9 * co, c)
    taps = torch.empty(n * h * w, 32, device=DEV, dtype=torch.float32)
    ops.gemm(a.to(DEV), wv.to(BF).to(DEV), taps, n * h * w, 32, c)
    y = torch.full((n * h * w, co), float("nan"), device=DEV)
    xr = torch.full((n * h * w, co), float("nan"), device=DEV)
    call("conv_taps_gather", taps, bias.to(DEV), y, xr, n, h, w, co)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(n * h * w, co)
    assert torch.allclose(y.cpu(), ref, atol=2e-4, rtol=2e-4), (y.cpu() - ref).abs().max()
    assert torch.equal(xr, ((y + 1) / 2).clamp(0, 1))
    y2 = torch.empty_like(y)
    call("conv_taps_gather", taps, None, y2, None, n, h, w, co)
    assert torch.allclose(y2.cpu() + bias, y.cpu(), atol=1e-6)


@pytest.mark.parametrize("N,T,heads,causal", [(2, 1024, 6, 1), (3, 200, 2, 0), (1, 130, 3, 1), (2, 64, 2, 1), (1, 257, 12, 0)])
def test_mha_flash_fwd_bwd(N, T, heads, causal):
    """tiled attention (ffvc_mha_flash_*) against torch: x-transformer shape (1024 tokens, 6 heads, causal), ragged lengths, and
    agreement with the small-sequence kernel at T = 64"""
    W = heads * 64
    qkv = rnd(N, T, 3 * W, seed=1)
    dout = rnd(N, T, W, seed=2)
    out = torch.full((N, T, W), float("nan"), device=DEV, dtype=BF)
    lse = torch.full((N, heads, T), float("nan"), device=DEV)
    call("mha_flash_fwd", qkv, out, lse, N, T, heads, 64, 0.125, causal)
    x = qkv.float().clone().requires_grad_(True)
    q, k, v = (x.view(N, T, 3, heads, 64)[:, :, i].transpose(1, 2) for i in range(3))
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool, device=DEV).triu_(1), float("-inf"))
    ref = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(N, T, W)
    close(out, ref, 2e-2)
    assert torch.allclose(lse, torch.logsumexp(s, dim=-1) * 1.4426950408889634, atol=2e-3, rtol=1e-4)
    ref.backward(dout.float())
    dqkv = torch.full((N, T, 3 * W), float("nan"), device=DEV, dtype=BF)
    call("mha_flash_bwd", qkv, out, dout, lse, torch.empty_like(lse), dqkv, N, T, heads, 64, 0.125, causal)
    close(dqkv, x.grad, 2e-2)
    if T <= 64 and not causal:
        o2 = torch.empty_like(out)
        call("mha_small_fwd", qkv, o2, N, T, heads, 64, 0.125)
        close(out, o2.float(), 1e-2)
