"""Parity of the tcgen05 GEMM (ffvc_gemm, through the C ABI) against a plain PyTorch fp32 matmul of the same
bf16 operands.  Tolerance: fp32 accumulation of bf16 products -> differences only from summation order and the
bf16 rounding of the output: |err| <= 2e-2 * max|ref| (bf16 out) / 1e-3 (fp32 out)."""
import pytest
import torch
import torch.nn.functional as F

from feed_forward_vqgan_clip_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rand(*shape, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV).to(torch.bfloat16)


def _check(out, ref, tol):
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    assert err <= tol * scale, "max err %g vs scale %g" % (err, scale)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (16384, 1024, 512),
                                   (200, 136, 72), (64, 768, 3072), (128, 32, 128), (384, 96, 200)])
def test_kmajor_plain(M, N, K):
    a, b = _rand(M, K, seed=1), _rand(N, K, seed=2)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K)
    _check(out, a.float() @ b.float().t(), 2e-2)


@pytest.mark.parametrize("bn", [32, 64, 128, 256])
def test_block_n_variants_fp32(bn):
    M, N, K = 256, 512, 320
    a, b = _rand(M, K, seed=3), _rand(N, K, seed=4)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, b, out, M, N, K, block_n=bn)
    _check(out, a.float() @ b.float().t(), 1e-3)


def test_b_mn_major():  # dgrad form: dX[m,k] = sum_n dY[m,n] W[n,k]
    M, N, K = 256, 1024, 512  # output (M x K), contraction N
    dy, w = _rand(M, N, seed=5), _rand(N, K, seed=6)
    out = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    ops.gemm(dy, w, out, M, K, N, b_mode=ops.MNMAJOR)
    _check(out, dy.float() @ w.float(), 2e-2)


def test_ab_mn_major():  # wgrad form: dW[n,k] = sum_m dY[m,n] X[m,k]
    M, N, K = 1024, 384, 256
    dy, x = _rand(M, N, seed=7), _rand(M, K, seed=8)
    out = torch.empty(N, K, device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, out, N, K, M, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR)
    _check(out, dy.float().t() @ x.float(), 1e-3)


def test_a_mn_major_only():
    M, N, K = 256, 128, 192  # A stored [K][M]
    at, b = _rand(K, M, seed=9), _rand(N, K, seed=10)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(at, b, out, M, N, K, a_mode=ops.MNMAJOR)
    _check(out, at.float().t() @ b.float().t(), 1e-3)


def test_batched_token_mix_form():
    # U[b,j,d] = sum_t W[j,t] H[b,t,d] + bias[j]   (mlp_mixer_pytorch.py:34, Conv1d k=1 over tokens)
    Bt, T, D, J = 3, 256, 512, 1024
    w, h = _rand(J, T, seed=11), _rand(Bt, T, D, seed=12)
    bias = torch.randn(J, device=DEV)
    out = torch.empty(Bt, J, D, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    ops.gemm(w, h, out, J, D, T, b_mode=ops.MNMAJOR, b_role=ops.ROLE_OUT, b_bs=T * D, b_ld=D, batch=Bt,
             out_bs=J * D, bias=bias, bias_mode=2, act=ops.ACT_GELU, pre_out=pre)
    ref_pre = torch.einsum("jt,btd->bjd", w.float(), h.float()) + bias[None, :, None]
    _check(pre, ref_pre, 2e-2)
    _check(out, F.gelu(ref_pre), 2e-2)


def test_segmented_k_split_atomic():
    # dW[j,t] = sum_{b,d} dU[b,j,d] H[b,t,d]  (token-mix wgrad): contraction over (segment b, k d)
    Bt, T, D, J = 8, 256, 256, 512
    du, h = _rand(Bt, J, D, seed=13), _rand(Bt, T, D, seed=14)
    out = torch.zeros(J, T, device=DEV, dtype=torch.float32)
    ops.gemm(du, h, out, J, T, D, a_role=ops.ROLE_SEG, a_bs=J * D, b_role=ops.ROLE_SEG, b_bs=T * D, k_segs=Bt,
             splits=4, atomic=True)
    _check(out, torch.einsum("bjd,btd->jt", du.float(), h.float()), 1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 256, 512), (1, 32, 32, 128, 256), (1, 128, 128, 128, 128),
                                            (1, 256, 256, 64, 32), (3, 16, 16, 64, 3)])
def test_conv3x3(n, h, w, cin, cout):
    x = _rand(n, h, w, cin, seed=15)
    wt = (_rand(cout, cin, 3, 3, seed=16).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()  # [Cout][tap][Cin]
    out = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, n * h * w, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(n, h, w, cin), bias=bias)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1)
    _check(out, ref, 2e-2)


def test_epilogue_residual_mulgrad_alpha():
    M, N, K = 256, 256, 128
    a, b = _rand(M, K, seed=17), _rand(N, K, seed=18)
    aux, res = _rand(M, N, seed=19), _rand(M, N, seed=20)
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K, bias=bias, aux=aux, mul_mode=ops.ACT_QUICKGELU, res=res, alpha=0.5)
    x = aux.float()
    s = torch.sigmoid(1.702 * x)
    ref = (0.5 * (a.float() @ b.float().t()) + bias) * (s + 1.702 * x * s * (1 - s)) + res.float()
    _check(out, ref, 2e-2)


@pytest.mark.parametrize("M,N,K", [(1024, 128, 256), (512, 64, 192), (300, 100, 64)])
def test_tile_m_256_kmajor(M, N, K):
    a, b = _rand(M, K, seed=21), _rand(N, K, seed=22)
    bias = torch.randn(N, device=DEV)
    res = _rand(M, N, seed=23)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, out, M, N, K, tile_m=256, block_n=128 if N > 64 else 64, bias=bias, res=res, act=ops.ACT_GELU)
    _check(out, F.gelu(a.float() @ b.float().t() + bias) + res.float(), 2e-2)


def test_tile_m_256_mn_major_wgrad():
    M, N, K = 640, 512, 128        # dW[n,k] = sum_m dY[m,n] X[m,k]  -> out (N x K), contraction M
    dy, x = _rand(M, N, seed=24), _rand(M, K, seed=25)
    out = torch.zeros(N, K, device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, out, N, K, M, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=N, b_ld=K, tile_m=256, block_n=128, atomic=True)
    _check(out, dy.float().t() @ x.float(), 1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 128, 128), (1, 256, 256, 64, 64), (2, 64, 64, 64, 128)])
def test_conv3x3_tile_m_256(n, h, w, cin, cout):
    x = _rand(n, h, w, cin, seed=26)
    wt = (_rand(cout, cin, 3, 3, seed=27).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    res = _rand(n, h, w, cout, seed=28)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, n * h * w, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(n, h, w, cin), bias=bias, res=res, tile_m=256,
             block_n=min(cout, 128))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    _check(out, ref, 2e-2)


def test_gelu_epilogue_matches_exact_erf_gelu():
    M, N, K = 256, 512, 64
    a, b = _rand(M, K, seed=29), (_rand(N, K, seed=30).float() * 0.5).to(torch.bfloat16)
    out = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, b, out, M, N, K, act=ops.ACT_GELU)
    ref = F.gelu(a.float() @ b.float().t())                 # exact (erf) GELU, nn.GELU() default
    assert (out - ref).abs().max().item() < 1e-4     # packed-polynomial CDF: |gelu error| <= 5.6e-5 (ptx.cuh: gelu_phi2)
    aux = _rand(M, N, seed=31)
    out2 = torch.empty(M, N, device=DEV, dtype=torch.float32)
    ops.gemm(a, b, out2, M, N, K, aux=aux, mul_mode=ops.ACT_GELU)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
    assert ((out2 - (a.float() @ b.float().t()) * gp).abs().max().item()) < 5e-4


def test_batch_inner_attention_heads():
    # scores[n,h,i,j] = sum_d Q[n,i,h,d] K[n,j,h,d] over a fused [N][T][3W] qkv tensor (CLIP ViT attention layout)
    Nn, T, Hh, dh = 5, 50, 12, 64
    W = Hh * dh
    qkv = _rand(Nn, T, 3 * W, seed=32)
    S = torch.zeros(Nn, Hh, T, 64, device=DEV, dtype=torch.float32)
    ops.gemm(qkv, qkv, S, T, T, dh, a_ld=3 * W, b_ld=3 * W, b_off=W, a_role=ops.ROLE_OUT, b_role=ops.ROLE_OUT,
             batch=Nn * Hh, batch_inner=Hh, a_bs=T * 3 * W, b_bs=T * 3 * W, a_bs_in=dh, b_bs_in=dh, ldc=64,
             out_bs=Hh * T * 64, out_bs_in=T * 64, alpha=0.125, block_n=64)
    q = qkv[..., :W].float().reshape(Nn, T, Hh, dh).transpose(1, 2)
    k = qkv[..., W:2 * W].float().reshape(Nn, T, Hh, dh).transpose(1, 2)
    ref = 0.125 * q @ k.transpose(-1, -2)
    _check(S[..., :T], ref, 1e-3)
    assert float(S[..., T:].abs().max()) == 0.0
    # O[n,i,h,:] = sum_j P[n,h,i,j] V[n,j,h,:]  (B read MN-major from the same fused tensor, K = 50 zero-extended to 64)
    P = torch.softmax(ref, -1)
    Pp = torch.zeros(Nn, Hh, T, 64, device=DEV, dtype=torch.bfloat16)
    Pp[..., :T] = P.to(torch.bfloat16)
    O = torch.empty(Nn, T, W, device=DEV, dtype=torch.bfloat16)
    ops.gemm(Pp, qkv, O, T, dh, T, a_ld=64, b_mode=ops.MNMAJOR, b_ld=3 * W, b_off=2 * W, a_role=ops.ROLE_OUT,
             b_role=ops.ROLE_OUT, batch=Nn * Hh, batch_inner=Hh, a_bs=Hh * T * 64, a_bs_in=T * 64, b_bs=T * 3 * W, b_bs_in=dh,
             ldc=W, out_bs=T * W, out_bs_in=dh, block_n=64)
    v = qkv[..., 2 * W:].float().reshape(Nn, T, Hh, dh).transpose(1, 2)
    refo = (Pp[..., :T].float() @ v).transpose(1, 2).reshape(Nn, T, W)
    _check(O, refo, 2e-2)


# ------------------------------------------------------------------ CTA-pair (cta_group::2) kernel, forced
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 256, 256), (1024, 1024, 512), (300, 200, 136), (4096, 128, 1024)])
def test_two_cta_kmajor(M, N, K):
    a, b = _rand(M, K, seed=41), _rand(N, K, seed=42)
    bias = torch.randn(N, device=DEV)
    res = _rand(M, N, seed=43)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    ops.gemm(a, b, out, M, N, K, two_cta=1, bias=bias, res=res, act=ops.ACT_GELU, pre_out=pre)
    ref_pre = a.float() @ b.float().t() + bias
    _check(pre, ref_pre, 2e-2)
    _check(out, F.gelu(ref_pre) + res.float(), 2e-2)


def test_two_cta_mn_major_forms():
    M, N, K = 512, 1024, 256        # dgrad: dX[m,k] = sum_n dY[m,n] W[n,k]
    dy, w = _rand(M, N, seed=44), _rand(N, K, seed=45)
    out = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    ops.gemm(dy, w, out, M, K, N, b_mode=ops.MNMAJOR, b_ld=K, two_cta=1)
    _check(out, dy.float() @ w.float(), 2e-2)
    M, N, K = 1024, 512, 384        # wgrad: dW[n,k] = sum_m dY[m,n] X[m,k]
    dy, x = _rand(M, N, seed=46), _rand(M, K, seed=47)
    o2 = torch.zeros(N, K, device=DEV, dtype=torch.float32)
    ops.gemm(dy, x, o2, N, K, M, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=N, b_ld=K, atomic=True, splits=2, two_cta=1)
    _check(o2, dy.float().t() @ x.float(), 1e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 256, 256), (1, 64, 64, 128, 128), (1, 256, 256, 64, 128)])
def test_two_cta_conv3x3(n, h, w, cin, cout):
    x = _rand(n, h, w, cin, seed=48)
    wt = (_rand(cout, cin, 3, 3, seed=49).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    res = _rand(n, h, w, cout, seed=50)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, n * h * w, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(n, h, w, cin), bias=bias, res=res, two_cta=1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    _check(out, ref, 2e-2)


def test_two_cta_batched_token_mix():
    Bt, T, D, J = 4, 256, 512, 1024
    w, h = _rand(J, T, seed=51), _rand(Bt, T, D, seed=52)
    bias = torch.randn(J, device=DEV)
    out = torch.empty(Bt, J, D, device=DEV, dtype=torch.bfloat16)
    ops.gemm(w, h, out, J, D, T, b_mode=ops.MNMAJOR, b_role=ops.ROLE_OUT, b_bs=T * D, b_ld=D, batch=Bt, out_bs=J * D, bias=bias,
             bias_mode=2, act=ops.ACT_GELU, two_cta=1)
    ref = F.gelu(torch.einsum("jt,btd->bjd", w.float(), h.float()) + bias[None, :, None])
    _check(out, ref, 2e-2)


@pytest.mark.parametrize("epi", [0, 8])
@pytest.mark.parametrize("two", [-1, 1])
def test_epilogue_warp_variants_agree(epi, two):
    """activation epilogues run on 16 warps (16-column chunks) by default; epi_warps=8 forces the 8-warp / 32-column form"""
    M, N, K = 512, 256, 128
    a, b = _rand(M, K, seed=61), _rand(N, K, seed=62)
    bias = torch.randn(N, device=DEV)
    aux, res = _rand(M, N, seed=63), _rand(M, N, seed=64)
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    pre = torch.empty_like(out)
    ops.gemm(a, b, out, M, N, K, bias=bias, act=ops.ACT_QUICKGELU, pre_out=pre, res=res, two_cta=two, epi_warps=epi)
    ref_pre = a.float() @ b.float().t() + bias
    _check(pre, ref_pre, 2e-2)
    _check(out, ref_pre * torch.sigmoid(1.702 * ref_pre) + res.float(), 2e-2)
    out2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(a, b, out2, M, N, K, aux=aux, mul_mode=ops.ACT_GELU, two_cta=two, epi_warps=epi, tile_m=0)
    x = aux.float()
    gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
    _check(out2, (a.float() @ b.float().t()) * gp, 2e-2)


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 4, 128, 64, 128), (2, 16, 128, 128, 128), (1, 6, 256, 128, 64), (2, 128, 128, 256, 128)])
def test_conv3x3_halo_reuse(n, h, w, cin, cout):
    from feed_forward_vqgan_clip_b200.ops import call
    x = _rand(n, h, w, cin, seed=71)
    wt = (_rand(cout, cin, 3, 3, seed=72).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    res = _rand(n, h, w, cout, seed=73)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.empty(n, h, w, cout, device=DEV, dtype=torch.bfloat16)
    call("conv3x3_halo", x, wp, out, n, h, w, cin, cout, cout, bias, res, None, 0, 0, 0)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1) + res.float()
    _check(out, ref, 2e-2)
    # and it agrees with the tap-by-tap implicit GEMM bit-for-bit up to accumulation order
    out2 = torch.empty_like(out)
    ops.gemm(x, wp, out2, n * h * w, cout, 9 * cin, a_mode=ops.CONV3X3, conv=(n, h, w, cin), bias=bias, res=res)
    _check(out, out2.float(), 1e-2)


def test_conv3x3_halo_three_channel_fp32_out():
    """the decoder's conv_out (128 -> 3 channels, fp32 image) on the halo kernel: N = 3 inside a 32-column tile, scalar stores"""
    from feed_forward_vqgan_clip_b200.ops import call
    n, h, w, cin, cout = 2, 8, 128, 128, 3
    x = _rand(n, h, w, cin, seed=81)
    wt = (_rand(cout, cin, 3, 3, seed=82).float() * 0.05).to(torch.bfloat16)
    bias = torch.randn(cout, device=DEV)
    wp = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()
    out = torch.full((n, h, w, cout), 7.0, device=DEV, dtype=torch.float32)
    call("conv3x3_halo", x, wp, out, n, h, w, cin, cout, cout, bias, None, None, 0, 0, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), bias, padding=1).permute(0, 2, 3, 1)
    _check(out, ref, 2e-2)


@pytest.mark.parametrize("M,N,K,two", [(300, 520, 192, -1), (1024, 2048, 128, 1), (257, 16384, 64, 0)])
def test_argmin_epilogue(M, N, K, two):
    """per-row arg-min of alpha * A.B^T + bias[n] without storing the product (the VQ distance search); ties -> lowest n"""
    a, b = _rand(M, K, seed=91), _rand(N, K, seed=92)
    bias = torch.randn(N, device=DEV)
    keys = torch.full((M,), -1, device=DEV, dtype=torch.int64)          # 0xFF.. = +inf key
    ops.gemm(a, b, None, M, N, K, bias=bias, alpha=-2.0, argmin_out=keys, two_cta=two)
    v = -2.0 * (a.float() @ b.float().t()) + bias
    idx = (keys & 0xFFFFFFFF).long()
    ref = v.argmin(-1)
    picked, best = v.gather(1, idx[:, None])[:, 0], v.min(-1).values
    assert (idx == ref).float().mean().item() > 0.99
    assert (picked - best).abs().max().item() < 1e-2 * max(1.0, best.abs().max().item())   # any disagreement is a near tie
    # a duplicated column must lose to its first occurrence
    b2 = b.clone()
    b2[N - 1] = b2[3]
    bias2 = bias.clone()
    bias2[N - 1] = bias2[3]
    keys.fill_(-1)
    ops.gemm(a, b2, None, M, N, K, bias=bias2, alpha=-2.0, argmin_out=keys, two_cta=two)
    assert int(((keys & 0xFFFFFFFF) == N - 1).sum()) == 0


@pytest.mark.parametrize("cfg", ["gelu_bias_pre", "gelu_rowbias_pre", "mul_gelu", "qgelu_bias_pre", "mul_qgelu"])
@pytest.mark.parametrize("M,N,K,batch", [(512, 512, 128, 1), (300, 520, 64, 1), (256, 1024, 256, 3), (1000, 264, 192, 2)])
def test_tma_store_epilogue_matches_direct_stores(cfg, M, N, K, batch):
    """the compile-time epilogues of the CTA-pair kernel write through shared memory + TMA stores (ragged M / N are clipped by
    the tensor map, batches are map dimensions): bit-identical to the direct-store epilogue and right against torch"""
    from feed_forward_vqgan_clip_b200 import _lib
    lib = _lib.load()
    a = _rand(batch, M, K, seed=101)
    b = _rand(N, K, seed=102)
    aux = _rand(batch, M, N, seed=103)
    bias_c, bias_r = torch.randn(N, device=DEV), torch.randn(M, device=DEV)
    kw = dict(two_cta=1, a_role=ops.ROLE_OUT, a_bs=M * K, batch=batch, out_bs=M * N)
    outs = []
    for tma in (2, 0):
        lib.ffvc_gemm_set_tma_store(tma)
        out = torch.full((batch, M, N), 7.0, device=DEV, dtype=torch.bfloat16)
        pre = torch.full((batch, M, N), 7.0, device=DEV, dtype=torch.bfloat16)
        if cfg == "gelu_bias_pre":
            ops.gemm(a, b, out, M, N, K, bias=bias_c, act=ops.ACT_GELU, pre_out=pre, **kw)
        elif cfg == "gelu_rowbias_pre":
            ops.gemm(a, b, out, M, N, K, bias=bias_r, bias_mode=2, act=ops.ACT_GELU, pre_out=pre, **kw)
        elif cfg == "mul_gelu":
            ops.gemm(a, b, out, M, N, K, aux=aux, mul_mode=ops.ACT_GELU, **kw)
        elif cfg == "qgelu_bias_pre":
            ops.gemm(a, b, out, M, N, K, bias=bias_c, act=ops.ACT_QUICKGELU, pre_out=pre, **kw)
        else:
            ops.gemm(a, b, out, M, N, K, aux=aux, mul_mode=ops.ACT_QUICKGELU, **kw)
        outs.append((out, pre))
    lib.ffvc_gemm_set_tma_store(1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    acc = torch.einsum("bmk,nk->bmn", a.float(), b.float())
    x = aux.float()
    if cfg == "gelu_bias_pre":
        _check(outs[0][1], acc + bias_c, 2e-2)
        _check(outs[0][0], F.gelu(acc + bias_c), 2e-2)
    elif cfg == "gelu_rowbias_pre":
        _check(outs[0][1], acc + bias_r[None, :, None], 2e-2)
        _check(outs[0][0], F.gelu(acc + bias_r[None, :, None]), 2e-2)
    elif cfg == "mul_gelu":
        gp = 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * 3.141592653589793) ** 0.5
        _check(outs[0][0], acc * gp, 2e-2)
    elif cfg == "qgelu_bias_pre":
        u = acc + bias_c
        _check(outs[0][1], u, 2e-2)
        _check(outs[0][0], u * torch.sigmoid(1.702 * u), 2e-2)
    else:
        sg = torch.sigmoid(1.702 * x)
        _check(outs[0][0], acc * (sg + 1.702 * x * sg * (1 - sg)), 2e-2)


@pytest.mark.parametrize("M,N,K,segs,two", [(1024, 1024, 4096, 1, 0), (4096, 1024, 2048, 1, 1), (256, 1024, 512, 8, 0), (304, 200, 1000, 1, 0)])
def test_stream_k_wgrad_matches_split_k(M, N, K, segs, two):
    """fp32-atomic (wgrad) GEMMs whose tiles do not fill the SMs evenly run stream-K: contiguous (tile, k-block) ranges per
    worker, partial tiles accumulated atomically — same result as split-K and as torch"""
    from feed_forward_vqgan_clip_b200 import _lib
    lib = _lib.load()
    # dW[m, n] += sum_seg sum_k A[seg][k][m] * B[seg][k][n]   (both operands MN-major, like linear_wgrad)
    a, b = _rand(segs, K, M, seed=111), _rand(segs, K, N, seed=112)
    ref = torch.einsum("skm,skn->mn", a.float(), b.float())
    outs = []
    for sk in (1, 0):
        lib.ffvc_gemm_set_stream_k(sk)
        out = torch.full((M, N), 0.25, device=DEV, dtype=torch.float32)
        ops.gemm(a, b, out, M, N, K, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=M, b_ld=N, a_role=ops.ROLE_SEG, a_bs=K * M,
                 b_role=ops.ROLE_SEG, b_bs=K * N, k_segs=segs, splits=4, atomic=True, two_cta=two)
        outs.append(out)
    lib.ffvc_gemm_set_stream_k(0)
    _check(outs[0] - 0.25, ref, 2e-3)
    _check(outs[1] - 0.25, ref, 2e-3)
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("epi16", [0, 2])
@pytest.mark.parametrize("n,h,w,cin,with_res", [(2, 128, 128, 128, True), (3, 4, 256, 64, False), (1, 256, 256, 128, True)])
def test_conv3x3_halo_epilogue_groupnorm_statistics(n, h, w, cin, with_res, epi16, ffvc_options):
    """ffvc_conv3x3_halo_gn: same output as ffvc_conv3x3_halo, and the (mean, rstd) its epilogue produces equal those of the
    separate statistics pass over the stored tensor (taming Normalize = GroupNorm(32, eps 1e-6))"""
    from feed_forward_vqgan_clip_b200.ops import call
    ffvc_options(halo_epi16=epi16)                       # 8 or 16 epilogue warps
    BF = torch.bfloat16
    cout = 128
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(n, h, w, cin, generator=g)).to(DEV).to(BF)
    wt = (torch.randn(cout, 9, cin, generator=g) * (9 * cin) ** -0.5).to(DEV).to(BF)
    bias = torch.randn(cout, generator=g).to(DEV)
    res = torch.randn(n * h * w, cout, generator=g).to(DEV).to(BF) if with_res else None
    out0 = torch.empty(n * h * w, cout, device=DEV, dtype=BF)
    call("conv3x3_halo", x, wt, out0, n, h, w, cin, cout, cout, bias, res, None, 0, 0, 0)
    out1 = torch.empty_like(out0)
    from feed_forward_vqgan_clip_b200 import _lib
    nws = int(_lib.load().ffvc_groupnorm_ws_doubles(n, h * w, 32))
    ws = torch.full((nws,), 123.0, device=DEV, dtype=torch.float64)             # every slot that is read is written first
    call("conv3x3_halo_gn", x, wt, out1, n, h, w, cin, cout, cout, bias, res, ws)
    assert torch.equal(out0, out1)
    mean, rstd = torch.empty(n * 32, device=DEV), torch.empty(n * 32, device=DEV)
    call("groupnorm_finalize", ws, mean, rstd, n, h * w, cout, 32, 1e-6)
    ws2 = torch.full((nws,), -5.0, device=DEV, dtype=torch.float64)
    mean2, rstd2 = torch.empty_like(mean), torch.empty_like(rstd)
    call("groupnorm_stats", out1, ws2, mean2, rstd2, n, h * w, cout, 32, 1e-6)
    assert torch.allclose(mean, mean2, atol=1e-5, rtol=1e-5), (mean - mean2).abs().max()
    assert torch.allclose(rstd, rstd2, rtol=1e-4), ((rstd - rstd2) / rstd2).abs().max()
    o = out1.float().view(n, h * w, 32, 4)
    assert torch.allclose(mean.view(n, 32), o.mean((1, 3)), atol=1e-4)


@pytest.mark.parametrize("epi16", [0, 1])
@pytest.mark.parametrize("n,h,w,cin", [(2, 128, 128, 128), (3, 4, 256, 64), (1, 256, 256, 128)])
def test_conv3x3_halo_epilogue_groupnorm_backward_statistics(n, h, w, cin, epi16, ffvc_options):
    """ffvc_conv3x3_halo_gnbwd: same dgrad output as ffvc_conv3x3_halo; the (sum g, sum g*xhat) its epilogue accumulates equal
    the first pass of ffvc_groupnorm_bwd over (dy, x), and ffvc_groupnorm_bwd_apply on them gives the same dx"""
    from feed_forward_vqgan_clip_b200.ops import call
    ffvc_options(halo_epi16=epi16)
    BF = torch.bfloat16
    cout = 128                                            # channels of dy = channels of the Normalize
    g = torch.Generator().manual_seed(5)
    dyin = torch.randn(n, h, w, cin, generator=g).to(DEV).to(BF)
    wt = (torch.randn(cout, 9, cin, generator=g) * (9 * cin) ** -0.5).to(DEV).to(BF)
    x = (0.5 + 1.5 * torch.randn(n * h * w, cout, generator=g)).to(DEV).to(BF)
    add = torch.randn(n * h * w, cout, generator=g).to(DEV).to(BF)
    gamma = (1 + 0.2 * torch.randn(cout, generator=g)).to(DEV)
    beta = (0.2 * torch.randn(cout, generator=g)).to(DEV)
    from feed_forward_vqgan_clip_b200 import _lib
    nws = int(_lib.load().ffvc_groupnorm_ws_doubles(n, h * w, 32))
    ws0 = torch.empty(nws, device=DEV, dtype=torch.float64)
    mean, rstd = torch.empty(n * 32, device=DEV), torch.empty(n * 32, device=DEV)
    call("groupnorm_stats", x, ws0, mean, rstd, n, h * w, cout, 32, 1e-6)
    dy0 = torch.empty(n * h * w, cout, device=DEV, dtype=BF)
    call("conv3x3_halo", dyin, wt, dy0, n, h, w, cin, cout, cout, None, None, None, 0, 0, 0)
    dx0 = torch.empty_like(dy0)
    ws_ref = torch.empty(nws, device=DEV, dtype=torch.float64)
    call("groupnorm_bwd", dy0, x, mean, rstd, gamma, beta, ws_ref, add, dx0, n, h * w, cout, 32, 1)
    dy1 = torch.empty_like(dy0)
    sums = torch.full((nws,), 7.0, device=DEV, dtype=torch.float64)
    call("conv3x3_halo_gnbwd", dyin, wt, dy1, n, h, w, cin, cout, cout, None, x, mean, rstd, gamma, beta, sums)
    assert torch.equal(dy0, dy1)
    dx1 = torch.empty_like(dy0)
    call("groupnorm_bwd_apply", dy1, x, mean, rstd, gamma, beta, sums, add, dx1, n, h * w, cout, 32, 1)
    ref = ws_ref[:n * 64]                                # both forms leave the folded (sum g, sum g * xhat) in ws[n][g][2]
    err = (sums[:n * 64] - ref).abs().max().item()
    assert err <= 2e-3 * ref.abs().max().item() + 1e-3, (err, ref.abs().max().item())
    # reproducible: a second run of the epilogue-statistics form gives the same bits (round 1 used double atomics)
    sums2 = torch.full((nws,), -3.0, device=DEV, dtype=torch.float64)
    dy2, dx2 = torch.empty_like(dy0), torch.empty_like(dy0)
    call("conv3x3_halo_gnbwd", dyin, wt, dy2, n, h, w, cin, cout, cout, None, x, mean, rstd, gamma, beta, sums2)
    call("groupnorm_bwd_apply", dy2, x, mean, rstd, gamma, beta, sums2, add, dx2, n, h * w, cout, 32, 1)
    assert torch.equal(sums[:n * 64], sums2[:n * 64]) and torch.equal(dx1, dx2)
    assert (dx1.float() - dx0.float()).abs().max().item() <= 2e-2 * dx0.float().abs().max().item()


@pytest.mark.parametrize("n,h,w,cin,cout,with_res,with_stats", [(2, 128, 128, 128, 128, True, True), (3, 4, 256, 64, 128, False, True),
                                                               (1, 256, 256, 128, 128, True, False), (2, 6, 128, 256, 128, False, True),
                                                               (5, 2, 128, 128, 64, False, False)])
def test_conv3x3_halo_with_fused_groupnorm_apply(n, h, w, cin, cout, with_res, with_stats):
    """ffvc_conv3x3_halo_xf — the conv of swish(GroupNorm(x)) with the apply done by transform warps on the halo tile in shared
    memory — against ffvc_groupnorm_apply followed by the plain halo conv (taming Normalize + nonlinearity + Conv2d): same output up
    to the bf16 rounding of the normalised tensor (the fused form computes swish as h + h tanh(h)), borders (zero padding of the
    ACTIVATED tensor) included, and the same epilogue statistics of the output; twice the same bits."""
    from feed_forward_vqgan_clip_b200 import _lib
    from feed_forward_vqgan_clip_b200.ops import call
    BF = torch.bfloat16
    g = torch.Generator().manual_seed(11)
    x = (0.3 + 1.7 * torch.randn(n, h, w, cin, generator=g)).to(DEV).to(BF)
    wt = (torch.randn(cout, 9, cin, generator=g) * (9 * cin) ** -0.5).to(DEV).to(BF)
    bias = torch.randn(cout, generator=g).to(DEV)
    gamma = (1 + 0.2 * torch.randn(cin, generator=g)).to(DEV)
    beta = (0.5 * torch.randn(cin, generator=g)).to(DEV)            # a large shift: a wrong padding value would show at the borders
    res = torch.randn(n * h * w, cout, generator=g).to(DEV).to(BF) if with_res else None
    nws = int(_lib.load().ffvc_groupnorm_ws_doubles(n, h * w, 32))
    ws = torch.empty(nws, device=DEV, dtype=torch.float64)
    mean, rstd = torch.empty(n * 32, device=DEV), torch.empty(n * 32, device=DEV)
    call("groupnorm_stats", x, ws, mean, rstd, n, h * w, cin, 32, 1e-6)
    a = torch.empty_like(x)
    call("groupnorm_apply", x, mean, rstd, gamma, beta, a, n, h * w, cin, 32, 1)
    out0 = torch.empty(n * h * w, cout, device=DEV, dtype=BF)
    call("conv3x3_halo", a, wt, out0, n, h, w, cin, cout, cout, bias, res, None, 0, 0, 0)
    outs = []
    for _ in range(2):
        out1 = torch.full((n * h * w, cout), float("nan"), device=DEV, dtype=BF)
        ws1 = torch.full((nws,), float("nan"), device=DEV, dtype=torch.float64) if with_stats else None
        call("conv3x3_halo_xf", x, wt, out1, n, h, w, cin, cout, cout, bias, res, mean, rstd, gamma, beta, 32, ws1)
        outs.append(out1)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    err = (outs[0].float() - out0.float()).abs().max().item()
    assert err <= 2e-2 * out0.float().abs().max().item(), err
    o4 = outs[0].float().view(n, h, w, cout)
    r4 = out0.float().view(n, h, w, cout)
    for sl in ((slice(None), 0), (slice(None), h - 1), (slice(None), slice(None), 0), (slice(None), slice(None), w - 1)):
        assert (o4[sl] - r4[sl]).abs().max().item() <= 2e-2 * r4.abs().max().item()          # image borders
    if with_stats:
        m1, r1 = torch.empty(n * 32, device=DEV), torch.empty(n * 32, device=DEV)
        call("groupnorm_finalize", ws1, m1, r1, n, h * w, cout, 32, 1e-6)
        m2, r2 = torch.empty_like(m1), torch.empty_like(r1)
        call("groupnorm_stats", outs[1], ws, m2, r2, n, h * w, cout, 32, 1e-6)
        assert torch.allclose(m1, m2, atol=1e-5, rtol=1e-5) and torch.allclose(r1, r2, rtol=1e-4)


@pytest.mark.parametrize("M,N,K,kind", [(1024, 512, 256, "kk"), (2048, 1024, 1024, "kk"), (2048, 768, 640, "km"), (1024, 512, 4096, "mm"),
                                        (1536, 256, 320, "kk")])
def test_gemm_cluster_of_four_multicast_matches_the_pair_kernel(M, N, K, kind, ffvc_options):
    """option gemm_quad: clusters of 4 CTAs, two M-adjacent pair tiles share their B tile through TMA multicast (each CTA fetches half
    of its B rows, UTMALDG.MULTICAST.2CTA) — bit-identical to the pair kernel for K-major / MN-major operands, fused epilogues and the
    fp32-atomic wgrad form (same K order per element); shapes whose tile rows are odd (1536 / 256 = 6 is even, 1280 would not be) or
    narrower than 256 columns fall back to pairs"""
    from feed_forward_vqgan_clip_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    BFD = torch.bfloat16

    def run(out):
        if kind == "mm":
            ops.gemm(a, b, out, M, N, K, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=M, b_ld=N, atomic=True, two_cta=1)
        elif kind == "km":
            ops.gemm(a, b, out, M, N, K, b_mode=ops.MNMAJOR, b_ld=N, res=res, two_cta=1)
        else:
            ops.gemm(a, b, out, M, N, K, bias=bias, act=ops.ACT_GELU, pre_out=pre, two_cta=1)

    if kind == "mm":
        a, b = torch.randn(K, M, generator=g).to(DEV).to(BFD), torch.randn(K, N, generator=g).to(DEV).to(BFD)
        outs = [torch.zeros(M, N, device=DEV) for _ in range(2)]
    elif kind == "km":
        a, b = torch.randn(M, K, generator=g).to(DEV).to(BFD), torch.randn(K, N, generator=g).to(DEV).to(BFD)
        res = torch.randn(M, N, generator=g).to(DEV).to(BFD)
        outs = [torch.empty(M, N, device=DEV, dtype=BFD) for _ in range(2)]
    else:
        a, b = torch.randn(M, K, generator=g).to(DEV).to(BFD), torch.randn(N, K, generator=g).to(DEV).to(BFD)
        bias = torch.randn(N, generator=g).to(DEV)
        pre = torch.empty(M, N, device=DEV, dtype=BFD)
        outs = [torch.empty(M, N, device=DEV, dtype=BFD) for _ in range(2)]
    ffvc_options(gemm_quad=0)
    run(outs[0])
    pre0 = pre.clone() if kind == "kk" else None
    ffvc_options(gemm_quad=1)
    run(outs[1])
    torch.cuda.synchronize()
    if kind == "mm":                       # fp32 atomics: the order of the k-block partial sums may differ
        assert torch.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-3)
        ref = a.float().t() @ b.float()
        assert torch.allclose(outs[1], ref, rtol=2e-2, atol=2e-1)
    else:
        assert torch.equal(outs[0], outs[1])
        if kind == "kk":
            assert torch.equal(pre0, pre)
