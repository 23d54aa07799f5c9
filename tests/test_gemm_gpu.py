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
