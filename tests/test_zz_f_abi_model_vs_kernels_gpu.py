"""tests/abi_model.py (the CPU statement of the C ABI the host-logic tests run against) checked against the kernels themselves
for the small entry points that the other GPU tests only reach through the engines: the same random buffers go through the
kernel on the GPU and through the model on the CPU.  Copies / gathers / casts must agree bit for bit, arithmetic within bf16 /
fp32 rounding.  Written after round 1's last GPU session (sorts last)."""
import ctypes as C

import pytest
import torch

import abi_model
from feed_forward_vqgan_clip_b200.ops import call

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


def _both(name, args, outs, exact=True, tol=1e-2):
    """run ffvc_<name> on CUDA copies and k_<name> on CPU copies of `args`; compare the tensors at positions `outs`"""
    gpu = [a.clone().cuda() if torch.is_tensor(a) else a for a in args]
    cpu = [a.clone() if torch.is_tensor(a) else a for a in args]
    call(name, *gpu)
    getattr(abi_model, "k_" + name)(*cpu)
    torch.cuda.synchronize()
    for i in outs:
        a, b = gpu[i].float().cpu(), cpu[i].float()
        if exact:
            assert torch.equal(a, b), (name, i, float((a - b).abs().max()))
        else:
            assert torch.allclose(a, b, atol=tol, rtol=tol), (name, i, float((a - b).abs().max()))


def test_copies_gathers_and_casts_match_the_model_bit_for_bit():
    g = torch.Generator().manual_seed(0)
    r = lambda *s, dt=BF: torch.randn(*s, generator=g).to(dt)                                   # noqa: E731
    N, T, W = 3, 5, 64
    _both("clip_assemble", [r(N * (T - 1), W), r(W, dt=F32), r(T, W, dt=F32), torch.zeros(N * T, W, dtype=BF), N, T, W], [3],
          exact=False, tol=1e-2)
    _both("copy_rows", [r(N * T, W), torch.zeros(N, W, dtype=BF), N, W, T * W, W], [1])
    tok = torch.randint(0, 50, (N, T), generator=g)
    _both("embed_tokens", [tok, r(50, W, dt=F32), r(T, W, dt=F32), torch.zeros(N * T, W, dtype=BF), N * T, T, W], [3], exact=False)
    _both("gather_rows", [r(N * T, W), torch.randint(0, T, (N,), generator=g), torch.zeros(N, W, dtype=BF), N, T, W], [2])
    _both("broadcast_rows", [r(T * W, dt=F32), torch.zeros(N, T * W, dtype=BF), N, T * W], [1])
    _both("cast_f32_bf16", [r(1000, dt=F32), torch.zeros(1000, dtype=BF), 1000], [1])
    _both("cast_bf16_f32", [r(1000), torch.zeros(1000, dtype=F32), 1000], [1])
    _both("cast_f32_bf16_pitched", [r(7, 13, dt=F32), torch.full((7, 16), 5.0, dtype=BF), 7, 13, 16], [1])
    _both("transpose", [r(2, 6, 40), torch.zeros(2, 40, 6, dtype=BF), 2, 6, 40, 0, 0], [1])
    _both("upsample2x_fwd", [r(2, 4, 4, 16), torch.zeros(2, 8, 8, 16, dtype=BF), 2, 4, 4, 16], [1])
    _both("relu_mask", [r(999), r(999), torch.zeros(999, dtype=BF), 999], [2])
    _both("im2col3x3_cin3", [r(2, 6, 6, 3, dt=F32), torch.full((2 * 36, 32), 3.0, dtype=BF), 2, 6, 6], [1])


def test_small_arithmetic_entry_points_match_the_model():
    g = torch.Generator().manual_seed(1)
    r = lambda *s, dt=BF: torch.randn(*s, generator=g).to(dt)                                   # noqa: E731
    _both("add_bf16", [r(1001), r(1001), torch.zeros(1001, dtype=BF), 1001], [2], exact=False)
    _both("upsample2x_bwd", [r(2, 8, 8, 16), torch.zeros(2, 4, 4, 16, dtype=BF), 2, 4, 4, 16], [1], exact=False, tol=2e-2)
    _both("axpy_f32", [r(1003, dt=F32), r(1003, dt=F32), 0.37, 1003], [1], exact=False, tol=1e-6)
    _both("sumsq", [r(100003, dt=F32), torch.zeros(1, dtype=F32), 100003], [1], exact=False, tol=1e-4)
    _both("rownorm2", [r(37, 64, dt=F32), torch.zeros(37, dtype=F32), 37, 64], [1], exact=False, tol=1e-5)
    _both("colsum", [r(300, 72), torch.ones(72, dtype=F32), 300, 72], [1], exact=False, tol=1e-3)
    _both("rowsum", [r(3, 20, 64), torch.ones(20, dtype=F32), 3, 20, 64], [1], exact=False, tol=1e-3)
    _both("image_post_fwd", [r(999, dt=F32) * 2, torch.zeros(999, dtype=F32), 999], [1], exact=False, tol=1e-6)
    _both("image_post_bwd", [r(999, dt=F32), r(999, dt=F32) * 2, torch.zeros(999, dtype=F32), 999], [2], exact=False, tol=1e-6)
    _both("clamp_bwd", [r(999, dt=F32), r(999, dt=F32) * 2, torch.zeros(999, dtype=F32), 999, -1.0, 1.5], [2], exact=False, tol=1e-6)
    mean, std = (C.c_float * 3)(0.48, 0.45, 0.40), (C.c_float * 3)(0.26, 0.26, 0.27)
    _both("normalize3_fwd", [r(300, dt=F32), torch.zeros(300, dtype=F32), 300, C.addressof(mean), C.addressof(std)], [1], exact=False,
          tol=1e-5)
    _both("normalize3_bwd", [r(300, dt=F32), torch.ones(300, dtype=F32), 300, C.addressof(std)], [1], exact=False, tol=1e-5)
    _both("softmax_fwd", [r(40, 24, dt=F32) * 3, torch.ones(40, 24, dtype=BF), 40, 20, 24], [1], exact=False, tol=1e-2)
    _both("softmax_causal_fwd", [r(2 * 5, 8, dt=F32) * 3, torch.ones(2 * 5, 8, dtype=BF), 10, 5, 8], [1], exact=False, tol=1e-2)
    hyper = torch.tensor([1e-3, .9, .999, 1e-8, 1, 1, 0.5, 0, 4, 2.0, 9.0, 1, 1e-3, 100.0, 0, 0.99])
    _both("adam_tick", [hyper], [0], exact=False, tol=1e-6)
    n = 1000
    p, gr, m, v = r(n, dt=F32), r(n, dt=F32), r(n, dt=F32) * 0.1, r(n, dt=F32).abs() * 0.01
    hy = torch.tensor([1e-3, .9, .999, 1e-8, 1 - .9 ** 5, (1 - .999 ** 5) ** .5, 0.5, 0, 5, 0, 0, 1, 1e-3, 0, 0, 0.99])
    _both("adam_step_ema", [p, gr, m, v, torch.zeros(n, dtype=BF), p * 0.9, n, hy], [0, 2, 3, 5], exact=False, tol=1e-5)
    _both("adam_step_ema", [p, gr, m, v, torch.zeros(n, dtype=BF), p * 0.9, n, hy], [4], exact=False, tol=1e-2)       # the bf16 shadow
