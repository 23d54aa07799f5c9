"""Thin Python wrappers over the C ABI (include/ffvc.h).  Tensors are torch CUDA tensors used purely as
device-memory handles; every computation happens in libffvc_sm100.so."""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmParams, check

KMAJOR, MNMAJOR, CONV3X3 = 0, 1, 2
ROLE_BCAST, ROLE_OUT, ROLE_SEG = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_QUICKGELU, ACT_SWISH = 0, 1, 2, 3


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "ffvc ops need CUDA tensors (no CPU fallback)"
    return C.c_void_p(t.data_ptr())


def gemm(a, b, out, M, N, K, *, a_mode=KMAJOR, b_mode=KMAJOR, a_ld=None, b_ld=None,
         a_role=ROLE_BCAST, b_role=ROLE_BCAST, a_bs=0, b_bs=0, batch=1, k_segs=1, splits=1, block_n=0,
         conv=None, pre_out=None, aux=None, res=None, bias=None, ldc=None, out_bs=0, atomic=False,
         bias_mode=1, act=ACT_NONE, mul_mode=ACT_NONE, alpha=1.0):
    """out[b,m,n] (+)= epilogue(alpha * sum_k A[m,k] B[n,k]); see include/ffvc.h:ffvc_gemm."""
    p = GemmParams()
    p.a, p.b = _ptr(a), _ptr(b)
    p.a_mode, p.b_mode = a_mode, b_mode
    if a_ld is None:
        a_ld = K if a_mode == KMAJOR else M
    if b_ld is None:
        b_ld = K if b_mode == KMAJOR else N
    p.a_ld, p.b_ld = a_ld, b_ld
    p.a_batch_role, p.b_batch_role = a_role, b_role
    p.a_batch_stride, p.b_batch_stride = a_bs, b_bs
    p.M, p.N, p.K = M, N, K
    p.batch, p.k_segs, p.splits, p.block_n = batch, k_segs, splits, block_n
    if conv is not None:
        p.conv_n, p.conv_h, p.conv_w, p.conv_c = conv
    p.out, p.pre_out, p.aux, p.res, p.bias = _ptr(out), _ptr(pre_out), _ptr(aux), _ptr(res), _ptr(bias)
    p.ldc = N if ldc is None else ldc
    p.out_batch_stride = out_bs
    p.out_fp32 = 1 if out.dtype == torch.float32 else 0
    assert out.dtype in (torch.float32, torch.bfloat16)
    p.atomic = 1 if atomic else 0
    p.bias_mode, p.act, p.mul_mode, p.alpha = bias_mode, act, mul_mode, alpha
    check(_lib.load().ffvc_gemm(C.byref(p), _stream()))
    return out
