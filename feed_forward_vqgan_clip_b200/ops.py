"""Thin Python wrappers over the C ABI (include/ffvc.h).  Tensors are torch CUDA tensors used purely as
device-memory handles; every computation happens in libffvc_sm100.so.  No fallback path exists."""
import ctypes as C
import json
import os

import torch

from . import _lib
from ._lib import GemmParams, check

KMAJOR, MNMAJOR, CONV3X3 = 0, 1, 2
ROLE_BCAST, ROLE_OUT, ROLE_SEG = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_QUICKGELU, ACT_SWISH, ACT_RELU = 0, 1, 2, 3, 4

BF16 = torch.bfloat16
F32 = torch.float32


def require_cuda(dev, what):
    """every engine calls this on construction: the product path exists on CUDA only"""
    if dev.type != "cuda":
        raise RuntimeError("%s runs on CUDA only (no CPU fallback): move the module to a B200 first" % what)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _arg(a):
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise RuntimeError("ffvc ops need CUDA tensors (there is no CPU fallback)")
        return a.data_ptr()
    return a


def call(name, *args):
    """Call `int ffvc_<name>(..., void* stream)`; tensors become device pointers, the stream is appended."""
    fn = getattr(_lib.load(), "ffvc_" + name)
    check(fn(*[_arg(a) for a in args], _stream()))


_NUM_SMS = None


def num_sms():
    global _NUM_SMS
    if _NUM_SMS is None:
        _NUM_SMS = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    return _NUM_SMS


# ------------------------------------------------------------------ side streams: independent launches next to the critical path
# The step's critical path is a chain of tensor-core GEMMs; several HBM-bound launches hang off it without feeding the next GEMM
# (bias-gradient column / row sums, the optimizer update of a finished bucket, zeroing the gradient arena, drawing the cutout
# noise).  fork() enqueues such work on a side stream ordered after everything issued so far on the current stream; join() makes
# the current stream wait for it.  The caller joins BEFORE freeing any tensor the forked work reads (so the caching allocator
# never hands its memory out again while the side stream is still reading) and before the step ends (so a CUDA-graph capture
# closes with one stream).
# Measured on the B200 (profiles/r02_ab_side_streams.md): NEUTRAL — the persistent tcgen05 kernels fill every SM's registers and
# shared memory, so forked launches cannot co-reside with them, and the GPU runs at its power cap, where step time follows the
# energy spent rather than the critical path.  Hence off by default (FFVC_SIDE_STREAMS=1 enables it).
SIDE_STREAMS = os.environ.get("FFVC_SIDE_STREAMS", "0") == "1"
_side = {}
_pending = set()


def fork(fn, key="aux"):
    if not (SIDE_STREAMS and torch.cuda.is_available()):
        return fn()
    dev = torch.cuda.current_device()
    s = _side.get((dev, key))
    if s is None:
        s = _side[(dev, key)] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    ev = torch.cuda.Event()
    ev.record(main)
    s.wait_event(ev)
    with torch.cuda.stream(s):
        fn()
    _pending.add((dev, key))


def join(key="aux"):
    if not _pending:
        return
    dev = torch.cuda.current_device()
    if (dev, key) in _pending:
        torch.cuda.current_stream().wait_stream(_side[(dev, key)])
        _pending.discard((dev, key))


def launch_count():
    return int(_lib.load().ffvc_launch_count())


def reset_launch_count():
    _lib.load().ffvc_reset_launch_count()


# ------------------------------------------------------------------ per-shape launch configurations measured on B200
# tools/tune_gemm.py times every legal (two_cta, tile_m, block_n, epi_warps, splits) combination of every distinct GEMM of the
# bench step on the GPU box and writes the winners to gemm_tuned.json; a call that leaves those knobs at their defaults picks
# its entry up here.  Shapes without an entry use the heuristics in ffvc_gemm.  FFVC_GEMM_TUNED=0 disables the table.
_TUNED = None
TUNED_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_tuned.json")


def tuned_table():
    global _TUNED
    if _TUNED is None:
        _TUNED = {}
        if os.environ.get("FFVC_GEMM_TUNED", "1") != "0" and os.path.exists(TUNED_PATH):
            with open(TUNED_PATH) as f:
                _TUNED = json.load(f).get("table", {})
    return _TUNED


def gemm_key(M, N, K, out_fp32, kw):
    """identity of a GEMM launch as far as the choice of tile configuration is concerned"""
    g = kw.get
    conv = g("conv")
    return ",".join(str(v) for v in (
        M, N, K, g("a_mode", 0), g("b_mode", 0), g("a_role", 0), g("b_role", 0), g("batch", 1), g("batch_inner", 1), g("k_segs", 1),
        int(out_fp32), int(bool(g("atomic", False))), 0 if g("bias") is None else g("bias_mode", 1), g("act", 0), g("mul_mode", 0),
        int(g("pre_out") is not None), int(g("res") is not None), int(g("aux") is not None), int(g("argmin_out") is not None),
        "x".join(str(c) for c in conv) if conv is not None else "-"))


def gemm(a, b, out, M, N, K, **kw):
    """ffvc_gemm with the measured per-shape configuration (if any) filled in for knobs the caller left at their defaults"""
    if not (kw.get("block_n") or kw.get("tile_m") or kw.get("two_cta") or kw.get("epi_warps")):
        t = tuned_table()
        if t:
            cfg = t.get(gemm_key(M, N, K, out is not None and out.dtype == F32, kw))
            if cfg:
                kw = dict(kw)
                for k_ in ("block_n", "tile_m", "two_cta", "epi_warps"):
                    if cfg.get(k_):
                        kw[k_] = cfg[k_]
                if cfg.get("splits") and kw.get("atomic"):
                    kw["splits"] = cfg["splits"]
    return gemm_raw(a, b, out, M, N, K, **kw)


def gemm_raw(a, b, out, M, N, K, *, a_mode=KMAJOR, b_mode=KMAJOR, a_ld=None, b_ld=None,
         a_role=ROLE_BCAST, b_role=ROLE_BCAST, a_bs=0, b_bs=0, batch=1, k_segs=1, splits=1, block_n=0,
         conv=None, pre_out=None, aux=None, res=None, bias=None, ldc=None, out_bs=0, atomic=False,
         bias_mode=1, act=ACT_NONE, mul_mode=ACT_NONE, alpha=1.0, a_off=0, b_off=0, out_off=0, batch_inner=1,
         a_bs_in=0, b_bs_in=0, out_bs_in=0, tile_m=0, two_cta=0, epi_warps=0, argmin_out=None):
    """out[b,m,n] (+)= epilogue(alpha * sum_k A[m,k] B[n,k]); see include/ffvc.h:ffvc_gemm.
    a_off / out_off: element offsets added to the base pointers."""
    p = GemmParams()
    p.a = a.data_ptr() + a_off * a.element_size()
    p.b = b.data_ptr() + b_off * b.element_size()
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
    p.batch_inner, p.tile_m, p.two_cta, p.epi_warps = batch_inner, tile_m, two_cta, epi_warps
    p.a_batch_stride_inner, p.b_batch_stride_inner, p.out_batch_stride_inner = a_bs_in, b_bs_in, out_bs_in
    if conv is not None:
        p.conv_n, p.conv_h, p.conv_w, p.conv_c = conv
    p.out = None if out is None else out.data_ptr() + out_off * out.element_size()
    p.argmin_out = None if argmin_out is None else argmin_out.data_ptr()
    p.pre_out = None if pre_out is None else pre_out.data_ptr()
    p.aux = None if aux is None else aux.data_ptr()
    p.res = None if res is None else res.data_ptr()
    p.bias = None if bias is None else bias.data_ptr()
    p.ldc = N if ldc is None else ldc
    p.out_batch_stride = out_bs
    assert out is None or out.dtype in (F32, BF16)
    p.out_fp32 = 1 if (out is not None and out.dtype == F32) else 0
    p.atomic = 1 if atomic else 0
    p.bias_mode, p.act, p.mul_mode, p.alpha = bias_mode, act, mul_mode, alpha
    check(_lib.load().ffvc_gemm(C.byref(p), _stream()))
    return out


# ------------------------------------------------------------------ convenience forms used by the model code
def linear_fwd(x, w, bias, out, M, N, K, **kw):
    """out[M,N] = x[M,K] @ w[N,K]^T + bias"""
    return gemm(x, w, out, M, N, K, bias=bias, **kw)


def linear_dgrad(dy, w, dx, M, N, K, **kw):
    """dx[M,K] = dy[M,N] @ w[N,K]  (w read MN-major straight from its forward layout)"""
    return gemm(dy, w, dx, M, K, N, b_mode=MNMAJOR, b_ld=K, **kw)


def linear_wgrad(dy, x, dw, M, N, K, splits=1):
    """dw[N,K] += dy[M,N]^T @ x[M,K]   (fp32 atomic accumulation)"""
    return gemm(dy, x, dw, N, K, M, a_mode=MNMAJOR, b_mode=MNMAJOR, a_ld=N, b_ld=K, atomic=True, splits=splits)


def auto_splits(M, N, K, sms=148):
    """split-K factor so that a wgrad GEMM with few output tiles still fills the SMs."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    kb = (K + 63) // 64
    s = 1
    while tiles * s * 2 <= sms and s * 2 <= kb:
        s *= 2
    return s
