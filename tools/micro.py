#!/usr/bin/env python
"""Micro-benchmark of single launches at the bench workload's shapes (config #2, 64 prompts), CUDA events over rotating buffer
sets larger than L2.  `python tools/micro.py [case ...]` prints one JSON line per case: microseconds, TFLOP/s or GB/s.
Under ncu: `ncu --set full -k regex:<kernel> -c 2 python tools/micro.py <case> --reps 1`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from feed_forward_vqgan_clip_b200 import _lib, ops  # noqa: E402
from feed_forward_vqgan_clip_b200.ops import call  # noqa: E402

DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32
REPS = 20


def timeit(fn, nsets):
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(REPS):
        fn(i % nsets)
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / REPS


def rnd(*shape, dtype=BF, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*shape, device=DEV, generator=g) * scale).to(dtype)


def case_halo(kind, n=64, h=256, w=256, cin=128, cout=128):
    nsets = 3
    xs = [rnd(n, h, w, cin, seed=i) for i in range(nsets)]
    wt = rnd(cout, 9, cin, scale=(9 * cin) ** -0.5, seed=9)
    bias = rnd(cout, dtype=F32, seed=8)
    out_f32 = cout < 8
    outs = [torch.empty(n * h * w, cout, device=DEV, dtype=F32 if out_f32 else BF) for _ in range(nsets)]
    res = rnd(n * h * w, cout, seed=7) if "res" in kind else None
    nws = int(_lib.load().ffvc_groupnorm_ws_doubles(n, h * w, 32))
    ws = torch.empty(nws, device=DEV, dtype=torch.float64)
    mean, rstd = torch.zeros(n * 32, device=DEV), torch.ones(n * 32, device=DEV)
    gamma, beta = torch.ones(128, device=DEV), torch.zeros(128, device=DEV)
    gx = rnd(n * h * w, 128, seed=5) if "gnbwd" in kind else None

    def fn(i):
        if kind.startswith("xf"):
            call("conv3x3_halo_xf", xs[i], wt, outs[i], n, h, w, cin, cout, cout, bias, res, mean, rstd, gamma, beta, 32,
                 ws if "gn" in kind else None)
        elif kind.startswith("gnbwd"):
            call("conv3x3_halo_gnbwd", xs[i], wt, outs[i], n, h, w, cin, cout, cout, res, gx, mean, rstd, gamma, beta, ws)
        elif kind.startswith("gn"):
            call("conv3x3_halo_gn", xs[i], wt, outs[i], n, h, w, cin, cout, cout, bias, res, ws)
        else:
            call("conv3x3_halo", xs[i], wt, outs[i], n, h, w, cin, cout, cout, bias, res, None, 0, 0, int(out_f32))
    us = timeit(fn, nsets)
    fl = 2.0 * n * h * w * cout * 9 * cin
    return dict(us=us, tflops=fl / us / 1e6)


def case_gemm(kind):
    nsets = 3
    if kind in ("mul1", "act1pre", "plain4096"):
        M, N, K = 16384, 4096, 1024
    else:
        M, N, K = 16384, 1024, 4096
    a = [rnd(M, K, seed=i) for i in range(nsets)]
    wK = rnd(N, K, scale=K ** -0.5, seed=11)            # forward layout [N][K]
    wM = rnd(K, N, scale=K ** -0.5, seed=12)            # dgrad reads the forward layout of the Linear whose output has K columns
    outs = [torch.empty(M, N, device=DEV, dtype=BF) for _ in range(nsets)]
    aux = [rnd(M, N, seed=20 + i) for i in range(nsets)]
    bias = rnd(N, dtype=F32, seed=13)

    def fn(i):
        if kind == "mul1":
            ops.gemm(a[i], wM, outs[i], M, N, K, b_mode=ops.MNMAJOR, b_ld=N, aux=aux[i], mul_mode=ops.ACT_GELU)
        elif kind == "act1pre":
            ops.gemm(a[i], wK, outs[i], M, N, K, bias=bias, act=ops.ACT_GELU, pre_out=aux[i])
        elif kind in ("plain4096", "plain1024"):
            ops.gemm(a[i], wK, outs[i], M, N, K, bias=bias)
        elif kind == "res":
            ops.gemm(a[i], wK, outs[i], M, N, K, bias=bias, res=aux[i])
        elif kind == "dgrad1024":
            ops.gemm(a[i], wM, outs[i], M, N, K, b_mode=ops.MNMAJOR, b_ld=N)
    us = timeit(fn, nsets)
    return dict(us=us, tflops=2.0 * M * N * K / us / 1e6)


def case_tok(kind, B=64, T=256, D=1024):
    nsets = 3
    W1 = rnd(4 * T, T, scale=T ** -0.5, seed=1)          # fc1 weight [4T][T]
    W2 = rnd(T, 4 * T, scale=(4 * T) ** -0.5, seed=2)    # fc2 weight [T][4T]
    n1 = [rnd(B, T, D, seed=i) for i in range(nsets)]
    big = [rnd(B, 4 * T, D, seed=10 + i) for i in range(nsets)]
    big2 = [torch.empty(B, 4 * T, D, device=DEV, dtype=BF) for _ in range(nsets)]
    small = [torch.empty(B, T, D, device=DEV, dtype=BF) for _ in range(nsets)]
    b1, b2 = rnd(4 * T, dtype=F32, seed=3), rnd(T, dtype=F32, seed=4)

    def fn(i):
        if kind == "fc1":
            ops.gemm(W1, n1[i], big2[i], 4 * T, D, T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT, b_bs=T * D, batch=B, out_bs=4 * T * D,
                     bias=b1, bias_mode=2, act=ops.ACT_GELU, pre_out=big[i])
        elif kind == "fc2":
            ops.gemm(W2, big[i], small[i], T, D, 4 * T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT, b_bs=4 * T * D, batch=B, out_bs=T * D,
                     bias=b2, bias_mode=2, res=n1[i])
        elif kind == "dgrad_mul1":
            ops.gemm(W2, n1[i], big2[i], 4 * T, D, T, a_mode=ops.MNMAJOR, a_ld=4 * T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT,
                     b_bs=T * D, batch=B, out_bs=4 * T * D, aux=big[i], mul_mode=ops.ACT_GELU)
        elif kind == "dgrad2":
            ops.gemm(W1, big[i], small[i], T, D, 4 * T, a_mode=ops.MNMAJOR, a_ld=T, b_mode=ops.MNMAJOR, b_ld=D, b_role=ops.ROLE_OUT,
                     b_bs=4 * T * D, batch=B, out_bs=T * D)
    us = timeit(fn, nsets)
    return dict(us=us, tflops=2.0 * B * 4 * T * T * D / us / 1e6)


CASES = {
    "conv_out_fwd": lambda: case_halo("plain", cout=3),
    "halo_plain": lambda: case_halo("plain"),
    "halo_gn": lambda: case_halo("gn"),
    "halo_gn_res": lambda: case_halo("gn_res"),
    "halo_gnbwd": lambda: case_halo("gnbwd"),
    "halo_plain_res": lambda: case_halo("plain_res"),
    "halo_xf": lambda: case_halo("xf"),
    "halo_xf_gn": lambda: case_halo("xf_gn"),
    "halo_xf_gn_res": lambda: case_halo("xf_gn_res"),
    "gemm_mul1": lambda: case_gemm("mul1"),
    "gemm_act1pre": lambda: case_gemm("act1pre"),
    "gemm_plain4096": lambda: case_gemm("plain4096"),
    "gemm_plain1024": lambda: case_gemm("plain1024"),
    "gemm_res": lambda: case_gemm("res"),
    "gemm_dgrad1024": lambda: case_gemm("dgrad1024"),
    "tok_fc1": lambda: case_tok("fc1"),
    "tok_fc2": lambda: case_tok("fc2"),
    "tok_dgrad_mul1": lambda: case_tok("dgrad_mul1"),
    "tok_dgrad2": lambda: case_tok("dgrad2"),
}


def main():
    global REPS
    args = sys.argv[1:]
    if "--reps" in args:
        i = args.index("--reps")
        REPS = int(args[i + 1])
        del args[i:i + 2]
    torch.cuda.set_device(0)
    for name in (args or list(CASES)):
        r = CASES[name]()
        print(json.dumps(dict(case=name, **{k: round(v, 2) for k, v in r.items()})), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
