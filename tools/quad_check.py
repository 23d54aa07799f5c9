#!/usr/bin/env python
"""quick check of the cluster-of-4 multicast GEMM (option gemm_quad) against the pair kernel: same bits, time"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from feed_forward_vqgan_clip_b200 import _lib, ops
lib = _lib.load()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
def rnd(*s): return torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / reps
for (M, N, K, kind) in [(1024, 512, 256, "kk"), (16384, 4096, 1024, "kk"), (16384, 1024, 4096, "kk"), (16384, 4096, 1024, "km"), (4096, 1024, 16384, "mm")]:
    if kind == "mm":
        a, b = rnd(K, M), rnd(K, N)
        out = [torch.zeros(M, N, device=dev) for _ in range(2)]
        run = lambda o: ops.gemm(a, b, o, M, N, K, a_mode=ops.MNMAJOR, b_mode=ops.MNMAJOR, a_ld=M, b_ld=N, atomic=True)
    elif kind == "km":
        a, b = rnd(M, K), rnd(K, N)
        out = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        run = lambda o: ops.gemm(a, b, o, M, N, K, b_mode=ops.MNMAJOR, b_ld=N)
    else:
        a, b = rnd(M, K), rnd(N, K)
        bias = torch.randn(N, device=dev, generator=g)
        out = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2)]
        run = lambda o: ops.gemm(a, b, o, M, N, K, bias=bias)
    lib.ffvc_set_option(b"gemm_quad", 0)
    run(out[0]); t0 = t(lambda: run(out[0])) if kind != "mm" else None
    lib.ffvc_set_option(b"gemm_quad", 1)
    if kind == "mm": out[0].zero_(); lib.ffvc_set_option(b"gemm_quad", 0); run(out[0]); lib.ffvc_set_option(b"gemm_quad", 1)
    run(out[1]); torch.cuda.synchronize()
    same = torch.equal(out[0], out[1]) if kind != "mm" else bool(torch.allclose(out[0], out[1], rtol=1e-4, atol=1e-3))
    t1 = t(lambda: run(out[1])) if kind != "mm" else None
    print(M, N, K, kind, "identical" if same else "DIFFERENT max %.4g" % float((out[0].float() - out[1].float()).abs().max()), "pair us", t0, "quad us", t1, flush=True)
print("max co-resident clusters of 4:", lib.ffvc_gemm_max_quads())
