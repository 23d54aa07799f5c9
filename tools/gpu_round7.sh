#!/bin/bash
mkdir -p gpurun_out
for f in test_gemm_gpu test_models_gpu; do
  timeout -k 10 300 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/$f.full 2>&1; rc=$?
  cut -c1-600 gpurun_out/$f.full | tail -100 > gpurun_out/$f.log; rm -f gpurun_out/$f.full
  echo "== $f (rc=$rc): $(tail -1 gpurun_out/$f.log)"
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/$f.log | head -10; fi
done
timeout -k 10 120 python - <<'PY' 2>&1 | tail -4
import torch, sys
sys.path.insert(0, ".")
from feed_forward_vqgan_clip_b200.ops import call
DEV, BF = "cuda", torch.bfloat16
n, h, w, c = 64, 256, 256, 128
x = torch.randn(n, h, w, c, device=DEV).to(BF)
wt = (torch.randn(c, 9, c, device=DEV) * (9 * c) ** -0.5).to(BF)
res = torch.randn(n * h * w, c, device=DEV).to(BF)
out = torch.empty(n * h * w, c, device=DEV, dtype=BF)
ws = torch.empty(n * 65, device=DEV, dtype=torch.float64)
mean, rstd = torch.empty(n * 32, device=DEV), torch.empty(n * 32, device=DEV)
def t(fn, reps=10):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / reps
gam, bet = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
call("groupnorm_stats", res, ws, mean, rstd, n, h * w, c, 32, 1e-6)
sums = torch.empty(n * 64, device=DEV, dtype=torch.float64)
a = t(lambda: call("conv3x3_halo", x, wt, out, n, h, w, c, c, c, None, None, None, 0, 0, 0))
b = t(lambda: call("conv3x3_halo_gnbwd", x, wt, out, n, h, w, c, c, c, None, res, mean, rstd, gam, bet, sums))
from feed_forward_vqgan_clip_b200 import _lib
_lib.load().ffvc_set_option(b"halo_epi16", 2)
b16 = t(lambda: call("conv3x3_halo_gnbwd", x, wt, out, n, h, w, c, c, c, None, res, mean, rstd, gam, bet, sums))
bias = torch.randn(c, device=DEV)
f16 = t(lambda: call("conv3x3_halo_gn", x, wt, out, n, h, w, c, c, c, bias, res, ws))
_lib.load().ffvc_set_option(b"halo_epi16", 0)
f8 = t(lambda: call("conv3x3_halo_gn", x, wt, out, n, h, w, c, c, c, bias, res, ws))
print("dgrad conv: plain %.1f us, with backward statistics 8 warps %.1f us, 16 warps %.1f us; fwd conv+res with statistics 8 warps %.1f, 16 warps %.1f" % (a, b, b16, f8, f16))
PY
run_bench() {
  FFVC_OPTS="halo_epi16=$2" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "== bench $1 rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/bench_$1.json')); print(round(d['value'],1), 'prompts/s', round(d['ms_per_step'],2), 'ms', 'gemm TF', round(d['roofline']['achieved'],1), d['roofline']['launches_per_step'])" 2>&1 | tail -1)"; tail -2 gpurun_out/bench_$1.err
}
run_bench e0 0
run_bench e1 1
run_bench e2 2
run_bench e0b 0
run_bench e2b 2
