#!/usr/bin/env python
"""bench.py — prompts/s for one train step of the feed-forward VQGAN-CLIP pipeline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched per rank by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): MLP-Mixer 32x1024 mapper, VQGAN f16/16384 decoder, 8 cutouts, CLIP ViT-B/32,
256x256, bf16 compute, batch 64 prompts per GPU, synthetic text embeddings, seeded random-init weights.
A step = mapper fwd -> clamp -> VQ -> decode -> cutouts -> CLIP -> loss -> full backward -> (NCCL grad all-reduce)
-> Adam.  Nothing is skipped inside the timed region.

`value`   : whole-job prompts/s, inputs resident in HBM, CUDA-graph replay, CUDA events, max over ranks.
`e2e`     : same metric through the public API with HOST inputs: per step a pinned-host -> device copy of the
            embeddings + augmentation parameters and a device -> host read of the loss.
`roofline`: the tcgen05 GEMM kernel (dominant), algorithmic FLOPs / CUDA-event time of its launches in one step.
`cpu_baseline` / `--impl reference`: the reference's CPU path (oracle restatement, torch fp32, all host threads)
            on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

MIXER = dict(input_dim=512, image_size=16, channels=256, patch_size=1, dim=1024, depth=32)
CUTN = 8
CLIP_DIM = 512
METRIC = "prompts/sec per train step (256x256, ViT-B/32)"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def load_traffic():
    """per-launch DRAM traffic of the tcgen05 GEMM family from the committed ncu launch list of one step
    (tools/one_step.py under ncu --metrics ...dram__bytes_read.sum,dram__bytes_write.sum; tools/summarize_launches.py --json)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith("_traffic.json"):
            best = os.path.join(pdir, f)
    if best is None:
        return None, None
    try:
        d = json.load(open(best))
        return d["gemm_family"]["dram_bytes_per_launch"], "profiles/" + os.path.basename(best)
    except Exception:
        return None, None


def flops_per_prompt():
    """SURVEY App. B formulas, config #2 (S=16, D=1024, L=32, cutn=8)."""
    T, C, D, L = 256, 256, 1024, 32
    mixer_fwd = 2 * 512 * T * C + 2 * T * C * D + L * (4 * T * 4 * T * D + 4 * T * D * 4 * D) + 2 * T * D * C
    vit = 2 * 49 * 3072 * 768 + 12 * (2 * 50 * 768 * 2304 + 4 * 50 * 50 * 768 + 2 * 50 * 768 * 768 + 16 * 50 * 768 * 768) + 2 * 768 * 512
    return dict(mapper=3 * mixer_fwd, decoder=506.0e9, vq=2.15e9, clip=CUTN * 2 * vit, total=3 * mixer_fwd + 506.0e9 + 2.15e9 + CUTN * 2 * vit)


# ------------------------------------------------------------------------------------------------ model construction
def build_b200(device, batch, world, pg, seed=0, rank=0):
    from feed_forward_vqgan_clip_b200.clip_vit import CLIP
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    from feed_forward_vqgan_clip_b200.train_step import TrainStep
    from feed_forward_vqgan_clip_b200.vqgan import VQModel
    torch.manual_seed(seed)                       # same seed on every rank = identical replicas (main.py:628)
    net = Mixer(**MIXER)
    vq = VQModel()
    with torch.no_grad():
        vq.quantize.embedding.weight.normal_(0, 1)   # N(0,1) codebook (taming's U(+-1/n) makes ties)
    clip = CLIP()
    net, vq, clip = net.to(device), vq.to(device).eval().requires_grad_(False), clip.to(device).eval().requires_grad_(False)
    # replicas share the weights' seed; the augmentation stream is per rank (every rank cuts its own prompts differently)
    return TrainStep(net, vq, clip, cutn=CUTN, lr=1e-3, world_size=world, process_group=pg, seed=seed + 17 + 1000 * rank)


def synthetic_embeddings(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, CLIP_DIM, generator=g) * 0.45).float()     # |x| ~ 10 like raw CLIP text embeddings


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(steps, warmup, sample_batch):
    import oracle.clip_vit as oclip
    import oracle.mixer as omix
    import oracle.vqgan as ovq
    from oracle.train_step import OracleTrainer
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    # torchrun exports OMP_NUM_THREADS=1: the reference arm must use all the host cores it can
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, cores))
    cores = torch.get_num_threads()
    sd_m = omix.init_mixer_state_dict(CLIP_DIM, MIXER["image_size"], MIXER["channels"], MIXER["dim"], MIXER["depth"], seed=0)
    sd_v = ovq.init_vqgan_state_dict(seed=1)
    sd_c = oclip.init_clip_state_dict(seed=2)
    tr = OracleTrainer(sd_m, sd_v, sd_c, MIXER["image_size"], MIXER["channels"], cutn=CUTN)
    g = torch.Generator().manual_seed(3)
    times = []
    for i in range(warmup + steps):
        x = synthetic_embeddings(sample_batch, 100 + i)
        prm = sample_params(CUTN * sample_batch, 224, g)
        t0 = time.perf_counter()
        tr.step(x, x, prm)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return dict(value=sample_batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=cores,
                sample="%d step(s) of %d prompt(s) (config #2 nets, fp32, torch CPU, %d threads) after %d warm-up"
                       % (len(times), sample_batch, cores, warmup))


def _finish(world):
    """Leave without tearing NCCL down: destroying the communicator while a captured CUDA graph still references its
    kernels can block forever (observed at N=2); the OS reclaims everything."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="prompts per GPU")
    ap.add_argument("--cpu-sample-batch", type=int, default=2, help="prompts per CPU step of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--bucket-layers", type=int, default=None,
                    help="N>1: mixer layers per gradient all-reduce bucket (overlapped with backward); 0 = one all-reduce after backward")
    ap.add_argument("--tail-overlap", action="store_true",
                    help="N>1: run Adam on the already-reduced slices while the last gradient bucket is in flight (opt-in, unmeasured)")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="N>1: cap the CTAs NCCL may use (NCCL_MAX_CTAS) so its kernels take fewer SMs from the persistent GEMMs they overlap; 0 = NCCL's default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "config#2: MLP-Mixer 32x1024 -> VQGAN f16/16384 decode -> 8 cutouts -> CLIP ViT-B/32, 256x256",
              "per_gpu_batch": args.batch, "global_batch": args.batch * world, "cutn": CUTN,
              "parallelism": "dp%d" % world, "l2": "inputs larger than L2 (tens of GB of activations per step)"}
    if world > 1:
        config["grad_allreduce"] = "flat fp32 arena, buckets of mixer layers overlapped with backward on a side stream (NCCL)"

    if args.impl == "reference":
        if rank != 0:
            return
        K, W = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
        r = cpu_reference(K, W, args.cpu_sample_batch)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "prompts/s", "n_gpus": args.gpus,
                          "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": "prompts/s", "cores": r["cores"], "kind": "port",
                                           "sample": r["sample"]},
                          "e2e": {"value": r["value"], "unit": "prompts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from feed_forward_vqgan_clip_b200 import ops
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        if args.nccl_max_ctas > 0:
            os.environ["NCCL_MAX_CTAS"] = str(args.nccl_max_ctas)
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    B = args.batch
    ts = build_b200(dev, B, world, pg, rank=rank)
    if args.bucket_layers is not None:
        ts.bucket_layers = args.bucket_layers
    ts.tail_overlap = bool(args.tail_overlap)
    x_host = [synthetic_embeddings(B, 1000 + 7919 * rank + i).pin_memory() for i in range(4)]

    ops.reset_launch_count()
    use_graph = not args.no_graph
    if use_graph:
        try:
            ts.capture(B, CLIP_DIM)
        except Exception as e:                                # e.g. a collective that refuses capture: run eagerly instead
            sys.stderr.write("CUDA-graph capture failed (%r); falling back to eager launches\n" % (e,))
            torch.cuda.synchronize()
            use_graph = False
            ts.graph = None
    if use_graph:
        launches_per_step = ops.launch_count() // 2          # capture() runs the body twice (warm-up + capture)
        run = lambda i: ts.replay(x_host[i % 4], None, None)               # noqa: E731
        ts.static["inp"].copy_(x_host[0])
        run_dev = lambda i: ts.graph.replay()                              # noqa: E731 (inputs already in HBM)
    else:
        xd = [x.to(dev) for x in x_host]
        ops.reset_launch_count()
        ts.step(xd[0])
        launches_per_step = ops.launch_count()
        run = lambda i: ts.step(x_host[i % 4].to(dev, non_blocking=True))  # noqa: E731
        run_dev = lambda i: ts.step(xd[i % 4])                             # noqa: E731

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value)
    for i in range(args.warmup):
        run_dev(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None      # one nvidia-smi poller per job, on the rank that reports
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run_dev(i)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    total_ms = ms.item()
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step / 1e3)

    # ---------------- end-to-end through the public API with host inputs (e2e)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = run(i)
        loss_host = loss.item()                                    # D2H read of the step's loss
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = B * world * args.steps / e2e_s.item()
    N = CUTN * B
    h2d = B * CLIP_DIM * 4 * 2 + N * (9 + 9 + 1 + 1) * 4 + 16

    # ---------------- roofline of the dominant kernel: instrumented eager step, CUDA events around every GEMM launch
    roof = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        ev, fl, by = [], [], []
        real_gemm = ops.gemm

        def timed_gemm(a, b, out, M, Nn, K, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = real_gemm(a, b, out, M, Nn, K, **kw)
            e.record()
            ev.append((s, e))
            nb, segs = kw.get("batch", 1), kw.get("k_segs", 1)       # `batch` is the total (outer x inner) batch count
            fl.append(2.0 * M * Nn * K * nb * segs)
            # algorithmic DRAM bytes of this launch: each operand once (broadcast operands are shared by the batch), the
            # output once, plus every extra epilogue stream (pre-activation copy, act' operand, residual)
            a_n = (nb if kw.get("a_role", 0) == 1 else 1) * (segs if kw.get("a_role", 0) == 2 else 1)
            b_n = (nb if kw.get("b_role", 0) == 1 else 1) * (segs if kw.get("b_role", 0) == 2 else 1)
            if kw.get("a_mode", 0) == 2:          # implicit-GEMM 3x3 conv: the NHWC tensor is read once, K = 9 * Cin
                a_bytes = 2.0 * M * K / 9
            else:
                a_bytes = 2.0 * M * K * a_n
            o_sz = 4.0 if (out is not None and out.dtype == torch.float32) else 2.0
            extra = sum(2.0 for k_ in ("pre_out", "aux", "res") if kw.get(k_) is not None)
            by.append(a_bytes + 2.0 * Nn * K * b_n + M * Nn * nb * (o_sz + extra))
            return r

        from feed_forward_vqgan_clip_b200 import vqgan as _vq_mod
        real_call = _vq_mod.call

        def timed_call(name, *a):
            if not name.startswith("conv3x3_halo"):   # the halo-reuse 3x3 conv (plain / + GroupNorm statistics) is the same tcgen05 family
                return real_call(name, *a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real_call(name, *a)
            e.record()
            ev.append((s, e))
            n_, h_, w_, cin_, cout_ = a[3:8]
            fl.append(2.0 * n_ * h_ * w_ * cout_ * 9 * cin_)
            if name == "conv3x3_halo":                     # (..., ldc, bias, res, aux, mul_mode, act, out_fp32)
                extra, o_sz = (a[10] is not None) + (a[11] is not None), (4.0 if a[14] else 2.0)
            elif name == "conv3x3_halo_gn":                # (..., ldc, bias, res, gn_ws)
                extra, o_sz = (a[10] is not None), 2.0
            else:                                          # conv3x3_halo_gnbwd: (..., ldc, res, gn_x, ...): reads the Normalize's input too
                extra, o_sz = (a[9] is not None) + 1, 2.0
            by.append(n_ * h_ * w_ * (2.0 * cin_ + o_sz * cout_ + 2.0 * cout_ * extra) + 2.0 * 9 * cin_ * cout_)

        xd0 = x_host[0].to(dev)
        if world == 1:
            ts.step(xd0)                       # first eager step after graph mode pays for fresh allocations: not timed
            torch.cuda.synchronize()
            ops.gemm = timed_gemm
            _vq_mod.call = timed_call
            del ev[:], fl[:], by[:]
            s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_all.record()
            ts.step(xd0)
            e_all.record()
            torch.cuda.synchronize()
            gemm_ms = sum(s.elapsed_time(e) for s, e in ev)
            gemm_flops = sum(fl)
            achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
            peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
            traffic, traffic_src = load_traffic()
            roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel + conv3x3_halo_kernel (tcgen05 GEMM family)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_unit": "DRAM bytes per launch (read + write), mean over the family's launches of one step",
                    "traffic_source": traffic_src, "algorithmic_bytes_per_launch": sum(by) / max(1, len(by)),
                    "flops_per_launch": gemm_flops / max(1, len(ev)), "peak_source": peak_src + " (sustained cuBLAS bf16)",
                    "launches_per_step": len(ev), "gemm_ms_per_step": gemm_ms, "eager_step_ms": s_all.elapsed_time(e_all),
                    "gemm_share_of_step": gemm_ms / ms_per_step, "flops_per_step_executed": gemm_flops,
                    "how": "CUDA events around every ffvc_gemm / ffvc_conv3x3_halo launch of one eager step; share = their time / graph step time"}
            _vq_mod.call = real_call
        ops.gemm = real_gemm

    if rank != 0:
        _finish(world)
        return

    fp = flops_per_prompt()
    peaks, peak_src = load_peaks()
    line = {"metric": METRIC, "value": value, "unit": "prompts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "prompts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step) * args.steps * 2,
            "launches_per_step": int(launches_per_step), "cuda_graph": use_graph, "last_loss": loss_host,
            "algorithmic_tflop_per_prompt": fp["total"] / 1e12,
            "step_tflops_achieved": fp["total"] * value / world / 1e12,
            "step_frac_of_bf16_peak": fp["total"] * value / world / 1e12 / peaks.get("bf16_tflops_sustained", 1400.0)}
    if roof is not None:
        line["roofline"] = roof
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(2, 1, args.cpu_sample_batch)          # ~10 s of host time: 1 warm-up + 2 timed steps of 2 prompts
        line["cpu_baseline"] = {"value": r["value"], "unit": "prompts/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line))
    _finish(world)


if __name__ == "__main__":
    main()
