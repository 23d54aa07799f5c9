#!/usr/bin/env python
"""bench.py — prompts/s for one train step of the feed-forward VQGAN-CLIP pipeline (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5]      (N>1: launched per rank by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W [--config C]

Workloads (BASELINE.json `configs`, numbered as there; `--config`, default 2 = the configuration the metric is quoted on):
  #2  MLP-Mixer 32x1024 mapper, VQGAN f16/16384 decoder, 8 cutouts, CLIP ViT-B/32, 256x256, 64 prompts per GPU
  #3  VitGAN 32x1024 mapper (main.py:459-468), otherwise as #2; 64 prompts per GPU (batch 512 over 8 GPUs)
  #4  X-transformer 256x16 mapper (main.py:488-499), 512x512 (32x32 latent grid), 16 prompts per GPU
  #5  MLP-Mixer 32x1024 at 512x512, OpenCLIP ViT-B-32 (exact GELU), LPIPS diversity (repeat 2) + TV loss (main.py:769-791),
      8 prompts x repeat 2 = 16 images per GPU
bf16 compute, synthetic text embeddings, seeded random-init weights.  A step = mapper fwd -> clamp -> VQ -> decode -> cutouts ->
CLIP -> loss (+ optional terms) -> full backward -> (NCCL grad all-reduce) -> Adam.  Nothing is skipped inside the timed region.

`value`   : whole-job prompts/s, inputs resident in HBM (real sampled augmentation parameters in the graph's static buffers),
            CUDA-graph replay, CUDA events, max over ranks.
`e2e`     : same metric through the public API with HOST inputs: per step a pinned-host -> device copy of the
            embeddings + augmentation parameters and a device -> host read of the loss.
`roofline`: the tcgen05 GEMM family (dominant), algorithmic FLOPs / CUDA-event time of its launches in one step.
`cpu_baseline` / `--impl reference`: the reference's CPU path (oracle restatement, torch fp32, all host threads)
            on a bounded sample of the same workload.
`--gpu-eager-baseline`: context only, never the target — the same oracle step executed by ATen / cuDNN / cuBLAS on this GPU
            (fp32 and bf16 autocast), i.e. what the reference's modules would do on the B200 box.
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

CUTN = 8
CLIP_DIM = 512
METRIC = "prompts/sec per train step (256x256, ViT-B/32)"

# per-prompt algorithmic FLOPs: BASELINE.md section 2 / SURVEY App. B (2*M*N*K; mapper x3, frozen nets x2, VQ forward only)
CONFIGS = {
    2: dict(workload="config#2: MLP-Mixer 32x1024 -> VQGAN f16/16384 decode -> 8 cutouts -> CLIP ViT-B/32, 256x256",
            cfg=dict(model_type="mlp_mixer", dim=1024, depth=32, vq_image_size=16), batch=64, kw={}, lpips=False,
            flops=dict(mapper=516.4e9, decoder=506.0e9, vq=2.15e9, clip=141.8e9, other=0.0)),
    3: dict(workload="config#3: VitGAN 32x1024 (6 heads) -> VQGAN f16/16384 decode -> 8 cutouts -> CLIP ViT-B/32, 256x256",
            cfg=dict(model_type="vitgan", dim=1024, depth=32, vq_image_size=16, num_heads=6), batch=64, kw={}, lpips=False,
            flops=dict(mapper=39.2e9, decoder=506.0e9, vq=2.15e9, clip=141.8e9, other=0.0)),
    4: dict(workload="config#4: X-transformer 256x16 (6 heads) -> VQGAN f16/16384 decode -> 8 cutouts -> CLIP ViT-B/32, 512x512",
            cfg=dict(model_type="xtransformer", dim=256, depth=16, vq_image_size=32, num_heads=6), batch=16, kw={}, lpips=False,
            flops=dict(mapper=168.7e9, decoder=2043.5e9, vq=8.59e9, clip=141.8e9, other=0.0)),
    5: dict(workload="config#5: MLP-Mixer 32x1024 -> VQGAN decode 512x512 -> 8 cutouts -> OpenCLIP ViT-B-32 (exact GELU), "
                     "+ LPIPS-VGG16 diversity (repeat 2) + TV loss",
            cfg=dict(model_type="mlp_mixer", dim=1024, depth=32, vq_image_size=32, clip_model="open_clip:ViT-B-32"), batch=8,
            kw=dict(tv_coef=0.1, diversity_coef=0.1, repeat=2), lpips=True,
            flops=dict(mapper=3302.6e9, decoder=2043.5e9, vq=8.59e9, clip=141.8e9, other=320.7e9)),
}


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


def load_traffic(config_id):
    """per-launch DRAM traffic of the tcgen05 GEMM family from the committed ncu launch list of one step of THIS config
    (tools/one_step.py under ncu --metrics ...dram__bytes_read.sum,dram__bytes_write.sum; tools/summarize_launches.py --json)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    suffix = "_traffic.json" if config_id == 2 else "_traffic_config%d.json" % config_id
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.endswith(suffix):
            best = os.path.join(pdir, f)
    if best is None:
        return None, None
    try:
        d = json.load(open(best))
        return d["gemm_family"]["dram_bytes_per_launch"], "profiles/" + os.path.basename(best)
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------ model construction
def build_b200(device, conf, world, pg, seed=0, rank=0):
    from feed_forward_vqgan_clip_b200 import api
    from feed_forward_vqgan_clip_b200.train_step import TrainStep
    torch.manual_seed(seed)                       # same seed on every rank = identical replicas (main.py:628)
    net = api.build_model(conf["cfg"])
    vq = api.load_vqgan_model()
    with torch.no_grad():
        vq.quantize.embedding.weight.normal_(0, 1)   # N(0,1) codebook (taming's U(+-1/n) makes ties)
    clip = api.load_clip_model(conf["cfg"].get("clip_model", "ViT-B/32"))
    net, vq, clip = net.to(device), vq.to(device).eval().requires_grad_(False), clip.to(device).eval().requires_grad_(False)
    kw = dict(conf["kw"])
    if conf["lpips"]:
        from feed_forward_vqgan_clip_b200.lpips import LpipsVGG16
        kw["lpips_net"] = LpipsVGG16().to(device).eval().requires_grad_(False)
    # replicas share the weights' seed; the augmentation stream is per rank (every rank cuts its own prompts differently)
    return TrainStep(net, vq, clip, cutn=CUTN, lr=1e-3, world_size=world, process_group=pg, seed=seed + 17 + 1000 * rank, **kw)


def synthetic_embeddings(n, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, CLIP_DIM, generator=g) * 0.45).float()     # |x| ~ 10 like raw CLIP text embeddings


# ------------------------------------------------------------------------------------------------ oracle trainers (CPU arm / GPU eager context)
def oracle_trainer(conf, device="cpu"):
    """the reference's step restated in plain torch (oracle/) for this config, on `device`"""
    import oracle.clip_vit as oclip
    import oracle.mixer as omix
    import oracle.vqgan as ovq
    from oracle.train_step import OracleTrainer
    cfg = conf["cfg"]
    S = cfg["vq_image_size"]
    if cfg["model_type"] == "mlp_mixer":
        sd_m, mapper = omix.init_mixer_state_dict(CLIP_DIM, S, 256, cfg["dim"], cfg["depth"], seed=0), "mixer"
    else:   # the other mappers: seeded parameter containers of this package (construction only; the arithmetic is the oracle's)
        from feed_forward_vqgan_clip_b200 import api
        torch.manual_seed(0)
        sd_m = {k: v.detach().clone() for k, v in api.build_model(cfg).state_dict().items()}
        mapper = cfg["model_type"]
    sd_v, sd_c = ovq.init_vqgan_state_dict(seed=1), oclip.init_clip_state_dict(seed=2)
    kw = dict(conf["kw"])
    sd_vgg = None
    if conf["lpips"]:
        import oracle.lpips as olpips
        sd_vgg = {k: v.to(device) for k, v in olpips.init_vgg_state_dict(seed=4).items()}
    mv = lambda sd: {k: v.to(device) for k, v in sd.items()}                                       # noqa: E731
    act = "gelu" if "open_clip" in cfg.get("clip_model", "") and "quickgelu" not in cfg.get("clip_model", "") else "quick_gelu"
    return OracleTrainer(mv(sd_m), mv(sd_v), mv(sd_c), S, 256, cutn=CUTN, act=act, mapper=mapper, num_heads=cfg.get("num_heads", 6),
                         tv_coef=kw.get("tv_coef", 0.0), repeat=kw.get("repeat", 1), diversity_coef=kw.get("diversity_coef", 0.0),
                         sd_vgg=sd_vgg)


def _oracle_step_time(tr, conf, batch, steps, warmup, device="cpu", autocast=None):
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    g = torch.Generator().manual_seed(3)
    rep = conf["kw"].get("repeat", 1)
    times = []
    for i in range(warmup + steps):
        x = synthetic_embeddings(batch, 100 + i).to(device)
        prm = sample_params(CUTN * batch * rep, 224, g)
        prm = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in prm.items()}
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if autocast is not None:
            with torch.autocast("cuda", dtype=autocast):
                tr.step(x, x, prm)
        else:
            tr.step(x, x, prm)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def cpu_reference(conf, steps, warmup, batches):
    """times the oracle step on the host cores for every batch in `batches`; the arm's value is the best of them"""
    # torchrun exports OMP_NUM_THREADS=1: the reference arm must use all the host cores it can
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    torch.set_num_threads(max(1, cores))
    cores = torch.get_num_threads()
    tr = oracle_trainer(conf)
    by_batch = {}
    for b in batches:
        s = _oracle_step_time(tr, conf, b, steps, warmup)
        by_batch[b] = dict(prompts_per_s=b / s, ms_per_step=1e3 * s)
    best = max(by_batch, key=lambda b: by_batch[b]["prompts_per_s"])
    return dict(value=by_batch[best]["prompts_per_s"], ms_per_step=by_batch[best]["ms_per_step"], cores=cores, batch=best,
                by_batch={str(b): v for b, v in by_batch.items()},
                sample="%d timed step(s) after %d warm-up at each of %s prompt(s) per step (%s nets, fp32, torch CPU, %d threads); "
                       "value = the best batch (%d)" % (steps, warmup, "/".join(str(b) for b in batches),
                                                        conf["workload"].split(":")[0], cores, best))


def gpu_eager_baseline(conf, device, batch):
    """CONTEXT, not the target: the oracle step (the reference's modules restated in plain torch) executed eagerly by ATen /
    cuDNN / cuBLAS on this GPU — fp32 (TF32 convolutions as torch defaults) and bf16 autocast."""
    out = {"batch": batch, "what": "oracle/ train step run eagerly on cuda by ATen / cuDNN / cuBLAS (torch defaults: cuDNN TF32 on, matmul fp32)"}
    tr = oracle_trainer(conf, device)
    for name, ac in (("fp32", None), ("bf16_autocast", torch.bfloat16)):
        try:
            s = _oracle_step_time(tr, conf, batch, 3, 2, device, ac)
            out[name] = dict(prompts_per_s=batch / s, ms_per_step=1e3 * s)
        except Exception as e:  # noqa: BLE001
            out[name] = dict(error=repr(e)[:200])
    return out


def _finish(world, ts=None):
    """Orderly NCCL teardown: drop the captured graph (it references NCCL's kernels and buffers) before the communicator goes.
    Round 1 left with os._exit(0) unconditionally because destroy_process_group() could block while the graph was alive; now
    the exit is only the watchdog of a teardown that takes more than 30 s (and says so)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world <= 1:
        return
    import gc
    torch.cuda.synchronize()

    def watchdog():
        time.sleep(30.0)
        sys.stderr.write("bench.py: NCCL teardown still blocked after 30 s - leaving with os._exit(0)\n")
        sys.stderr.flush()
        os._exit(0)

    threading.Thread(target=watchdog, daemon=True).start()
    if ts is not None:
        ts.graph = None
        ts.static = None
    gc.collect()
    torch.cuda.synchronize()
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[] number (2 = the metric's)")
    ap.add_argument("--batch", type=int, default=0, help="prompts per GPU (0 = the config's default)")
    ap.add_argument("--cpu-sample-batch", type=str, default="", help="comma-separated prompts per CPU step of the reference arm / cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-eager-baseline", action="store_true", help="also time the oracle step eagerly on this GPU (context only)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--bucket-layers", type=int, default=None,
                    help="N>1: mixer layers per gradient all-reduce bucket (overlapped with backward); 0 = one all-reduce after backward")
    ap.add_argument("--tail-overlap", action="store_true",
                    help="N>1: run Adam on the already-reduced slices while the last gradient bucket is in flight")
    ap.add_argument("--no-adam-overlap", action="store_true",
                    help="N>1: one Adam launch after the last all-reduce instead of per-bucket updates behind each bucket's all-reduce")
    ap.add_argument("--comm-sms", type=int, default=0,
                    help="N>1: SMs left to NCCL while gradient buckets are in flight — sets NCCL_MAX_CTAS and sizes the persistent GEMM grids "
                         "to (SMs - this) during the overlapped part of backward; 0 = off")
    ap.add_argument("--nccl-max-ctas", type=int, default=0,
                    help="N>1: cap the CTAs NCCL may use (NCCL_MAX_CTAS) so its kernels take fewer SMs from the persistent GEMMs they overlap; 0 = NCCL's default")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    conf = CONFIGS[args.config]
    B = args.batch or conf["batch"]
    rep = conf["kw"].get("repeat", 1)
    config = {"workload": conf["workload"], "per_gpu_batch": B, "global_batch": B * world, "cutn": CUTN,
              "parallelism": "dp%d" % world, "l2": "inputs larger than L2 (tens of GB of activations per step)"}
    if rep > 1:
        config["repeat"] = rep
        config["images_per_gpu_step"] = B * rep
    if world > 1:
        config["grad_allreduce"] = "flat fp32 arena, buckets of mapper layers overlapped with backward on a side stream (NCCL)"

    if args.impl == "reference":
        if rank != 0:
            return
        K, W = max(1, min(args.steps, 2)), max(0, min(args.warmup, 1))
        if args.cpu_sample_batch:
            batches = [int(b) for b in args.cpu_sample_batch.split(",")]
        else:
            batches = [2, 8] if args.config in (2, 3) else [1, 2]
        r = cpu_reference(conf, K, W, batches)
        config = dict(config, per_gpu_batch=r["batch"], global_batch=r["batch"], parallelism="cpu",
                      note="CPU arm: the batch it actually ran (best of %s prompts per step); the GPU arm runs %d per GPU" % (batches, B))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "prompts/s", "n_gpus": args.gpus,
                          "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": r["value"], "unit": "prompts/s", "cores": r["cores"], "kind": "port",
                                           "sample": r["sample"], "by_batch": r["by_batch"]},
                          "e2e": {"value": r["value"], "unit": "prompts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from feed_forward_vqgan_clip_b200 import ops
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        if args.nccl_max_ctas > 0 or args.comm_sms > 0:
            os.environ["NCCL_MAX_CTAS"] = str(args.nccl_max_ctas or args.comm_sms)
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    ts = build_b200(dev, conf, world, pg, rank=rank)
    if args.bucket_layers is not None:
        ts.bucket_layers = args.bucket_layers
    ts.tail_overlap = bool(args.tail_overlap)
    ts.comm_sms = args.comm_sms if world > 1 else 0
    ts.adam_overlap = not args.no_adam_overlap
    if world > 1:
        config.update(bucket_layers=ts.bucket_layers, tail_overlap=ts.tail_overlap, comm_sms=ts.comm_sms, adam_overlap=ts.adam_overlap,
                      nccl_max_ctas=os.environ.get("NCCL_MAX_CTAS"))
    x_host = [synthetic_embeddings(B, 1000 + 7919 * rank + i).pin_memory() for i in range(4)]
    Bimg = B * rep                                            # rows of the (repeated) batch = generated images per step

    ops.reset_launch_count()
    use_graph = not args.no_graph
    if use_graph:
        try:
            ts.capture(Bimg, CLIP_DIM)
        except Exception as e:                                # e.g. a collective that refuses capture: run eagerly instead
            sys.stderr.write("CUDA-graph capture failed (%r); falling back to eager launches\n" % (e,))
            torch.cuda.synchronize()
            use_graph = False
            ts.graph = None
    if use_graph:
        launches_per_step = ops.launch_count() // 2          # capture() runs the body twice (warm-up + capture)
        run = lambda i: ts.replay(x_host[i % 4], None, None)               # noqa: E731
        # device-resident leg: embeddings AND one real draw of the augmentation parameters (affine / perspective maps, jitter,
        # erase rectangle) sit in the graph's static buffers before the timed region starts
        ts.load_static(x_host[0], None, ts.new_params(Bimg))
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
    N = CUTN * Bimg
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
            traffic, traffic_src = load_traffic(args.config)
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
        _finish(world, ts)
        return

    fp = conf["flops"]
    fp_total = (fp["mapper"] + fp["decoder"] + fp["vq"] + fp["clip"] + fp["other"]) * rep      # per PROMPT: `repeat` images each
    peaks, peak_src = load_peaks()
    line = {"metric": METRIC, "value": value, "unit": "prompts/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "prompts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches_per_step) * args.steps * 2,
            "launches_per_step": int(launches_per_step), "cuda_graph": use_graph, "last_loss": loss_host,
            "algorithmic_tflop_per_prompt": fp_total / 1e12,
            "step_tflops_achieved": fp_total * value / world / 1e12,
            "step_frac_of_bf16_peak": fp_total * value / world / 1e12 / peaks.get("bf16_tflops_sustained", 1400.0)}
    if rep > 1:
        line["images_per_s"] = value * rep
    if roof is not None:
        line["roofline"] = roof
    if world == 1 and not args.no_cpu_baseline:
        if args.cpu_sample_batch:
            batches = [int(b) for b in args.cpu_sample_batch.split(",")]
        else:
            batches = [2, 8] if args.config in (2, 3) else [1]
        r = cpu_reference(conf, 1, 1, batches)        # ~25 s of host time at config #2: (1 warm-up + 1 timed step) x (2 and 8 prompts)
        line["cpu_baseline"] = {"value": r["value"], "unit": "prompts/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                "by_batch": r["by_batch"]}
    if world == 1 and args.gpu_eager_baseline:
        del ts
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(conf, dev, 8 if args.config in (2, 3) else 2)
        ts = None
    print(json.dumps(line))
    _finish(world, ts)


if __name__ == "__main__":
    main()
