"""Data-parallel plumbing of the train step (replaces the Horovod calls of main.py:528-531,627-629,668-674,838-842).

One process per GPU (torchrun), full replicas, the prompt batch is sharded across ranks, and ONE all-reduce of the
mapper's flat gradient arena per step; the 1/world average is folded into the fused Adam's grad_scale.  The process
group is NCCL on the GPU box; the same functions run over gloo on CPU for the host-logic tests."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (torch.distributed.run)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of a global batch, like DistributedSampler with drop_last (main.py:668-674)."""
    per = n_total // world
    return rank * per, (rank + 1) * per


def shard_cutout_params(prm, cutn, n_total, lo, hi):
    """Augmentation parameters of the prompts [lo, hi) of a global batch of n_total prompts.
    MakeCutouts orders its output cutout-major — cutout k of prompt j is row k * B + j (main.py:219, `repeat(cutn, ...)`) —
    so a rank that owns prompts [lo, hi) needs rows {k * n_total + j : lo <= j < hi} of every per-cutout tensor, re-packed
    as k * (hi - lo) + (j - lo).  The erase rectangle is one per batch (RandomErasing(same_on_batch=True), main.py:189-190)
    and is shared.  With parameters sharded this way a data-parallel step over any world size reproduces the single-process
    step on the global batch (tests/test_parallel_cpu.py, tests/test_zz_d_full_size_gpu.py)."""
    out = {}
    for k, v in prm.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == cutn * n_total:
            out[k] = v.reshape(cutn, n_total, *v.shape[1:])[:, lo:hi].reshape(cutn * (hi - lo), *v.shape[1:]).contiguous()
        else:
            out[k] = v
    return out


def allreduce_flat_grads(flat_grad, world, group=None):
    """Sum-all-reduce the flat gradient arena in place; returns the scale Adam must apply (1/world)."""
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def bucket_ranges(layer_starts, total, layers_per_bucket):
    """Contiguous [lo, hi) slices of the flat gradient arena in the order backward completes them.
    layer_starts: arena offset of the first parameter of every mixer layer, ascending (layer 0 first); backward finishes
    the layers last-to-first, so bucket k covers layers [L - (k+1)*n, L - k*n) plus, for k = 0, everything registered
    after the last layer (final norm, output projection); the head of the arena (everything registered before the first layer,
    finished last) is the final slice.  The slices tile [0, total) exactly once.  Parameters that are registered after the
    layers but whose gradient is only complete at the very end of backward must be carved out: see bucket_slices."""
    L = len(layer_starts)
    out, hi, i = [], total, L
    while i > 0:
        j = max(0, i - layers_per_bucket)
        out.append((layer_starts[j], hi))
        hi, i = layer_starts[j], j
    if hi > 0:
        out.append((0, hi))
    return out


def bucket_slices(layer_starts, total, layers_per_bucket, late=()):
    """bucket_ranges with the `late` ranges carved out: a list of buckets, each a list of [lo, hi) slices.  Bucket k < last may be
    all-reduced as soon as backward has finished layer L - (k+1)*n; the LAST bucket — the head of the arena plus every `late`
    range — only after backward has ended.  `late`: arena ranges whose gradient is complete only then although they are
    registered after the layers: the mixer's input projection `proj` sits between the final norm and `final_proj` in the
    reference's registration order (mlp_mixer_pytorch.py:73-76) but its wgrad is the last GEMM of backward.  Reducing it with
    the first bucket would sum zeros and leave every replica with its own local gradient.
    The slices of all buckets together tile [0, total) exactly once."""
    late = sorted((int(lo), int(hi)) for lo, hi in late if hi > lo)
    ranges = bucket_ranges(layer_starts, total, layers_per_bucket)
    head = []
    if ranges and ranges[-1][0] == 0 and (not layer_starts or ranges[-1][1] <= layer_starts[0]):
        head = [ranges.pop()]
    buckets = []
    for lo, hi in ranges:
        pieces, cur = [], lo
        for a, b in late:
            a, b = max(a, lo), min(b, hi)
            if b <= a:
                continue
            if a > cur:
                pieces.append((cur, a))
            cur = max(cur, b)
        if hi > cur:
            pieces.append((cur, hi))
        buckets.append(pieces)
    buckets.append(head + list(late))
    return buckets


def broadcast_flat(flat, src=0, group=None):
    """hvd.broadcast_parameters / broadcast_optimizer_state equivalent on a flat arena (main.py:628-629)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def allreduce_scalars(values, world, group=None):
    """The 4 logging scalars of main.py:838-842 in ONE all-reduce (average)."""
    t = torch.stack([v.reshape(()) for v in values]).clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= world
    return t


def install_horovod_shim():
    """Make `import horovod.torch as hvd` (main.py:45) resolve to the torch.distributed-backed shim shipped in
    feed_forward_vqgan_clip_b200/hvd_shim, unless a real Horovod is installed.  Returns the module."""
    import importlib
    import sys
    try:
        import horovod.torch as hvd                      # a real installation wins
        return hvd
    except Exception:
        pass
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hvd_shim")
    if here not in sys.path:
        sys.path.insert(0, here)
    for k in [k for k in sys.modules if k == "horovod" or k.startswith("horovod.")]:
        del sys.modules[k]
    return importlib.import_module("horovod.torch")
