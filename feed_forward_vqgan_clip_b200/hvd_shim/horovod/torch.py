"""`horovod.torch` surface used by the reference's train loop, implemented on torch.distributed (NCCL on the GPU box, gloo
on CPU) so that main.py runs data-parallel WITHOUT edits (the reference imports Horovod opportunistically at
main.py:45 and drives it from USE_HOROVOD):

    hvd.init / rank / local_rank / size                    main.py:304-307,529-531,620,671-672
    hvd.DistributedOptimizer(opt)                          main.py:627   gradient averaging before opt.step()
    hvd.broadcast_parameters / broadcast_optimizer_state   main.py:628-629
    hvd.broadcast(tensor, root_rank)                       main.py:686
    hvd.allreduce(tensor, average=True)                    main.py:367,839-842
    hvd.join()                                             main.py:375,390,1362

Semantics follow Horovod's defaults: allreduce / DistributedOptimizer AVERAGE over ranks.  When every parameter of the
wrapped optimizer is a view into one flat gradient arena (the B200 mappers re-point their parameters that way:
mixer.py / vitgan_mapper.py / xtransformer.py), DistributedOptimizer issues ONE all-reduce over the arena per step instead
of one per parameter.  One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).
"""
import os

import torch
import torch.distributed as dist

Average, Sum = "average", "sum"


def init(backend=None):
    if dist.is_available() and dist.is_initialized():
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank())
        kw["device_id"] = torch.device("cuda", local_rank())
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend, **kw)


def _on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def size():
    return dist.get_world_size() if _on() else 1


def rank():
    return dist.get_rank() if _on() else 0


def local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def local_size():
    return int(os.environ.get("LOCAL_WORLD_SIZE", str(size())))


def is_initialized():
    return True


def shutdown():
    if _on():
        dist.destroy_process_group()


def _comm_tensor(t):
    """NCCL reduces device tensors only; gloo host tensors."""
    if _on() and dist.get_backend() == "nccl" and not t.is_cuda:
        return t.cuda(), True
    return t, False


def allreduce(tensor, average=None, name=None, op=None):
    """Returns a new tensor; average=True (Horovod's default) divides by size()."""
    avg = (op in (None, Average)) if average is None else bool(average)
    if torch.is_tensor(tensor):
        out = tensor.detach().clone()
    else:
        out = torch.tensor(tensor)
    if not _on():
        return out
    buf, moved = _comm_tensor(out)
    if not buf.is_floating_point() and avg:
        buf = buf.float()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    if avg:
        buf = buf / size()
    return buf.to(out.device) if moved else buf


def allreduce_(tensor, average=None, name=None, op=None):
    tensor.copy_(allreduce(tensor, average=average, op=op))
    return tensor


def broadcast(tensor, root_rank=0, name=None):
    out = tensor.detach().clone()
    if not _on():
        return out
    buf, moved = _comm_tensor(out)
    dist.broadcast(buf, src=root_rank)
    return buf.to(out.device) if moved else buf


def broadcast_(tensor, root_rank=0, name=None):
    with torch.no_grad():                      # not through .data: the version counter must move (engines key their caches on it)
        tensor.copy_(broadcast(tensor, root_rank))
    return tensor


def broadcast_object(obj, root_rank=0, name=None):
    if not _on():
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=root_rank)
    return box[0]


def _flat_arena_of(tensors):
    """If all tensors are views into ONE contiguous storage (the mappers' flat arenas), return a 1-D view spanning them."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return None
    try:
        ptr0 = tensors[0].untyped_storage().data_ptr()
        if any(t.untyped_storage().data_ptr() != ptr0 or t.dtype != tensors[0].dtype or not t.is_contiguous() for t in tensors):
            return None
        lo = min(t.storage_offset() for t in tensors)
        hi = max(t.storage_offset() + t.numel() for t in tensors)
        return torch.as_strided(tensors[0], (hi - lo,), (1,), lo)
    except Exception:
        return None


def broadcast_parameters(params, root_rank=0):
    """params: a state_dict or an iterable of (name, tensor) (main.py:628)."""
    if not _on():
        return
    items = params.items() if hasattr(params, "items") else params
    tensors = [t for _, t in items if torch.is_tensor(t)]
    flat = _flat_arena_of(tensors)
    if flat is not None:
        broadcast_(flat, root_rank)
        return
    for t in tensors:
        broadcast_(t, root_rank)


def broadcast_optimizer_state(optimizer, root_rank=0):
    """main.py:629 — a fresh optimizer has no state yet; existing state tensors and the hyper-parameters are synchronised."""
    if not _on():
        return
    opt = getattr(optimizer, "_opt", optimizer)
    for group in opt.param_groups:
        hp = {k: v for k, v in group.items() if k != "params"}
        hp = broadcast_object(hp, root_rank)
        group.update(hp)
        for p in group["params"]:
            st = opt.state.get(p, {})
            for k in sorted(st):
                if torch.is_tensor(st[k]):
                    broadcast_(st[k], root_rank)


def join(device=None):
    if _on():
        dist.barrier()
    return 0


class _DistributedOptimizer:
    """opt.step() first averages the gradients over all ranks (Horovod op=Average): one all-reduce over the flat gradient
    arena when there is one, else one flattened bucket per dtype."""

    def __init__(self, optimizer, named_parameters=None, op=Average, **_unused):
        self._opt = optimizer
        self._average = op in (None, Average)

    def __getattr__(self, name):
        return getattr(self._opt, name)

    def synchronize(self):
        if not _on():
            return
        grads = [p.grad for g in self._opt.param_groups for p in g["params"] if p.grad is not None]
        if not grads:
            return
        flat = _flat_arena_of(grads)
        if flat is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if self._average:
                flat.div_(size())
            return
        by_type = {}
        for g in grads:
            by_type.setdefault((g.dtype, g.device), []).append(g)
        for bucket in by_type.values():
            buf = torch.cat([g.reshape(-1) for g in bucket])
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            if self._average:
                buf.div_(size())
            off = 0
            for g in bucket:
                g.copy_(buf[off:off + g.numel()].view_as(g))
                off += g.numel()

    def step(self, closure=None):
        self.synchronize()
        return self._opt.step(closure) if closure is not None else self._opt.step()

    def zero_grad(self, *a, **kw):
        return self._opt.zero_grad(*a, **kw)

    def state_dict(self):
        return self._opt.state_dict()

    def load_state_dict(self, sd):
        return self._opt.load_state_dict(sd)

    @property
    def param_groups(self):
        return self._opt.param_groups

    @property
    def state(self):
        return self._opt.state


def DistributedOptimizer(optimizer, named_parameters=None, op=Average, **kw):
    return _DistributedOptimizer(optimizer, named_parameters=named_parameters, op=op, **kw)
