"""Checkpoint I/O of the train step in the reference's own file formats, without stalling the GPU.

The reference saves synchronously on rank 0 every `log_interval` steps (main.py:904-911):

    checkpoint.th       {"state_dict": net.state_dict(), "config", "step", "epoch"}
    checkpoint_ema.th   the same dict with the EMA weights copied in (torch_ema `average_parameters`)
    opt.th              torch.optim.Adam.state_dict()

and resumes from them (main.py:577-596,598-616).  Here the mapper's weights, Adam moments and EMA live in flat fp32 arenas on
the device, so a save is: (1) ONE device-to-device snapshot of those arenas on the training stream (a few GB at HBM speed,
~1 ms), (2) a device-to-host copy of the snapshot into pinned memory on a side stream while training continues, (3) a host
thread that waits for the copy, rebuilds the three dictionaries above as views of the host buffers and `torch.save`s them
(write to a temporary name, then an atomic rename).  Files written this way load with the reference's resume code unchanged.

Sharded form (`CheckpointWriter(..., rank, world, sharded=True)`): data-parallel replicas hold identical arenas, so rank r
snapshots and writes only elements [r*n/world, (r+1)*n/world) of every arena to `checkpoint.shard-RR-of-WW.th`; rank 0 adds
`checkpoint.index.th` (parameter names, shapes, arena offsets, config, step, epoch).  `load_sharded(folder)` reassembles the flat
arenas and returns the three reference-format dictionaries.  Each rank moves 1/world of the bytes over its own PCIe link.
"""
import os
import threading

import torch


def _atomic_save(obj, path):
    tmp = path + ".tmp"
    torch.save(obj, tmp)
    os.replace(tmp, path)


def param_layout(engine):
    """[(name, arena offset in elements, numel, shape)] of the mapper's parameters inside the engine's flat arenas."""
    out = []
    base = engine.arena.data_ptr()
    for (name, _), p in zip(engine.m.named_parameters(), engine.params):
        off = p.data_ptr() - base
        assert off % 4 == 0 and 0 <= off // 4 and off // 4 + p.numel() <= engine.arena.numel(), name
        out.append((name, off // 4, p.numel(), tuple(p.shape)))
    return out


def state_dict_from_arena(layout, arena, extra=None):
    """the module's state_dict() rebuilt from a flat (host) arena: parameters are views at their offsets, `extra` carries
    whatever else the module's state_dict holds (buffers; the mappers have none)"""
    sd = dict(extra or {})
    for name, off, n, shape in layout:
        sd[name] = arena[off:off + n].view(shape)
    return sd


def adam_state_from_arenas(layout, m, v, step, group):
    """torch.optim.Adam.state_dict() layout (what main.py:593-596 loads and :911 saves as opt.th)"""
    state = {}
    if step > 0:
        for i, (_, off, n, shape) in enumerate(layout):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m[off:off + n].view(shape),
                        "exp_avg_sq": v[off:off + n].view(shape)}
    g = dict(group)
    g["params"] = list(range(len(layout)))
    return {"state": state, "param_groups": [g]}


class _CudaSnapshot:
    """device snapshot + pinned host buffers for a fixed list of fp32 slices; copy() is stream-ordered, wait() blocks the calling
    host thread until the host buffers hold the snapshot"""

    def __init__(self, sizes, device):
        if device.type != "cuda":
            raise RuntimeError("checkpoints are taken from the device-resident arenas of a CUDA train step")
        self.dev_buf = [torch.empty(n, device=device, dtype=torch.float32) for n in sizes]
        self.host_buf = [torch.empty(n, dtype=torch.float32).pin_memory() for n in sizes]
        self.side = torch.cuda.Stream(device=device)
        self.done = torch.cuda.Event()

    def copy(self, slices):
        main = torch.cuda.current_stream()
        for d, s in zip(self.dev_buf, slices):
            d.copy_(s, non_blocking=True)                  # (1) D2D on the training stream: a consistent cut between two steps
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            for h, d in zip(self.host_buf, self.dev_buf):
                h.copy_(d, non_blocking=True)              # (2) D2H on the side stream, overlapping the next steps
            self.done.record(self.side)
        return self.host_buf

    def wait(self):
        self.done.synchronize()


class CheckpointWriter:
    """ckpt = CheckpointWriter(train_step, folder, config); ckpt.save(step, epoch) returns at once; ckpt.wait() joins the
    write in flight (save() does so itself before it reuses the buffers)."""

    def __init__(self, train_step, folder, config=None, rank=0, world=1, sharded=False, snapshot_cls=None):
        self.ts, self.folder, self.config = train_step, folder, config
        self.rank, self.world, self.sharded = rank, world, bool(sharded) and world > 1
        self.eng, self.opt = train_step.mix, train_step.opt
        self.layout = param_layout(self.eng)
        n = self.eng.arena.numel()
        if self.sharded:
            per = (n + world - 1) // world
            self.lo, self.hi = min(n, rank * per), min(n, (rank + 1) * per)
        else:
            self.lo, self.hi = 0, n
        self.names = ["arena", "m", "v"] + (["ema"] if self.opt.ema is not None else [])
        self.active = self.sharded or rank == 0                       # main.py:904-911 runs on rank 0 only
        sizes = [self.hi - self.lo] * len(self.names) + [self.opt.hyper.numel()]      # + Adam's device-side scalar block
        self.snap = (snapshot_cls or _CudaSnapshot)(sizes, self.eng.arena.device) if self.active else None
        self.thread, self.error = None, None
        self._steps = []
        os.makedirs(folder, exist_ok=True)
        import atexit
        atexit.register(self._finish_at_exit)

    def _finish_at_exit(self):
        try:
            self.wait()
        except Exception as e:                                        # nobody is left to call wait(): say it
            import sys
            sys.stderr.write("CheckpointWriter: the last save failed: %r\n" % (e,))

    def _arenas(self):
        a = {"arena": self.eng.arena, "m": self.opt.m, "v": self.opt.v}
        if self.opt.ema is not None:
            a["ema"] = self.opt.ema
        return a

    def wait(self):
        if self.thread is not None:
            self.thread.join()
            self.thread = None
        if self.error is not None:
            err, self.error = self.error, None
            raise err

    def save(self, step, epoch, blocking=False):
        if not self.active:
            return
        self.wait()                                                   # the buffers of the previous save are free again
        arenas = self._arenas()
        host = self.snap.copy([arenas[k][self.lo:self.hi] for k in self.names] + [self.opt.hyper])
        extra = {k: v.detach().cpu() for k, v in self.eng.m.state_dict().items() if k not in {n for n, _, _, _ in self.layout}}

        def work():
            try:
                self.snap.wait()
                bufs = dict(zip(self.names, host))
                adam_step, group = _adam_group(self.opt, host[len(self.names)])
                if self.sharded:
                    # shards carry their step in the file name and the previous generation is kept: ranks write independently (no
                    # barrier), so after a crash the newest index may point at a step some rank never finished — load_sharded then
                    # falls back to the previous index, whose shards are all still there
                    _atomic_save({"lo": self.lo, "hi": self.hi, "total": self.eng.arena.numel(), "step": step, "epoch": epoch,
                                  **{k: bufs[k] for k in self.names}},
                                 os.path.join(self.folder, _shard_name(self.rank, self.world, step)))
                    if self.rank == 0:
                        idx = os.path.join(self.folder, "checkpoint.index.th")
                        if os.path.exists(idx):
                            os.replace(idx, os.path.join(self.folder, "checkpoint.index.prev.th"))
                        _atomic_save({"layout": self.layout, "config": self.config, "step": step, "epoch": epoch, "world": self.world,
                                      "adam_step": adam_step, "adam_group": group, "names": self.names, "extra": extra}, idx)
                    self._steps.append(step)
                    for old in self._steps[:-2]:                      # keep this generation and the previous one
                        try:
                            os.remove(os.path.join(self.folder, _shard_name(self.rank, self.world, old)))
                        except OSError:
                            pass
                    del self._steps[:-2]
                    return
                write_reference_files(self.folder, self.layout, bufs, self.config, step, epoch, adam_step, group, extra)
            except Exception as e:                                     # surfaced by the next wait() / save()
                self.error = e

        self.thread = threading.Thread(target=work, daemon=False)    # a process that exits right after save() still finishes the write
        self.thread.start()
        if blocking:
            self.wait()


def _adam_group(opt, hyper):
    """step count and param_group of torch.optim.Adam.state_dict() from a host copy of FusedAdam's scalar block
    (include/ffvc.h: [0] current lr, [8] step, [12] initial lr, [13] cosine T_max) — same fields as FusedAdam.state_dict()"""
    group = {"lr": float(hyper[0]), "betas": tuple(opt.betas), "eps": opt.eps, "weight_decay": opt.wd, "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "decoupled_weight_decay": False}
    if float(hyper[13]) > 0:
        group["initial_lr"] = float(hyper[12])
    return int(hyper[8]), group


def write_reference_files(folder, layout, bufs, config, step, epoch, adam_step, group, extra=None):
    """checkpoint.th / checkpoint_ema.th / opt.th exactly as main.py:904-911 writes them"""
    _atomic_save({"state_dict": state_dict_from_arena(layout, bufs["arena"], extra), "config": config, "step": step, "epoch": epoch},
                 os.path.join(folder, "checkpoint.th"))
    if "ema" in bufs:
        _atomic_save({"state_dict": state_dict_from_arena(layout, bufs["ema"], extra), "config": config, "step": step,
                      "epoch": epoch}, os.path.join(folder, "checkpoint_ema.th"))
    _atomic_save(adam_state_from_arenas(layout, bufs["m"], bufs["v"], adam_step, group), os.path.join(folder, "opt.th"))


def _shard_name(rank, world, step):
    return "checkpoint.shard-%02d-of-%02d.step-%d.th" % (rank, world, step)


def load_sharded(folder):
    """Reassemble a sharded checkpoint: returns (checkpoint dict, ema checkpoint dict or None, Adam state_dict) in the
    reference's formats — what `write_reference_files` would have written from one rank.  Uses the newest index whose shards
    are all present (a crash between the ranks' writes leaves the previous generation complete)."""
    index = None
    for name in ("checkpoint.index.th", "checkpoint.index.prev.th"):
        path = os.path.join(folder, name)
        if not os.path.exists(path):
            continue
        cand = torch.load(path, map_location="cpu", weights_only=False)
        if all(os.path.exists(os.path.join(folder, _shard_name(r, cand["world"], cand["step"]))) for r in range(cand["world"])):
            index = cand
            break
    if index is None:
        raise RuntimeError("no complete sharded checkpoint in %s" % folder)
    world, names = index["world"], index["names"]
    flat = None
    for r in range(world):
        sh = torch.load(os.path.join(folder, _shard_name(r, world, index["step"])), map_location="cpu", weights_only=False)
        if sh["step"] != index["step"]:
            raise RuntimeError("shard %d is from step %d, the index from step %d" % (r, sh["step"], index["step"]))
        if flat is None:
            flat = {k: torch.empty(sh["total"], dtype=torch.float32) for k in names}
        for k in names:
            flat[k][sh["lo"]:sh["hi"]] = sh[k]
    layout = [(n, o, c, tuple(s)) for n, o, c, s in index["layout"]]
    meta = dict(config=index["config"], step=index["step"], epoch=index["epoch"])
    ckpt = {"state_dict": state_dict_from_arena(layout, flat["arena"], index.get("extra")), **meta}
    ema = {"state_dict": state_dict_from_arena(layout, flat["ema"], index.get("extra")), **meta} if "ema" in flat else None
    opt = adam_state_from_arenas(layout, flat["m"], flat["v"], index["adam_step"], index["adam_group"])
    return ckpt, ema, opt
