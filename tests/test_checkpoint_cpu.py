"""Host logic of checkpoint.py on stand-in objects (no GPU): the files are the reference's formats (main.py:904-911) and load
with the reference's resume code (main.py:577-596); sharded writes reassemble to the same dictionaries.  The device snapshot
(D2D + pinned D2H on a side stream) is replaced by a host copy with the same copy()/wait() contract."""
import os
from types import SimpleNamespace

import torch

from feed_forward_vqgan_clip_b200 import checkpoint as ck


class HostSnapshot:
    def __init__(self, sizes, device):
        self.host_buf = [torch.empty(n, dtype=torch.float32) for n in sizes]

    def copy(self, slices):
        for h, s in zip(self.host_buf, slices):
            h.copy_(s)
        return self.host_buf

    def wait(self):
        pass


def _stand_in(use_ema=True, cosine=False):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.LayerNorm(3), torch.nn.Linear(3, 7))
    params = list(net.parameters())
    offs, o = [], 0
    for p in params:                                     # 8-element aligned slots like the engines' arenas
        offs.append(o)
        o += (p.numel() + 7) // 8 * 8
    arena = torch.zeros(o)
    for p, off in zip(params, offs):
        view = arena[off:off + p.numel()].view(p.shape)
        view.copy_(p.data)
        p.data = view
    eng = SimpleNamespace(m=net, params=params, arena=arena, total=o)
    hyper = torch.tensor([2e-4, .9, .999, 1e-8, 1, 1, 1, 0, 12, 0, 0, 1, 1e-3, 100.0 if cosine else 0.0, 0, 0.995])
    opt = SimpleNamespace(m=torch.randn(o), v=torch.rand(o), ema=(arena * 0.5 if use_ema else None), hyper=hyper, t=12, lr=1e-3,
                          betas=(0.9, 0.999), eps=1e-8, wd=0.0)
    return net, SimpleNamespace(mix=eng, opt=opt)


def test_async_writer_produces_the_reference_files(tmp_path):
    net, ts = _stand_in(use_ema=True, cosine=True)
    w = ck.CheckpointWriter(ts, str(tmp_path), config={"dim": 3}, snapshot_cls=HostSnapshot)
    w.save(step=40, epoch=2)
    ts.mix.arena.add_(1.0)                               # training goes on: the snapshot must not see it
    w.wait()
    assert sorted(os.listdir(tmp_path)) == ["checkpoint.th", "checkpoint_ema.th", "opt.th"]
    c = torch.load(tmp_path / "checkpoint.th", weights_only=False)
    assert c["step"] == 40 and c["epoch"] == 2 and c["config"] == {"dim": 3}
    assert list(c["state_dict"].keys()) == list(net.state_dict().keys())
    fresh = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.LayerNorm(3), torch.nn.Linear(3, 7))
    fresh.load_state_dict(c["state_dict"])               # main.py:581
    for (k, a), b in zip(fresh.state_dict().items(), net.state_dict().values()):
        assert torch.allclose(a, b - 1.0, atol=1e-6), k
    e = torch.load(tmp_path / "checkpoint_ema.th", weights_only=False)
    assert torch.equal(e["state_dict"]["0.weight"], 0.5 * c["state_dict"]["0.weight"])
    o = torch.load(tmp_path / "opt.th", weights_only=False)
    t = torch.optim.Adam(fresh.parameters(), lr=1.0)
    t.load_state_dict(o)                                 # main.py:593-596
    g = t.param_groups[0]
    assert abs(g["lr"] - 2e-4) < 1e-10 and abs(g["initial_lr"] - 1e-3) < 1e-9 and int(t.state_dict()["state"][0]["step"]) == 12
    lay = ck.param_layout(ts.mix)
    assert torch.equal(t.state_dict()["state"][2]["exp_avg"], ts.opt.m[lay[2][1]:lay[2][1] + lay[2][2]].view(lay[2][3]))
    # a second save reuses the buffers only after the first write has been joined
    w.save(step=41, epoch=2, blocking=True)
    assert torch.load(tmp_path / "checkpoint.th", weights_only=False)["step"] == 41


def test_sharded_writers_reassemble_to_the_single_writer_files(tmp_path):
    net, ts = _stand_in(use_ema=True)
    one, many = tmp_path / "one", tmp_path / "many"
    ck.CheckpointWriter(ts, str(one), config="cfg", snapshot_cls=HostSnapshot).save(7, 1, blocking=True)
    world = 3                                            # 72 arena elements do not divide evenly: ragged last shard
    for r in range(world):
        w = ck.CheckpointWriter(ts, str(many), config="cfg", rank=r, world=world, sharded=True, snapshot_cls=HostSnapshot)
        assert (w.hi - w.lo) <= (ts.mix.arena.numel() + world - 1) // world
        w.save(7, 1, blocking=True)
    assert sorted(os.listdir(many)) == ["checkpoint.index.th"] + ["checkpoint.shard-%02d-of-03.step-7.th" % r for r in range(3)]
    c, e, o = ck.load_sharded(str(many))
    c1 = torch.load(one / "checkpoint.th", weights_only=False)
    e1 = torch.load(one / "checkpoint_ema.th", weights_only=False)
    o1 = torch.load(one / "opt.th", weights_only=False)
    assert c["step"] == c1["step"] == 7 and c["config"] == "cfg"
    for k in c1["state_dict"]:
        assert torch.equal(c["state_dict"][k], c1["state_dict"][k]) and torch.equal(e["state_dict"][k], e1["state_dict"][k]), k
    assert o["param_groups"] == o1["param_groups"]
    for i in o1["state"]:
        assert torch.equal(o["state"][i]["exp_avg_sq"], o1["state"][i]["exp_avg_sq"])
    # a later generation whose shards are incomplete (a rank died before writing): the previous complete one is loaded
    writers = [ck.CheckpointWriter(ts, str(many), config="cfg", rank=r, world=world, sharded=True, snapshot_cls=HostSnapshot) for r in range(2)]
    for w in writers:                                    # rank 2 never writes step 9
        w.save(9, 1, blocking=True)
    assert os.path.exists(many / "checkpoint.index.prev.th")
    c2, _, _ = ck.load_sharded(str(many))
    assert c2["step"] == 7
    ck.CheckpointWriter(ts, str(many), config="cfg", rank=2, world=world, sharded=True, snapshot_cls=HostSnapshot).save(9, 1, blocking=True)
    assert ck.load_sharded(str(many))[0]["step"] == 9
    # replicas other than rank 0 stay silent in the reference's (unsharded) mode
    w = ck.CheckpointWriter(ts, str(tmp_path / "none"), rank=1, world=2, sharded=False, snapshot_cls=HostSnapshot)
    w.save(1, 0, blocking=True)
    assert os.listdir(tmp_path / "none") == []


def test_write_errors_surface_on_wait(tmp_path):
    net, ts = _stand_in(use_ema=False)
    w = ck.CheckpointWriter(ts, str(tmp_path / "d"), snapshot_cls=HostSnapshot)
    os.rmdir(tmp_path / "d")
    w.save(1, 0)
    try:
        w.wait()
        raise AssertionError("expected the failed write to be reported")
    except (OSError, RuntimeError):
        pass


def test_cuda_snapshot_refuses_host_arenas():
    net, ts = _stand_in()
    try:
        ck.CheckpointWriter(ts, "/tmp/never", config=None)
        raise AssertionError("expected a RuntimeError")
    except RuntimeError:
        pass
