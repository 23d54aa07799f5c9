"""The data-parallel branch of TrainStep._device_step (main.py:627: Horovod averages the ranks' gradients) run for
real with world_size 2 over gloo on the CPU: bucketed all-reduces of the flat gradient arena issued from the per-layer
completion callback of the mixer's backward, the head-of-arena slice after it, the 1/world average folded into the fused Adam —
against the single-process step on the global batch.  Device kernels are replaced by tests/abi_model.py (see
test_engine_orchestration_cpu.py); the CUDA stream / event calls of that branch, which only order work on the device, are
replaced by no-ops (on the CPU everything runs in program order)."""
import contextlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

VQ = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(16,), resolution=32, z_channels=64, out_ch=3,
          embed_dim=64, n_embed=512)
CLIPCFG = dict(input_resolution=64, patch_size=32, width=128, layers=1, heads=2, output_dim=64)
CUTN, B, CUT, DEPTH = 2, 4, 64, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _NoStream:
    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass


class _NoEvent:
    def record(self, stream=None):
        pass


def _patch(setattr_=setattr):
    """in the spawned workers the replacements are permanent; in the pytest process they go through monkeypatch.setattr"""
    import abi_model
    from feed_forward_vqgan_clip_b200 import (clip_vit, cutouts, mixer, ops, simple_vitgan_mapper, train_step, vitgan_mapper, vqgan,
                                              xtransformer)
    setattr_(ops, "gemm_raw", abi_model.gemm_raw)
    setattr_(ops, "gemm", lambda a, b, out, M, N, K, **kw: abi_model.gemm_raw(a, b, out, M, N, K, **kw))
    setattr_(ops, "call", abi_model.call)
    setattr_(ops, "require_cuda", lambda dev, what: None)
    for mod in (mixer, vqgan, cutouts, train_step, clip_vit, vitgan_mapper, simple_vitgan_mapper, xtransformer):
        setattr_(mod, "call", abi_model.call)
    setattr_(torch.cuda, "Stream", lambda device=None: _NoStream())
    setattr_(torch.cuda, "current_stream", lambda device=None: _NoStream())
    setattr_(torch.cuda, "Event", lambda *a, **k: _NoEvent())
    setattr_(torch.cuda, "stream", lambda s: contextlib.nullcontext())


def _mapper(name):
    from feed_forward_vqgan_clip_b200 import mixer, simple_vitgan_mapper, vitgan_mapper, xtransformer
    if name == "mixer":
        net = mixer.Mixer(input_dim=64, image_size=16, channels=64, patch_size=1, dim=64, depth=DEPTH)
        out = net.final_proj.weight
    elif name == "vitgan":
        net = vitgan_mapper.Generator(initialize_size=2, dim=48, blocks=DEPTH, num_heads=3, out_channels=64, input_dim=64)
        out = net.w_out[0].weight
    elif name == "simple_vitgan":
        net = simple_vitgan_mapper.SimpleGenerator(size=16, dim=48, blocks=DEPTH, num_heads=3, out_channels=64, input_dim=64)
        out = net.w_out[0].weight
    else:
        net = xtransformer.XTransformer(input_dim=64, image_size=16, channels=64, dim=64, depth=DEPTH, heads=3, initial_proj=True,
                                        add_input=False)
        out = net.transformer.project_out.weight
    with torch.no_grad():
        out.mul_(5.0)                                        # spread z over the codebook range
    return net


def _build(world, pg, bucket_layers, mapper="mixer", tail_overlap=False):
    import oracle.clip_vit as oclip
    import oracle.vqgan as ovq
    from feed_forward_vqgan_clip_b200 import clip_vit, train_step, vqgan
    torch.manual_seed(0)                                     # identical replicas on every rank (main.py:628 broadcasts)
    net = _mapper(mapper)
    vq = vqgan.VQModel(VQ)
    vq.load_state_dict(ovq.init_vqgan_state_dict(VQ, seed=8))
    clip = clip_vit.CLIP(CLIPCFG)
    clip.visual.load_state_dict(oclip.init_clip_state_dict(CLIPCFG, seed=9))
    ts = train_step.TrainStep(net, vq.eval().requires_grad_(False), clip.eval().requires_grad_(False), cutn=CUTN, lr=1e-3,
                              cut_size=CUT, world_size=world, process_group=pg)
    ts.bucket_layers = bucket_layers
    ts.tail_overlap = tail_overlap
    return net, ts


def _inputs():
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, 64, generator=g) * 0.45
    return x, sample_params(CUTN * B, CUT, g)


def _worker(rank, world, port, bucket_layers, mapper, tail_overlap, q):
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.set_num_threads(2)
        _patch()
        from feed_forward_vqgan_clip_b200 import parallel
        dist.init_process_group("gloo", rank=rank, world_size=world)
        calls = []
        real = dist.all_reduce
        dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
        torch.distributed.all_reduce = dist.all_reduce
        net, ts = _build(world, dist.group.WORLD, bucket_layers, mapper, tail_overlap)
        x, prm = _inputs()
        lo, hi = parallel.shard_range(B, rank, world)
        loss = float(ts.step(x[lo:hi].contiguous(), None, parallel.shard_cutout_params(prm, CUTN, B, lo, hi)))
        eng = ts.mix
        # numpy arrays travel through the queue by value (torch tensors would be shared-memory handles that die with the worker)
        q.put((rank, loss, eng.grad.numpy().copy(), eng.arena.numpy().copy(), list(calls), eng.total,
               ts.last_indices.numpy().copy()))
        dist.destroy_process_group()
    except Exception as e:                                   # surface the failure instead of a queue time-out
        import traceback
        q.put((rank, "ERROR", traceback.format_exc()))


def _run(bucket_layers, mapper="mixer", tail_overlap=False):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bucket_layers, mapper, tail_overlap, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] != "ERROR", r[2]
    return [(r, l, torch.from_numpy(g), torch.from_numpy(p), c, t, torch.from_numpy(i)) for r, l, g, p, c, t, i in res]


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))


import pytest


@pytest.mark.parametrize("mapper", ["mixer", "vitgan", "simple_vitgan", "xtransformer"])
def test_two_rank_step_equals_the_global_batch_step(monkeypatch, mapper):
    res = _run(bucket_layers=2, mapper=mapper)                              # 3 mixer layers in buckets of 2: [layers 1-2 + tail], [layer 0], [head]
    (_, loss0, g0, p0, calls0, total, idx0), (_, loss1, g1, p1, calls1, _, idx1) = res
    assert torch.equal(g0, g1) and torch.equal(p0, p1)       # both replicas hold the same reduced gradient and take the same step
    assert calls0 == calls1 and sum(calls0) == total         # the all-reduced slices tile the arena exactly once
    # single process, global batch
    _patch(monkeypatch.setattr)
    net, ts = _build(1, None, 0, mapper)
    keep = ts.mix.arena.clone()
    x, prm = _inputs()
    loss = float(ts.step(x, None, prm))
    g_full, p_full = ts.mix.grad, ts.mix.arena
    from feed_forward_vqgan_clip_b200 import parallel
    eng = ts.mix
    buckets = parallel.bucket_slices(eng.layer_starts(), eng.total, 2, late=eng.late_ranges())
    assert calls0 == [hi - lo for b in buckets for lo, hi in b] and len(buckets) == 3
    # an input projection registered after the layers finishes last: it must travel in the LAST bucket, not the first
    for lo, hi in eng.late_ranges():
        assert (lo, hi) in buckets[-1] and all(not (a < hi and lo < b) for bk in buckets[:-1] for a, b in bk)
        assert eng.offs["proj.weight"] == lo and float(g_full[lo:hi].abs().max()) > 0
    assert bool(eng.late_ranges()) == (mapper in ("mixer", "xtransformer"))
    same = (torch.cat([idx0.view(B // 2, -1), idx1.view(B // 2, -1)]) == ts.last_indices.view(B, -1)).float().mean()
    assert float(same) > 0.99
    assert abs((loss0 + loss1) / 2 - loss) < 5e-3 * abs(loss)
    assert _cos(g0 / 2, g_full) > 0.99                       # sum over ranks, averaged by Adam's grad_scale = 1 / world
    assert _cos(p0 - keep, p_full - keep) > 0.9              # the same Adam update (sign-like first step: tiny gradients may flip)


def test_single_allreduce_form_gives_the_same_reduced_gradient():
    res_b = _run(bucket_layers=2)
    res_s = _run(bucket_layers=0)                            # one all-reduce of the whole arena after backward
    assert len(res_s[0][4]) == 1 and res_s[0][4][0] == res_s[0][5]
    assert torch.allclose(res_b[0][2], res_s[0][2], rtol=1e-5, atol=1e-7)
    # opt-in: Adam on the already-reduced slices while the last bucket is in flight — the same update, slice by slice
    res_t = _run(bucket_layers=2, tail_overlap=True)
    assert torch.equal(res_t[0][2], res_b[0][2]) and torch.equal(res_t[0][3], res_b[0][3]) and torch.equal(res_t[1][3], res_b[1][3])
