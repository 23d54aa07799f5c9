"""The data-parallel branch of TrainStep on the GPU with its real CUDA streams and events: 2 processes share cuda:0 and exchange
gradients over gloo (NCCL refuses two ranks on one device; gloo all-reduces CUDA tensors through the host, stream-ordered), which
is enough to check what a one-GPU box can check of SURVEY §8e — the bucketed all-reduces issued from backward's per-layer
callbacks on the side stream, the late `proj` slice in the last bucket, the 1/world average in the fused Adam — against the
single-process step on the global batch.  (tests/test_dp_step_cpu.py checks the same branch on the CPU; the N-GPU NCCL runs are
bench.py's.)  Sorts last; every wait is bounded and the workers are killed on a time-out so that a stuck collective cannot
hang the suite."""
import os
import queue
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

VQ = dict(ch=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(16,), resolution=32, z_channels=64, out_ch=3,
          embed_dim=64, n_embed=512)
CLIPCFG = dict(input_resolution=224, patch_size=32, width=128, layers=2, heads=2, output_dim=64)
CUTN, B, DEPTH = 2, 4, 3


def _build(world, pg, bucket_layers):
    import oracle.clip_vit as oclip
    import oracle.vqgan as ovq
    from feed_forward_vqgan_clip_b200.clip_vit import CLIP
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    from feed_forward_vqgan_clip_b200.train_step import TrainStep
    from feed_forward_vqgan_clip_b200.vqgan import VQModel
    torch.manual_seed(0)
    net = Mixer(input_dim=64, image_size=16, channels=64, patch_size=1, dim=128, depth=DEPTH)
    with torch.no_grad():
        net.final_proj.weight.mul_(6.0)
    vq = VQModel(VQ)
    vq.load_state_dict(ovq.init_vqgan_state_dict(VQ, seed=8))
    clip = CLIP(CLIPCFG)
    clip.visual.load_state_dict(oclip.init_clip_state_dict(CLIPCFG, seed=9))
    dev = "cuda:0"
    ts = TrainStep(net.to(dev), vq.to(dev).eval().requires_grad_(False), clip.to(dev).eval().requires_grad_(False), cutn=CUTN,
                   lr=1e-3, world_size=world, process_group=pg)
    ts.bucket_layers = bucket_layers
    return net, ts


def _inputs():
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, 64, generator=g) * 0.45
    return x, sample_params(CUTN * B, 224, g)


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from feed_forward_vqgan_clip_b200 import parallel
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        net, ts = _build(world, dist.group.WORLD, 2)
        x, prm = _inputs()
        lo, hi = parallel.shard_range(B, rank, world)
        loss = float(ts.step(x[lo:hi].contiguous().cuda(), None, parallel.shard_cutout_params(prm, CUTN, B, lo, hi)).item())
        torch.cuda.synchronize()
        q.put((rank, loss, ts.mix.grad.cpu().numpy().copy(), ts.mix.arena.cpu().numpy().copy()))
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put((rank, "ERROR", traceback.format_exc(), None))


def test_two_rank_step_on_one_gpu_equals_the_global_batch_step():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        try:
            res = sorted((q.get(timeout=240) for _ in range(2)), key=lambda t: t[0])
        except queue.Empty:
            pytest.fail("the two-rank step did not finish within 240 s")
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    for r in res:
        assert not isinstance(r[1], str), r[2]
    (_, loss0, g0, p0), (_, loss1, g1, p1) = res
    g0, g1, p0, p1 = (torch.from_numpy(a) for a in (g0, g1, p0, p1))
    assert torch.equal(g0, g1) and torch.equal(p0, p1)       # every slice of the arena was reduced: identical replicas after the step
    net, ts = _build(1, None, 0)
    keep = ts.mix.arena.clone()
    x, prm = _inputs()
    loss = float(ts.step(x.cuda(), None, prm).item())
    g_full, p_full = ts.mix.grad.cpu().double(), ts.mix.arena.cpu()
    assert abs((loss0 + loss1) / 2 - loss) <= 5e-3 * abs(loss)
    a = (g0 / 2).double()
    assert float(torch.dot(a, g_full) / (a.norm() * g_full.norm())) > 0.99
    lo, hi = ts.mix.late_ranges()[0]                          # the input projection really carries a gradient from BOTH ranks
    assert float(torch.dot(a[lo:hi], g_full[lo:hi]) / (a[lo:hi].norm() * g_full[lo:hi].norm())) > 0.99
    du, dr = (p0 - keep.cpu()).double(), (p_full - keep.cpu()).double()
    assert float(torch.dot(du, dr) / (du.norm() * dr.norm())) > 0.9
