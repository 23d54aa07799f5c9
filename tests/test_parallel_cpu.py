"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, the single flat-gradient all-reduce with
the average folded into the optimizer scale, parameter broadcast, scalar reduction, identical seeded replicas."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from feed_forward_vqgan_clip_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, lr, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(128, rank, world)
    # per-rank "gradient": rank-dependent, the average is known in closed form
    g = torch.full((1000,), float(rank + 1))
    scale = parallel.allreduce_flat_grads(g, world)
    avg = g * scale
    from feed_forward_vqgan_clip_b200.mixer import Mixer
    torch.manual_seed(0)
    net = Mixer(32, 4, 16, 1, 32, 1)
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    ref = flat.clone()
    if rank != 0:
        flat.add_(1.0)                      # corrupt, then restore from rank 0
    parallel.broadcast_flat(flat, 0)
    sc = parallel.allreduce_scalars([torch.tensor(float(rank)), torch.tensor(2.0)], world)
    q.put((rank, lo, hi, avg[0].item(), torch.equal(flat, ref), sc.tolist(), ref.sum().item()))
    dist.destroy_process_group()


def test_two_rank_gloo_data_parallel_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, a0, ok0, s0, sum0), (r1, lo1, hi1, a1, ok1, s1, sum1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 64, 64, 128)
    assert a0 == a1 == 1.5                      # mean of the per-rank gradients 1 and 2
    assert ok0 and ok1                          # broadcast restored rank 1's parameters
    assert s0 == s1 == [0.5, 2.0]
    assert sum0 == sum1                         # same seed -> identical replicas


# ----------------------------------------------------------------------------------------------------- horovod.torch shim
def _hvd_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    hvd = parallel.install_horovod_shim()
    hvd.init("gloo")
    assert (hvd.rank(), hvd.size(), hvd.local_rank()) == (rank, world, rank)
    torch.manual_seed(rank)                                   # DIFFERENT initial weights per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)         # main.py:591
    opt = hvd.DistributedOptimizer(opt)                       # main.py:627
    hvd.broadcast_parameters(net.state_dict(), root_rank=0)   # main.py:628
    hvd.broadcast_optimizer_state(opt, root_rank=0)           # main.py:629
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    x = torch.full((4, 8), float(rank + 1))                   # rank-dependent batch shard
    loss = net(x).pow(2).mean()
    opt.zero_grad()
    loss.backward()
    g_local = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    opt.step()                                                # averages the gradients first
    g_avg = torch.cat([p.grad.flatten() for p in net.parameters()]).clone()
    w1 = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    noise = hvd.broadcast(torch.full((3,), float(rank)), root_rank=0)          # main.py:686
    lsum = hvd.allreduce(torch.tensor(float(rank + 1)), average=False)         # main.py:367
    lavg = hvd.allreduce(torch.tensor(float(rank + 1)))                        # main.py:839
    # flat-arena fast path: parameters that are views of one buffer get a single all-reduce
    arena = torch.zeros(24)
    pa, pb = torch.nn.Parameter(arena[:16].view(4, 4)), torch.nn.Parameter(arena[16:24])
    garena = torch.full((24,), float(rank + 1))
    pa.grad, pb.grad = garena[:16].view(4, 4), garena[16:24]
    o2 = hvd.DistributedOptimizer(torch.optim.SGD([pa, pb], lr=1.0))
    o2.step()
    hvd.join()
    # plain lists: tensors on an mp queue are shared through file descriptors that die with this process
    q.put((rank, w0.tolist(), g_local.tolist(), g_avg.tolist(), w1.tolist(), noise.tolist(), float(lsum), float(lavg), arena.tolist()))
    hvd.shutdown()


def test_horovod_shim_two_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_hvd_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    T = torch.tensor
    (_, w0a, gla, gaa, w1a, na, sa, aa, ara), (_, w0b, glb, gab, w1b, nb, sb, ab, arb) = res
    w0a, gla, gaa, w1a, ara, w0b, glb, gab, w1b, arb = (T(v) for v in (w0a, gla, gaa, w1a, ara, w0b, glb, gab, w1b, arb))
    assert torch.equal(w0a, w0b)                                       # broadcast_parameters made the replicas identical
    assert not torch.allclose(gla, glb)                                # the shards really differ
    assert torch.allclose(gaa, (gla + glb) / 2, atol=1e-7) and torch.equal(gaa, gab)   # Horovod default: average
    assert torch.equal(w1a, w1b)                                       # so the replicas stay identical after the step
    assert na == nb == [0.0, 0.0, 0.0]
    assert sa == sb == 3.0 and aa == ab == 1.5
    assert torch.allclose(ara, torch.full((24,), -1.5)) and torch.equal(ara, arb)      # SGD lr=1 on the averaged arena gradient


def test_horovod_shim_single_process_is_a_no_op():
    hvd = parallel.install_horovod_shim()
    assert hvd.size() == 1 and hvd.rank() == 0
    t = torch.arange(4.0)
    assert torch.equal(hvd.allreduce(t), t) and torch.equal(hvd.broadcast(t, 0), t)
    opt = hvd.DistributedOptimizer(torch.optim.SGD([torch.nn.Parameter(torch.ones(2))], lr=0.1))
    opt.step()
    hvd.join()


# ----------------------------------------------------------------------------------------------------- bucketed all-reduce
def test_bucket_ranges_tile_the_arena_in_backward_order():
    import random
    rnd = random.Random(0)
    for _ in range(50):
        L = rnd.randint(1, 40)
        starts, o = [], rnd.choice([0, 8, 4096])
        for _l in range(L):
            starts.append(o)
            o += 8 * rnd.randint(1, 50)
        total = o + 8 * rnd.randint(0, 20)
        n = rnd.randint(1, 12)
        r = parallel.bucket_ranges(starts, total, n)
        assert r[0][1] == total and r[-1][0] == 0                       # tail first (finished first), head last
        assert all(a[0] == b[1] for a, b in zip(r, r[1:]))              # contiguous, descending, no overlap
        assert all(lo < hi for lo, hi in r)
        assert all(lo in starts or lo == 0 for lo, hi in r)             # cuts only at layer boundaries


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env("gloo")
    g = torch.Generator().manual_seed(100 + rank)
    grad = torch.randn(5000, generator=g)
    whole = grad.clone()
    dist.all_reduce(whole)
    starts = [40 + 155 * i for i in range(32)]
    for lo, hi in parallel.bucket_ranges(starts, 5000, 8):
        dist.all_reduce(grad[lo:hi])                                   # slices are views: reduced in place
    q.put((rank, torch.equal(grad, whole)))
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_single_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_sharded_step_reproduces_the_global_batch_step_oracle():
    """Size-independent property of the path: the train step is a mean over independent prompts, so the
    gradient of the global batch equals the average of the shard gradients when every shard gets ITS prompts' cutout
    parameters (parallel.shard_cutout_params).  Checked here on the CPU oracle; tests/test_zz_d_full_size_gpu.py checks the
    same property on the CUDA path at BASELINE config #2's full size."""
    import oracle.clip_vit as oclip
    import oracle.mixer as omix
    import oracle.vqgan as ovq
    from oracle.train_step import OracleTrainer
    from feed_forward_vqgan_clip_b200.cutouts import sample_params
    vq_cfg = dict(ch=32, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(4,), resolution=8, z_channels=32, out_ch=3,
                  embed_dim=32, n_embed=64)
    clip_cfg = dict(input_resolution=64, patch_size=32, width=64, layers=1, heads=2, output_dim=32)
    S, C, cutn, B, cut = 4, 32, 3, 4, 64
    sd_m = omix.init_mixer_state_dict(32, S, C, 64, 1, seed=0)
    sd_m["final_proj.weight"] = sd_m["final_proj.weight"] * 6.0
    sd_v, sd_c = ovq.init_vqgan_state_dict(vq_cfg, seed=1), oclip.init_clip_state_dict(clip_cfg, seed=2)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, 32, generator=g) * 0.45
    prm = sample_params(cutn * B, cut, g)

    def run(xs, ps):
        tr = OracleTrainer(sd_m, sd_v, sd_c, S, C, vq_cfg, clip_cfg, cutn=cutn, cut_size=cut)
        loss = tr.step(xs, xs, ps)
        return loss, tr.grads, tr.last_indices

    loss_full, g_full, idx_full = run(x, prm)
    world = 2
    losses, grads, idx = [], [], []
    for r in range(world):
        lo, hi = parallel.shard_range(B, r, world)
        ps = parallel.shard_cutout_params(prm, cutn, B, lo, hi)
        assert ps["affine_inv"].shape[0] == cutn * (hi - lo) and ps["erase"] == prm["erase"]
        # row k * Bs + (j - lo) of the shard is row k * B + j of the global batch
        assert torch.equal(ps["hue"][1 * (hi - lo) + 1], prm["hue"][1 * B + lo + 1])
        l, gr, ix = run(x[lo:hi], ps)
        losses.append(l), grads.append(gr), idx.append(ix.reshape(hi - lo, -1))
    assert torch.equal(torch.cat(idx).reshape(-1), idx_full.reshape(-1))
    assert abs(sum(losses) / world - loss_full) < 1e-5 * abs(loss_full)
    for k in g_full:
        avg = sum(gr[k] for gr in grads) / world
        assert torch.allclose(avg, g_full[k], rtol=1e-3, atol=1e-6 * float(g_full[k].abs().max()) + 1e-9), k


def test_bucket_slices_carve_out_late_ranges_and_still_tile_the_arena():
    """parallel.bucket_slices: ranges whose gradient completes only at the end of backward (the mixer's input projection, which
    the reference registers after the layers) leave the early buckets and join the head of the arena in the last one."""
    starts, total = [100, 300, 500, 700, 900], 1500          # 5 layers of 200, head [0, 100), tail [1100, 1500)
    late = [(1200, 1400)]
    for n in (1, 2, 3, 8):
        buckets = parallel.bucket_slices(starts, total, n, late=late)
        assert len(buckets) == -(-5 // n) + 1
        flat = sorted(s for b in buckets for s in b)
        assert flat[0][0] == 0 and flat[-1][1] == total and all(a[1] == b[0] for a, b in zip(flat, flat[1:]))
        assert buckets[-1] == [(0, 100), (1200, 1400)]
        assert buckets[0][-2:] == [(starts[max(0, 5 - n)], 1200), (1400, 1500)] if n < 5 else True
        assert all(not (lo < 1400 and 1200 < hi) for b in buckets[:-1] for lo, hi in b)
    # without late ranges the slices are bucket_ranges' own
    assert [s for b in parallel.bucket_slices(starts, total, 2) for s in b] == parallel.bucket_ranges(starts, total, 2)
    # no head (first layer starts at 0): the last bucket holds the late ranges only
    assert parallel.bucket_slices([0, 200], 500, 1, late=[(450, 500)])[-1] == [(450, 500)]
