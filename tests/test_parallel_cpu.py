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
