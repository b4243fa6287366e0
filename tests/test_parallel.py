"""N > 1 host logic on CPU: world-size-2 gloo all-reduce of gradients (what bench.py --gpus N does over NCCL)
and ray sharding."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from nsr_b200.parallel import GradSync, shard_rays
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(1000)), torch.nn.Parameter(torch.zeros(7, 3))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    GradSync(params, world).all_reduce_mean()
    # make_grad_sync: CPU parameters / no symmetric memory -> the NCCL-style path (here gloo) with a description saying so
    from nsr_b200.parallel import make_grad_sync
    s2, desc = make_grad_sync(params, world)
    assert isinstance(s2, GradSync) and 'all-reduce' in desc
    rays = torch.arange(8 * 6, dtype=torch.float32).view(8, 6)
    mine = shard_rays(rays, rank, world)
    q.put((rank, [float(p.grad.flatten()[0]) for p in params], mine[:, 0].tolist()))
    dist.destroy_process_group()


def test_grad_allreduce_mean_and_ray_sharding_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g, shard in res:
        assert g == [1.5, 3.0]            # mean of (1,2) and of (2,4)
        assert shard == [rank * 24.0 + 6.0 * i for i in range(4)]
