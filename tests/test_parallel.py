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


def test_level_group_ranges_tile_the_exchange_buffer():
    """host logic of the pipelined exchange (P2PGradSync.bind_pipelined): the ranges of the level groups tile the buffer, the table's tail
    goes first, everything in front of the table rides with the last group"""
    import pytest
    from nsr_b200 import configs, ops
    from nsr_b200.parallel import level_group_ranges
    grid = ops.GridSpec(configs.nerf_blender()['geometry']['xyz_encoding_config'])
    n_small = 7168            # the colour network's parameters sit in front of the table's tensor
    first = n_small + 3072    # + the density network's weights inside the table's tensor
    n_total = (first + grid.n_params + 31) // 32 * 32
    r = level_group_ranges(((12, 16), (8, 12), (0, 8)), grid.offset, first, n_total)
    assert r[0][0] + r[0][1] == n_total and r[2][0] == 0
    assert r[2][0] + r[2][1] == r[1][0] and r[1][0] + r[1][1] == r[0][0]          # contiguous, no overlap
    assert sum(c for _, c in r) == n_total
    assert r[0][0] == first + 2 * int(grid.offset[12]) and r[1][0] == first + 2 * int(grid.offset[8])
    assert all(b % 4 == 0 and c % 4 == 0 for b, c in r)
    assert level_group_ranges(((0, 16),), grid.offset, first, n_total) == [(0, n_total)]
    for bad in (((8, 16), (0, 7)), ((0, 8), (8, 16)), ((12, 16), (8, 12)), ((15, 16), (14, 15), (13, 14), (12, 13), (0, 12))):
        with pytest.raises(ValueError):
            level_group_ranges(bad, grid.offset, first, n_total)
