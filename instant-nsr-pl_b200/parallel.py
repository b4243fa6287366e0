"""Data-parallel gradient exchange: rays shard across GPUs (one process per GPU), every rank holds a full
replica, and the parameter gradients are averaged once per step -- what the reference gets from Lightning DDP
(launch.py:98).  NCCL over NVLink 5 / NVSwitch via torch.distributed; no data-path collective
other than this one exists on the path."""
import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params, world_size=None, group=None):
        self.params = list(params)
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)

    def all_reduce_mean(self):
        """in-place mean of every .grad over the ranks (largest tensor first so NCCL starts on the 50 MB
        hash-table gradient immediately)."""
        if self.world <= 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        grads.sort(key=lambda g: -g.numel())
        works = [dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for g in grads]
        for w in works:
            w.wait()
        inv = 1.0 / self.world
        for g in grads:
            g.mul_(inv)


def shard_rays(rays, rank, world):
    """contiguous shard of a [N, 6] ray batch for this rank (N divisible by world, SURVEY 8e: 65,536 / 8)."""
    n = rays.shape[0] // world
    return rays[rank * n:(rank + 1) * n]
