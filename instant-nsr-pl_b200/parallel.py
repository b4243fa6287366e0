"""Data-parallel gradient exchange: rays shard across GPUs (one process per GPU), every rank holds a full
replica, and the parameter gradients are averaged once per step -- what the reference gets from Lightning DDP
(launch.py:98).  NCCL over NVLink 5 / NVSwitch via torch.distributed; no data-path collective
other than this one exists on the path."""
import torch
import torch.distributed as dist


class GradSync:
    """comm_dtype=None: all-reduce the fp32 gradients as they are (what the reference's DDP does).
    comm_dtype=torch.bfloat16: large gradients (the hash table) travel as bf16 -- half the bytes on the wire, 8-bit mantissa;
    an opt-in wire-compression hook, not the default."""

    def __init__(self, params, world_size=None, group=None, comm_dtype=None, compress_min_numel=1 << 20):
        self.params = list(params)
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.comm_dtype = comm_dtype
        self.compress_min_numel = compress_min_numel

    def all_reduce_mean(self):
        """in-place mean of every .grad over the ranks (largest tensor first so NCCL starts on the 50 MB
        hash-table gradient immediately)."""
        if self.world <= 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        grads.sort(key=lambda g: -g.numel())
        inv = 1.0 / self.world
        works, packed = [], []
        for g in grads:
            if self.comm_dtype is not None and g.numel() >= self.compress_min_numel:
                buf = g.to(self.comm_dtype)
                packed.append((g, buf))
                works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            else:
                works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()
        for g, buf in packed:
            g.copy_(buf)
        for g in grads:
            g.mul_(inv)


def shard_rays(rays, rank, world):
    """contiguous shard of a [N, 6] ray batch for this rank (N divisible by world, SURVEY 8e: 65,536 / 8)."""
    n = rays.shape[0] // world
    return rays[rank * n:(rank + 1) * n]
