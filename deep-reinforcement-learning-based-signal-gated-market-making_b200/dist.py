"""Population sharding across the GPUs of one box.

The path shards by contiguous global index: rank r evaluates children [first, first+count) of
the SAME (master, sigma, seed, generation) -- children are regenerated from the counter-based
stream, so no genome ever crosses NVLink.  Per generation the only exchange is an all-gather of
the fitness (f64) and trade-count (i32) slices, after which every rank runs the identical
argmax / tell / validate / select locally.  Works with any torch.distributed backend (NCCL on
the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(pop_size: int, world_size: int, rank: int):
    """Equal contiguous shards of ``stride = ceil(P/R)``; the last ranks may be short or empty."""
    stride = (pop_size + world_size - 1) // world_size
    first = min(rank * stride, pop_size)
    count = min(stride, pop_size - first)
    return first, count, stride


def all_gather_slices(local: torch.Tensor, pop_size: int, stride: int, fill, group=None) -> torch.Tensor:
    """All-gather per-rank slices (padded to ``stride``) and return the first ``pop_size``
    entries in global-index order."""
    world = dist.get_world_size(group)
    padded = torch.full((stride,), fill, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    out = torch.empty(world * stride, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:pop_size]


def device_tensor(ptr: int, n: int, dtype: torch.dtype, device: int) -> torch.Tensor:
    """Zero-copy torch view of a raw device buffer owned by the C library."""
    class _Raw:
        pass
    r = _Raw()
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.float32: "<f4"}[dtype]
    r.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(r, device=f"cuda:{device}")


class ShardedGA:
    """One process per GPU: evaluate the local shard, all-gather fitness/trades into the
    library's gather buffers, select locally (identical on every rank)."""

    def __init__(self, make_ga, pop_size, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.pop_size = pop_size
        self.first, self.count, self.stride = shard_bounds(pop_size, self.world, self.rank)
        self.ga = make_ga((self.first, self.count))
        fs, ts, fa, ta = self.ga.buffers()
        dev = self.ga.device
        cap = self.world * self.stride
        assert cap <= pop_size + 64, "gather buffers hold pop_size + 64 entries"
        self.fit_all = device_tensor(fa, cap, torch.float64, dev)
        self.trd_all = device_tensor(ta, cap, torch.int32, dev)
        # this rank's stride-sized window of the gather buffers (in-place all-gather)
        self.fit_mine = self.fit_all[self.rank * self.stride:(self.rank + 1) * self.stride]
        self.trd_mine = self.trd_all[self.rank * self.stride:(self.rank + 1) * self.stride]

    def generation(self, train, val):
        self.ga.evaluate(train)
        if self.world > 1:
            dist.all_gather_into_tensor(self.fit_all, self.fit_mine, group=self.group)
            dist.all_gather_into_tensor(self.trd_all, self.trd_mine, group=self.group)
        self.ga.select(val)
