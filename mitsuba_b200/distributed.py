"""Multi-GPU sharding of the hot path: one process per GPU, sample-index ranges, one film reduce.

The reference shards an image into 32x32 work units handed to worker threads / network nodes and merges the
ImageBlocks under a mutex (src/librender/imageproc.cpp:43-78, renderproc.cpp:142-149, sched_remote.cpp).
Here every rank renders the sample indices [lo, hi) of EVERY pixel into a full-frame (R,G,B,alpha,weight) film
(exactly the single-GPU sample set: Sobol' indices are a pure function of (pixel, sample)), and a single
reduce(SUM) over NCCL/NVLink merges the films -- no other exchange.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional, Tuple


def shard_range(spp: int, rank: int, world: int, lo: int = 0, hi: int = 0) -> Tuple[int, int]:
    """Contiguous sample-index range of `rank` out of [lo, hi) (hi=0 -> spp); ranges tile the whole range."""
    hi = hi if hi > 0 else spp
    n = hi - lo
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    if n < world:
        raise ValueError(f"cannot shard {n} samples per pixel over {world} ranks")
    a = lo + (n * rank) // world
    b = lo + (n * (rank + 1)) // world
    return a, b


def render_sharded(render_fn: Callable, rp, rank: int, world: int, reduce_fn: Optional[Callable] = None):
    """render_fn(rp_shard) -> film tensor/array for the shard; reduce_fn(film) -> reduced film (or None on non-root).

    Returns whatever reduce_fn returns (the full film on the root rank)."""
    lo, hi = shard_range(rp.spp, rank, world, rp.sample_lo, rp.sample_hi)
    shard = dataclasses.replace(rp, sample_lo=lo, sample_hi=hi)
    film = render_fn(shard)
    if reduce_fn is None or world == 1:
        return film
    return reduce_fn(film)


def torch_reduce_sum(film, dst: int = 0, group=None):
    """SUM-reduce a film tensor to `dst` over torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film if dist.get_rank(group) == dst else None
