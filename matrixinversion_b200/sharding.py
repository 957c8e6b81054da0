"""Multi-GPU partitioning of the batch.

The path shards by independent units: GPU g of G owns the contiguous slice
[g*ceil(B/G), min(B, (g+1)*ceil(B/G))) of the batch.  No data-path collective exists;
the only exchange is an optional reduction of two scalars (bad-matrix count, max residual).
The reference itself is single-GPU (one cudaMalloc on the default device,
templated/luBatchedInplace.cu:63).
"""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the batch owned by `rank`; empty (lo == hi) when batch < world*..."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad shard arguments")
    per = -(-batch // world)
    lo = min(batch, rank * per)
    hi = min(batch, lo + per)
    return lo, hi


def reduce_verdict(n_bad: int, max_dev: float, group=None):
    """Sum of incorrect inversions and max |r - delta| over all ranks (torch.distributed)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return n_bad, max_dev
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    bad = torch.tensor([float(n_bad)], dtype=torch.float64, device=dev)
    mx = torch.tensor([float(max_dev)], dtype=torch.float64, device=dev)
    dist.all_reduce(bad, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return int(bad.item()), float(mx.item())
