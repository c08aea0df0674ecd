"""Multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by ciphertext -- every rank owns a
contiguous slice of the batch -- and the only collective is the one-time broadcast of the evaluation keys at setup
(SURVEY 8(e)). Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
from __future__ import annotations


def shard_range(global_batch: int, world: int, rank: int):
    """Contiguous, balanced slice [lo, hi) of a global batch owned by `rank` (first `global_batch % world` ranks
    get one extra ciphertext)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_key(t, src: int = 0):
    """One-time broadcast of an evaluation-key tensor from `src` (in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def max_over_ranks(seconds: float, device=None) -> float:
    """Step time of the job = slowest rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(units_per_rank, seconds: float, device=None) -> float:
    """Whole-job throughput: units processed by all ranks / max-over-ranks time."""
    import torch
    import torch.distributed as dist
    n = float(units_per_rank)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([n], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n = float(t.item())
    return n / max_over_ranks(seconds, device)
