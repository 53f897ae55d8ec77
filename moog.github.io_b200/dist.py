"""Multi-GPU plumbing: envs shard by index, nothing else is shared.

MOOG's Environment.step reads nothing outside its own env
(moog/environment.py:98-126), so N envs split into contiguous index ranges, one
process (rank) per GPU, with no data-path collective.  The only collective is
the sum of the 4 episode statistics, and timings are reported as the maximum
over ranks.  Works on any torch.distributed backend (NCCL on the GPUs, gloo in
the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of env indices owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(seed, rank):
    """Per-rank RNG stream offset (1234 + 1000 * rank)."""
    return int(seed) + 1000 * int(rank)


def reduce_stats(stats):
    """In-place sum of the [4] episode-statistics vector over all ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def max_over_ranks(value, device=None):
    """Scalar max over ranks (device timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
