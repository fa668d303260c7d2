"""Multi-GPU plumbing for the one way this path shards: independent frame-pair problems.

One process per GPU.  A problem is owned by exactly one rank (static round-robin by problem index, the
reference has no notion of ranks); there is NO data-path collective.  torch.distributed (NCCL on the GPU box,
gloo in the CPU tests) is used only for the start barrier and for folding the per-rank timings / counters.
"""
import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(n_problems, world, rank):
    """Problem indices refined by `rank`: p with p % world == rank (every problem exactly once)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_problems, world))


def fold(dist, device, times_ms, counters):
    """max over ranks of every entry of times_ms, sum over ranks of every entry of counters."""
    if dist is None:
        return list(times_ms), list(counters)
    import torch
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(list(counters), dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()], [float(v) for v in c.tolist()]


def throughput(units_all_ranks, max_ms):
    """whole-job rate: units processed by all ranks / slowest rank's time."""
    return units_all_ranks / (max_ms * 1e-3)
