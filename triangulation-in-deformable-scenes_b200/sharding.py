"""Multi-GPU plumbing for the two ways this path shards: independent frame-pair problems (BASELINE.json
configs[4], below) and ONE frame pair split by points over the GPUs (ShardedPair at the end of this file: the data plane
is inside the library -- kernels writing peer-mapped memory over NVLink; torch.distributed only carries the 64-byte
memory handles once, at set-up).

One process per GPU.  A problem is owned by exactly one rank (static round-robin by problem index: the reference has no
notion of ranks, and on one GPU the device-side queue of lm_batch_kernel balances the load between clusters); there is
NO data-path collective -- a frame pair never needs another pair's data.  torch.distributed (NCCL on the GPU box, gloo in
the CPU tests) carries three things only: the start barrier, the fold of the per-rank timings / counters, and ONE gather
of the small per-problem results at the end ("one ncclGather / host gather of results").

  shard / owner            problem index <-> rank
  ShardedBatch             a rank's share of a batch: uploads its pairs into a dsc_batch, refines them with one launch,
                           hands back per-problem results in GLOBAL problem order after gather_by_problem
  fold / throughput        max-over-ranks timing, whole-job rate
"""
import os

import numpy as np


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(n_problems, world, rank):
    """Problem indices refined by `rank`: p with p % world == rank (every problem exactly once)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_problems, world))


def owner(p, world):
    """(rank, position inside that rank's shard) of problem p"""
    return p % world, p // world


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


def gather_by_problem(dist, device, n_problems, world, rank, local_rows):
    """The one collective of the batched path: every rank contributes a [len(shard), C] float64 block of per-problem
    results (row k = its k-th problem, i.e. global problem rank + k * world); every rank gets the [n_problems, C] table
    in global problem order.  One all_gather of equally sized (zero-padded) blocks."""
    local_rows = np.ascontiguousarray(local_rows, np.float64)
    if local_rows.ndim == 1:
        local_rows = local_rows[:, None]
    mine = shard(n_problems, world, rank)
    if local_rows.shape[0] != len(mine):
        raise ValueError(f"rank {rank} owns {len(mine)} problems, got {local_rows.shape[0]} rows")
    c = local_rows.shape[1]
    if dist is None or world == 1:
        return local_rows.copy()
    import torch
    per = (n_problems + world - 1) // world
    block = torch.zeros((per, c), dtype=torch.float64, device=device)
    if len(mine):
        block[:len(mine)] = torch.from_numpy(local_rows).to(device)
    table = torch.empty((world, per, c), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(table.view(world * per, c), block)
    t = table.cpu().numpy()
    out = np.empty((n_problems, c), np.float64)
    for p in range(n_problems):
        r, k = owner(p, world)
        out[p] = t[r, k]
    return out


class ShardedBatch:
    """A rank's share of a batch of independent frame pairs.

    problems_of(p) -> problem dict (Batch.upload layout) is called for the problems this rank owns only, so host-side
    synthesis / loading is sharded too."""

    def __init__(self, pkg, device, n_problems, world, rank):
        self.pkg, self.n_problems, self.world, self.rank = pkg, n_problems, world, rank
        self.mine = shard(n_problems, world, rank)
        self.batch = pkg.Batch(device)
        self.problems = []

    def build(self, problem_of):
        self.problems = [problem_of(p) for p in self.mine]
        return self

    def upload(self):
        return self.batch.upload(self.problems)

    def optimize(self, weights, n_iters):
        return self.batch.optimize(weights, n_iters)

    def result_rows(self, stats):
        """per-problem rows of this rank for gather_by_problem: final chi2, LM iterations, trials, PCG iterations"""
        return np.array([[s.final_chi2, s.iterations, s.total_trials, s.total_pcg_iters] for s in stats], np.float64).reshape(len(stats), 4)

    def close(self):
        self.batch.close()


# ---------------------------------------------------------------------------------------------- one pair over several GPUs
def exchange_handles(dist, device, handle, world, rank):
    """all-gather of the ranks' 64-byte arena handles -> [world][64] uint8 in rank order (the only use of torch.distributed
    by the point-sharded path)."""
    handle = np.ascontiguousarray(handle, np.uint8).reshape(-1)
    if dist is None or world == 1:
        return handle[None, :].copy()
    import torch
    mine = torch.from_numpy(handle.copy()).to(device)
    table = torch.empty((world, handle.size), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(table.view(-1), mine)
    out = table.cpu().numpy()
    if not np.array_equal(out[rank], handle):
        raise RuntimeError("handle exchange: own handle came back changed")
    return out


class ShardedPair:
    """One rank of a frame pair split by points (dsc.h: dsc_shard_*).  Every rank builds it with the same problem; after
    attach() the Context is used exactly like a single-GPU one (problem_upload, set_graph, compute_rotations, optimize,
    download): every rank gets the same records and, after download, the whole refined pair."""

    def __init__(self, pkg, device, max_points, dist=None, world=1, rank=0, torch_device=None):
        self.pkg, self.world, self.rank = pkg, world, rank
        self.ctx = pkg.Context(device)
        handle = self.ctx.shard_init(rank, world, max_points)
        self.handles = exchange_handles(dist, torch_device or f"cuda:{device}", handle, world, rank)
        self.ctx.shard_attach(self.handles)
        if dist is not None and world > 1:
            dist.barrier()                                  # every arena is mapped everywhere before anyone writes

    def close(self, dist=None):
        if dist is not None and self.world > 1:
            dist.barrier()                                  # nobody writes into an arena that is about to be freed
        self.ctx.close()
