"""Multi-GPU host logic of the MRFP path: batch sharding, per-rank RNG streams, max-over-ranks timing.

The path shards by batch with NO collective (every NP+ / HRFP-BN statistic is local to the replica's batch,
as under the reference's nn.DataParallel scatter, main.py:824); the only cross-rank operations here are the
ones a benchmark or a trainer needs around it.
"""
import random

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, world: int, rank: int):
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    if hi - lo < 2:
        # torch.std over a single sample is NaN (deepv3.py:272): NP+ needs a local batch of at least 2
        raise ValueError(f"per-rank batch {hi - lo} < 2: NP+ statistics are over the LOCAL batch and need >= 2 samples")
    return lo, hi


def seed_rank_streams(base_seed: int, rank: int):
    """The reference re-seeds every module to 0 at import (deepv3.py:39-44: random, numpy, torch), which would make all
    ranks draw identical gates / alpha / eps / HRFP weights; give every rank its own streams instead.  The three MRFP
    gates p, p2, p3 come from Python's `random.random()` (deepv3.py:281-283), so that generator is seeded too."""
    seed = base_seed + 1000003 * rank
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def max_over_ranks(value: float, device="cpu") -> float:
    """Device-timed durations are reported as the maximum over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(local_items: int, steps: int, local_ms: float, device="cpu") -> float:
    """Whole-job items/s: sum of the items of all ranks over the slowest rank's time."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    n = torch.tensor([float(local_items)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()) * steps / (max_over_ranks(local_ms, device) * 1e-3)
