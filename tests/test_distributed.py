"""N>1 host logic on CPU: world_size-2 gloo.  The MRFP path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mrfp_b200 import dist as D
    from oracle import torch_port as T
    lo, hi = D.shard_bounds(5, world, rank)
    # per-rank statistics only: NP+ on the local shard equals the reference run on that shard
    g = torch.Generator().manual_seed(0)
    feat = torch.relu(torch.randn(5, 6, 8, 8, generator=g))
    alpha = 1 + 0.75 * torch.randn(5, 6, 1, 1, generator=g)
    eps = 0.75 * torch.randn(5, 6, 1, 1, generator=g)
    local = T.np_plus(feat[lo:hi], alpha[lo:hi], eps[lo:hi])
    full = T.np_plus(feat, alpha, eps)
    D.seed_rank_streams(7, rank)
    draw = torch.randn(4)
    gathered = [torch.zeros(4) for _ in range(world)]
    dist.all_gather(gathered, draw)
    ms = D.max_over_ranks(10.0 + rank)
    thr = D.aggregate_throughput(hi - lo, 3, 10.0 + rank)
    q.put((rank, lo, hi, bool(torch.allclose(local, full[lo:hi])), bool(torch.equal(gathered[0], gathered[1])), ms, thr))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, same0, eq0, ms0, thr0), (r1, lo1, hi1, same1, eq1, ms1, thr1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 3, 3, 5)                 # disjoint, complete, sizes differ by <= 1
    assert not same0 and not same1                              # local-batch statistics != global-batch statistics
    assert not eq0                                              # ranks draw different random factors
    assert ms0 == ms1 == 11.0                                   # max over ranks
    assert abs(thr0 - 5 * 3 / 11e-3) < 1e-6 and thr0 == thr1    # all items / slowest rank


def test_local_batch_of_one_is_rejected():
    from mrfp_b200 import dist as D
    with pytest.raises(ValueError):
        D.shard_bounds(3, 2, 1)
    assert D.shard_bounds(16, 8, 7) == (14, 16)
