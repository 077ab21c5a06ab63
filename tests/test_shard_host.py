"""CPU: host-side logic of the multi-GPU particle sharding (genparticlefilters.jl_b200/sharded.py) under the
gloo backend at world_size 2, plus the pure range arithmetic mirrored from k_shard_ranges."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exchange_plan_properties():
    from genpf_b200.sharded import cross_shard_fraction, exchange_plan
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        n_local = 4096
        n_total = world * n_local
        for _ in range(50):
            cuts = np.sort(rng.integers(0, n_total + 1, world - 1)) if world > 1 else np.array([], dtype=int)
            oend = list(cuts) + [int(rng.integers(n_total - 5, n_total + 1))]
            if world > 2 and rng.random() < 0.3:  # ulp-level inversion between neighbouring shards
                i = rng.integers(0, world - 2)
                oend[i + 1] = max(0, oend[i] - 1)
            ranges = exchange_plan(oend, n_total)
            assert ranges[0][0] == 0 and ranges[-1][1] == n_total
            assert all(b <= e for b, e in ranges)
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))  # exact cover, no overlap
            f = cross_shard_fraction(ranges, n_local)
            assert 0.0 <= f <= 1.0
    # balanced shards: nothing crosses
    assert cross_shard_fraction(exchange_plan([4096, 8192], 8192), 4096) == 0.0
    # all mass in shard 0: half of the offspring cross
    assert cross_shard_fraction(exchange_plan([8192, 8192], 8192), 4096) == 0.5


def test_closing_counts_from_totals_match_the_unsharded_ancestors():
    """The two-synchronisation-point protocol exchanges only per-shard totals; every rank derives all closing counts
    from them (sharded.closing_counts_from_totals mirrors the device code).  Against the CPU oracle's stratified
    ancestors of the WHOLE population: the derived ranges are an exact cover, and the outputs of range g are parented
    by shard g (a boundary output may differ only at a cumulative-sum tie)."""
    from genpf_b200.sharded import closing_counts_from_totals, exchange_plan
    from oracle import oracle as orc
    rng = np.random.default_rng(5)
    for world, n_local, tilt in ((2, 4096, 0.0), (4, 2048, 0.7), (8, 2048, 0.35), (8, 2048, 0.0)):
        n = world * n_local
        lw = rng.normal(0, 1.5, n) + tilt * np.repeat(np.arange(world), n_local)  # imbalanced shards when tilt > 0
        r = rng.random(n)
        totals = []
        for g in range(world):
            v = lw[g * n_local:(g + 1) * n_local]
            totals.append((v.max(), np.exp(v - v.max()).sum()))
        oend = closing_counts_from_totals(totals, n, r)
        assert all(oend[i] <= oend[i + 1] for i in range(world - 1)) and oend[-1] == n
        ranges = exchange_plan(oend, n)
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        p_ref = orc.resample("stratified", lw, r)[0]
        owner_of_parent = p_ref // n_local
        bad = 0
        for g, (b, e) in enumerate(ranges):
            bad += int(np.sum(owner_of_parent[b:e] != g))
        assert bad <= world - 1, bad  # at most one tie per shard boundary (observed: 0)
        if tilt > 0:
            sizes = [e - b for b, e in ranges]
            assert sizes[-1] > sizes[0]  # mass, and with it offspring, moved towards the tilted shards


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from genpf_b200.sharded import ShardExchange, exchange_plan
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = ShardExchange(torch.device("cpu"))
    ex.stats_local[:] = torch.tensor([1.0 + rank, 10.0 * (rank + 1), 100.0 * (rank + 1)], dtype=torch.float64)
    ex.gather_stats()
    ex.oend_local[0] = 3000 if rank == 0 else 8192
    ex.gather_oend()
    ex.barrier()
    blobs = ex.all_gather_bytes(bytes([rank]) * 64)
    ranges = exchange_plan(ex.oend_all.numpy(), 8192)
    # global logsumexp pieces combine like k_shard_combine
    a = ex.stats_all.numpy().reshape(world, 3)
    M = a[:, 0].max()
    S = float((a[:, 1] * np.exp(a[:, 0] - M)).sum())
    q.put((rank, ex.stats_all.tolist(), ex.oend_all.tolist(), ranges, [b[0] for b in blobs], M, S))
    dist.destroy_process_group()


def test_shard_exchange_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29641, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, stats_all, oend_all, ranges, blobs, M, S in res:
        assert stats_all == [1.0, 10.0, 100.0, 2.0, 20.0, 200.0]
        assert oend_all == [3000, 8192]
        assert ranges == [(0, 3000), (3000, 8192)]
        assert blobs == [0, 1]
        assert M == 2.0 and S == pytest.approx(20.0 + 10.0 * np.exp(-1.0))
