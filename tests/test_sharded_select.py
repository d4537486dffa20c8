"""CPU tests of the multi-GPU host logic: contiguous sharding by bases and the
split-phase normalise/select protocol with an all-reduce between phases, on gloo with world size 2.
The per-rank compute is the numpy restatement in tests/numpy_phases.py; the expected result is the
oracle's single-process sort + prefix walk on the whole read set."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from filtlong_b200 import sharding
from oracle import oracle as orc
from tests import parity, util
from tests.numpy_phases import NumpyPhases


def make_rows(seed, n=400, ties=False):
    rng = np.random.default_rng(seed)
    genome = util.rand_seq(rng, 30000)
    reads = [(s, q) for _, s, q in util.long_reads(rng, genome, n, max_len=5000)]
    if ties:                                # exact duplicates -> tie classes at the cut-off
        reads = reads[: n // 4] * 4
    return reads


def oracle_rows(reads, opts):
    p = orc.make_params(**opts)
    return orc.finalize(orc.score(reads, p, None), p), p


def test_shard_by_bases_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    L = rng.integers(100, 50000, size=1000)
    for world in (1, 2, 3, 8):
        cuts = sharding.shard_by_bases(L, world)
        assert len(cuts) == world and cuts[0][0] == 0 and cuts[-1][1] == len(L)
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        per = [int(L[lo:hi].sum()) for lo, hi in cuts]
        assert max(per) - min(per) <= 2 * L.max()
    assert sharding.shard_by_bases([], 4) == [(0, 0)] * 4


@pytest.mark.parametrize("seed,opts,ties", [
    (1, dict(target_bases=300000), False),
    (2, dict(keep_percent=60.0, min_length=400), False),
    (3, dict(keep_percent=35.0), True),
    (4, dict(target_bases=10 ** 12), False),
    (5, dict(target_bases=500, min_length=10 ** 6), False),
    (6, dict(keep_percent=50.0, length_weight=2.0, window_q_weight=0.0), True),
])
def test_numpy_phases_single_rank_matches_oracle(seed, opts, ties):
    reads = make_rows(seed, ties=ties)
    sc, p = oracle_rows(reads, opts)
    rows = sc.rows
    ph = NumpyPhases([r.mean_q for r in rows], [r.window_q for r in rows], [r.length for r in rows],
                     [r.passed for r in rows], p)
    buf = sharding.Buffers(torch, "cpu", 1)
    s = sharding.sharded_finalize(ph, None, buf, 0, 1, sc.total_bases)
    assert s.status == sc.summary.status
    parity.check_selection([int(x) for x in ph.pfinal], [r.passed_final for r in rows],
                           [r.final_score for r in rows], [r.length for r in rows])
    if s.status == 3:
        assert s.keeping == sc.summary.keeping and s.target == sc.summary.target


def _worker(rank, world, port, seed, opts, ties, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = make_rows(seed, ties=ties)
    sc, p = oracle_rows(reads, opts)            # every rank scores everything here only to get the rows;
    rows = sc.rows                              # each rank then keeps just its shard
    lo, hi = sharding.shard_by_bases([r.length for r in rows], world)[rank]
    mine = rows[lo:hi]
    ph = NumpyPhases([r.mean_q for r in mine], [r.window_q for r in mine], [r.length for r in mine],
                     [r.passed for r in mine], p)
    buf = sharding.Buffers(torch, "cpu", world)
    s = sharding.sharded_finalize(ph, dist, buf, rank, world, sc.total_bases)
    out[rank] = (lo, hi, [int(x) for x in ph.pfinal], s.status, s.keeping, s.target)
    dist.destroy_process_group()


@pytest.mark.parametrize("seed,opts,ties", [
    (11, dict(target_bases=250000), False),
    (12, dict(keep_percent=45.0), True),
    (13, dict(keep_percent=80.0, min_mean_q=85.0), False),
])
def test_two_rank_gloo_protocol_matches_single_process_oracle(seed, opts, ties):
    world = 2
    port = 29500 + (os.getpid() + seed) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, seed, opts, ties, out), nprocs=world, join=True)
    reads = make_rows(seed, ties=ties)
    sc, _ = oracle_rows(reads, opts)
    got = [0] * len(sc.rows)
    for rank in range(world):
        lo, hi, flags, status, keeping, target = out[rank]
        got[lo:hi] = flags
        assert status == sc.summary.status
        if status == 3:
            assert keeping == sc.summary.keeping and target == sc.summary.target
    parity.check_selection(got, [r.passed_final for r in sc.rows], [r.final_score for r in sc.rows],
                           [r.length for r in sc.rows])
