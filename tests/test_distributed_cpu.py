"""CPU, world_size 2 over gloo: the multi-rank host logic (block plan, ragged all-gather, merges).

Each rank accumulates its own time blocks with the C grid oracle standing in for a GPU engine
(same table schema); the reduced tables must equal a single-process run over the whole day."""
import os
import socket
import sys

import numpy as np
import pytest

from citam_b200 import distributed as D
from citam_b200._lib import COORD, PAIR


def test_plan_blocks_tiles_the_day():
    T, block = 2400, 350
    seen = []
    for r in range(3):
        seen += D.plan_blocks(T, block, 3, r)
    seen.sort()
    assert seen[0][0] == 0 and seen[-1][1] == T
    assert all(a[1] == b[0] for a, b in zip(seen[:-1], seen[1:]))
    assert D.plan_replicas(10, 4, 1) == [1, 5, 9]
    with pytest.raises(ValueError):
        D.plan_blocks(10, 0, 1, 0)


def test_plan_range_tiles_the_day_contiguously():
    for T, G in ((28800, 8), (2401, 3), (5, 8)):
        r = [D.plan_range(T, G, k) for k in range(G)]
        assert r[0][0] == 0 and r[-1][1] == T
        assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_part_of_is_the_device_hash_and_spreads_pairs():
    # known answers of mix64 (the 64-bit finaliser both sides use), computed with Python integers
    def mix(k):
        M = (1 << 64) - 1
        k ^= k >> 33; k = (k * 0xFF51AFD7ED558CCD) & M
        k ^= k >> 33; k = (k * 0xC4CEB9FE1A85EC53) & M
        k ^= k >> 33
        return k
    a1 = np.array([0, 1, 7, 123456, 999_998], np.uint32)
    a2 = np.array([1, 2, 900_000, 123457, 999_999], np.uint32)
    for G in (2, 3, 8):
        want = [(mix((int(x) << 32) | int(y)) >> 40) % G for x, y in zip(a1, a2)]
        assert D.part_of(a1, a2, G).tolist() == want
    rng = np.random.default_rng(0)
    x = rng.integers(0, 10**6, 200_000)
    parts = np.bincount(D.part_of(x, x + 1 + rng.integers(0, 1000, len(x)), 8), minlength=8)
    assert parts.min() > 0.9 * len(x) / 8 and parts.max() < 1.1 * len(x) / 8


def test_merge_records():
    a = np.array([(1, 2, 5, 1, 10), (3, 4, 7, 2, 3)], PAIR)
    b = np.array([(1, 2, 2, 1, 4), (0, 9, 2, 1, 1)], PAIR)
    m = D.merge_pair_records([a, b, np.zeros(0, PAIR)])
    assert m.tolist() == [(0, 9, 2, 1, 1), (1, 2, 2, 2, 14), (3, 4, 7, 2, 3)]
    c = D.merge_coord_records([np.array([(0, 4, 6, 0, 2)], COORD), np.array([(0, 4, 6, 0, 5), (1, 0, 0, 0, 1)], COORD)])
    assert [(r["floor"], r["sx"], r["sy"], r["count"]) for r in c] == [(0, 4, 6, 7), (1, 0, 0, 1)]
    assert len(D.merge_pair_records([])) == 0 and len(D.merge_coord_records([])) == 0


def _worker(rank, world, port, out, block):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    from citam_b200 import synthetic as S
    from oracle_util import grid_oracle_for

    dist.init_process_group("gloo", rank=rank, world_size=world)
    fac, seg_off, segs, p = S.make_config("tiny", seed=2)
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)

    class Eng:  # the engine surface ShardedRun uses
        def run(self, t0, t1):
            g.run(t0, t1, 1)

        pairs = staticmethod(lambda: g.pairs().astype(PAIR))
        coords = staticmethod(lambda: g.coords().astype(COORD))
        agent_durations = staticmethod(lambda: g.agent_durations())

    run = D.ShardedRun(Eng(), p["total_timesteps"], block=block).run()
    share = D.exchange_pairs_host(Eng.pairs())  # this rank's share of the merged pair table
    assert (D.part_of(share["a1"], share["a2"], world) == rank).all()
    shares = D.all_gather_records(share)
    pairs, coords, dur = run.gather_tables()
    if rank == 0:
        np.savez(out, pairs=pairs, coords=coords, dur=dur, shares=D.merge_pair_records(shares),
                 n_shares=sum(len(x) for x in shares))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("block", [211, None])
def test_two_rank_time_sharding_equals_single_run(tmp_path, block):
    """block=211: round-robin time blocks; None: one contiguous range per rank."""
    import torch.multiprocessing as mp

    from citam_b200 import synthetic as S
    from oracle_util import grid_oracle_for

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(2, port, out, block), nprocs=2, join=True)
    z = np.load(out)
    fac, seg_off, segs, p = S.make_config("tiny", seed=2)
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(0, p["total_timesteps"], 1)
    assert z["pairs"].tolist() == g.pairs().astype(PAIR).tolist()
    # the hash-partitioned shares are disjoint and together are the whole table
    assert z["shares"].tolist() == g.pairs().astype(PAIR).tolist() and int(z["n_shares"]) == len(z["pairs"])
    assert z["coords"].tolist() == g.coords().astype(COORD).tolist()
    assert z["dur"].tolist() == g.agent_durations().astype(np.int64).tolist()
