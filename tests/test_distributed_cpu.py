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


def test_merge_records():
    a = np.array([(1, 2, 5, 1, 10), (3, 4, 7, 2, 3)], PAIR)
    b = np.array([(1, 2, 2, 1, 4), (0, 9, 2, 1, 1)], PAIR)
    m = D.merge_pair_records([a, b, np.zeros(0, PAIR)])
    assert m.tolist() == [(0, 9, 2, 1, 1), (1, 2, 2, 2, 14), (3, 4, 7, 2, 3)]
    c = D.merge_coord_records([np.array([(0, 4, 6, 0, 2)], COORD), np.array([(0, 4, 6, 0, 5), (1, 0, 0, 0, 1)], COORD)])
    assert [(r["floor"], r["sx"], r["sy"], r["count"]) for r in c] == [(0, 4, 6, 7), (1, 0, 0, 1)]
    assert len(D.merge_pair_records([])) == 0 and len(D.merge_coord_records([])) == 0


def _worker(rank, world, port, out):
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

    run = D.ShardedRun(Eng(), p["total_timesteps"], block=211).run()
    pairs, coords, dur = run.reduce()
    if rank == 0:
        np.savez(out, pairs=pairs, coords=coords, dur=dur)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_time_block_sharding_equals_single_run(tmp_path):
    import torch.multiprocessing as mp

    from citam_b200 import synthetic as S
    from oracle_util import grid_oracle_for

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "reduced.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    fac, seg_off, segs, p = S.make_config("tiny", seed=2)
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(0, p["total_timesteps"], 1)
    assert z["pairs"].tolist() == g.pairs().astype(PAIR).tolist()
    assert z["coords"].tolist() == g.coords().astype(COORD).tolist()
    assert z["dur"].tolist() == g.agent_durations().astype(np.int64).tolist()
