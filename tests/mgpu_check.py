"""Run under torchrun on >= 2 GPUs: the time-sharded run with the device-side exchange (NCCL
all-reduce of counters and histogram, all-to-all of the hash-partitioned pair table, device merge)
equals the single-process C oracle, table for table."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def check(block, seed, world, rank, local):
    import torch.distributed as dist

    from citam_b200 import distributed as D
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    fac, seg_off, segs, p = S.make_config("config2", seed=seed, n_agents=2500, total_timesteps=2400)
    eng = DeviceEngine(p["n_agents"], p["n_floors"], 6.0, device=local)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    run = D.ShardedRun(eng, p["total_timesteps"], block=block).run()
    sums = run.reduce()
    share = eng.pairs()  # this rank's share, in first-contact order
    assert (D.part_of(share["a1"], share["a2"], world) == rank).all(), "a pair landed on the wrong rank"
    pairs, coords, dur = run.gather_tables()
    err = None
    if rank == 0:
        try:
            _compare(fac, seg_off, segs, p, pairs, coords, dur, sums)
        except AssertionError as e:
            err = str(e) or "mismatch"
    import torch

    flag = torch.tensor([0 if err is None else 1], dtype=torch.int64, device="cuda")
    dist.broadcast(flag, 0)  # every rank learns the verdict: nobody is left waiting in a collective
    eng.close()
    assert int(flag.item()) == 0, err or "rank 0 found a mismatch"
    return len(pairs)


def _compare(fac, seg_off, segs, p, pairs, coords, dur, sums):
    from citam_b200.engine import statistics_from_sums
    from oracle_util import grid_oracle_for, tables_of

    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(0, p["total_timesteps"], 8)
    want = tables_of(g)
    got_pairs = [tuple(int(v) for v in r) for r in pairs.tolist()]
    assert got_pairs == want[0], "pairs differ"
    got_coords = {(int(r["floor"]), int(r["sx"]), int(r["sy"])): int(r["count"]) for r in coords}
    assert got_coords == want[1], "coords differ"
    assert dur.tolist() == want[2], "durations differ"
    # statistics from the partitioned tables == statistics of the whole table
    wp = np.array(want[0], np.int64).reshape(-1, 5)
    nev = np.zeros(p["n_agents"], np.int64)
    np.add.at(nev, wp[:, 0], wp[:, 3])
    np.add.at(nev, wp[:, 1], wp[:, 3])
    has = np.zeros(p["n_agents"], bool)
    has[wp[:, 0]] = True
    has[wp[:, 1]] = True
    want_sums = {"total_duration": int(wp[:, 4].sum()), "n_pairs": len(wp), "n_agents_with_contacts": int(has.sum()),
                 "sum_events_per_agent": int(nev.sum()), "max_events_per_agent": int(nev.max())}
    assert sums == want_sums, (sums, want_sums)
    statistics_from_sums(sums)


def main():
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n1 = check(None, 4, world, rank, local)  # one contiguous range per rank
    n2 = check(173, 5, world, rank, local)   # round-robin time blocks
    if rank == 0:
        print("mgpu ok: world", world, "pairs", n1, n2)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
