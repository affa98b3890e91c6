"""Run under torchrun on >= 2 GPUs: time-block sharding over NCCL equals the single-engine run."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.dirname(HERE), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    from citam_b200 import distributed as D
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine
    from oracle_util import grid_oracle_for, tables_of

    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    fac, seg_off, segs, p = S.make_config("config2", seed=4, n_agents=2500, total_timesteps=2400)
    eng = DeviceEngine(p["n_agents"], p["n_floors"], 6.0, device=local)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    run = D.ShardedRun(eng, p["total_timesteps"], block=173).run()
    # device view of the per-agent seconds, all-reduced over NCCL
    ptr, n = eng.agent_durations_device()

    class Holder:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}

    t = torch.as_tensor(Holder(), device="cuda").clone()  # the engine's own counters stay per-rank
    dist.all_reduce(t)
    dur_nccl = t.cpu().numpy().tolist()
    pairs, coords, dur = run.reduce(into_engine=True)
    if rank == 0:
        g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
        g.run(0, p["total_timesteps"], 8)
        want = tables_of(g)
        got = tables_of(eng)
        assert got[0] == want[0], "pairs differ"
        assert got[1] == want[1], "coords differ"
        assert got[2] == want[2], "durations differ"
        assert dur_nccl == want[2], "NCCL all-reduce of the device counters differs"
        assert dur.tolist() == want[2]
        print("mgpu ok: world", dist.get_world_size(), "pairs", len(got[0]))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
