"""Where does the ordered pair export (finalize) spend its time?  config 3 whole day, then two exports."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from citam_b200 import synthetic as S
from citam_b200.engine import DeviceEngine

cap_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 30
fac, seg_off, segs, p = S.make_config("config3", seed=0)
N, T = p["n_agents"], p["total_timesteps"]
eng = DeviceEngine(N, p["n_floors"], 6.0, pair_capacity=1 << cap_log2, timing=True)
eng.load_geometry(fac.geometry); eng.load_segments(seg_off, segs)
eng.run(0, T)
n = eng.pair_count()
buf = torch.empty((n, 3), dtype=torch.int64, device="cuda")
for k in range(3):
    t0 = eng.timings().get("finalize", {"ms": 0.0})["ms"]
    torch.cuda.synchronize(); w0 = time.perf_counter()
    got = eng.export_pairs_device(buf.data_ptr(), n)
    torch.cuda.synchronize(); w1 = time.perf_counter()
    t1 = eng.timings().get("finalize", {"ms": 0.0})["ms"]
    print("cap 2^%d export %d: pairs %d wall %.1f ms, finalize-kind kernels %.1f ms" % (cap_log2, k, got, 1e3 * (w1 - w0), t1 - t0), flush=True)
# the bench's sequence: pinned host copies of the store, upload + reset + day, then an export
so_p = torch.from_numpy(np.ascontiguousarray(seg_off, np.int64)).pin_memory()
sg_p = torch.from_numpy(np.ascontiguousarray(segs).view(np.uint8)).pin_memory()
out_dur = torch.empty(N, dtype=torch.int64).pin_memory()
for k in range(2):
    eng.load_segments_raw(so_p.data_ptr(), sg_p.data_ptr(), len(segs))
    eng.reset(); eng.run(0, T); eng.statistics_sums(); eng.agent_durations_into(out_dur.data_ptr())
    del buf
    buf = torch.empty((n, 3), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize(); w0 = time.perf_counter()
    got = eng.export_pairs_device(buf.data_ptr(), n)
    torch.cuda.synchronize(); w1 = time.perf_counter()
    print("after an e2e day %d: export wall %.1f ms" % (k, 1e3 * (w1 - w0)), flush=True)
eng.close()
