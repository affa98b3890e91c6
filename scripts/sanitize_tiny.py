#!/usr/bin/env python
"""The `tiny` workload through every device path, small enough to run under compute-sanitizer:
  compute-sanitizer --tool memcheck  python scripts/sanitize_tiny.py
  compute-sanitizer --tool racecheck python scripts/sanitize_tiny.py
(memcheck / racecheck evidence; logs are kept in profiles/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from citam_b200 import _lib as L  # noqa: E402
from citam_b200 import synthetic as S  # noqa: E402
from citam_b200.engine import DeviceEngine  # noqa: E402


def main():
    fac, seg_off, segs, p = S.make_config("tiny", seed=1)
    N, T = p["n_agents"], 600
    # block path: chunks of 37 steps, contact runs crossing chunk borders, side-stream accumulation
    eng = DeviceEngine(N, p["n_floors"], 6.0, steps_per_chunk=37)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(300, 300 + T)
    pairs, coords, dur = eng.pairs(), eng.coords(), eng.agent_durations()
    sums = eng.statistics_sums()
    assert sums["n_pairs"] == len(pairs) and int(dur.sum()) == 2 * int(pairs["duration"].sum())
    # exchange path on one device: partition, clear, merge
    import ctypes as C

    lib = L.load()
    rt = C.CDLL("libcudart.so")
    buf = C.c_void_p()
    assert rt.cudaMalloc(C.byref(buf), max(1, len(pairs)) * 24) == 0
    counts = eng.partition_pairs_device(buf.value, len(pairs), 3)
    eng.clear_pairs()
    eng.merge_pairs_device(buf.value, int(counts.sum()))
    assert eng.pairs().tolist() == pairs.tolist()
    rt.cudaFree(buf)
    states = eng.states_range(300, 320)
    eng.close()
    # full-fidelity path: contact log, every step evaluated, device sort of the log
    eng = DeviceEngine(N, p["n_floors"], 6.0, log_capacity=1 << 20, steps_per_chunk=16)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(300, 420)
    log = eng.contact_log()
    assert (np.lexsort((log["a2"], log["a1"], log["step"])) == np.arange(len(log))).all()
    eng.close()
    # plugin path: explicit states, one step per call
    eng = DeviceEngine(N, p["n_floors"], 6.0, hashed_coords=True, coord_capacity=1 << 16)
    eng.load_geometry(fac.geometry)
    for k in range(6):
        eng.step_states(300 + k, states[k])
    eng.coords()
    eng.close()
    print("sanitize_tiny ok:", len(pairs), "pairs,", len(log), "log records")


if __name__ == "__main__":
    main()
