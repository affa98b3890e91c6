#!/usr/bin/env python
"""BASELINE config 4 in small: a slice of the 1024-replica sweep on this rank's GPU.
Prints replicas/s and aggregate agent-steps/s (itinerary generation on the host included)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from citam_b200 import sweep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
specs = sweep.make_sweep(1024)[:n]
sweep.run_sweep(specs[:2], workers=2)  # warm-up (library load, first allocations)
t0 = time.perf_counter()
res = sweep.run_sweep(specs, workers=workers)
dt = time.perf_counter() - t0
steps = sum(r["agent_steps"] for r in res)
print("replicas %d workers %d seconds %.2f replicas/s %.2f agent-steps/s %.3g contact-steps %d"
      % (len(res), workers, dt, len(res) / dt, steps / dt, sum(r["contact_steps"] for r in res)))
