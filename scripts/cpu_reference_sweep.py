#!/usr/bin/env python
"""CPU timings of the reference's own formulation beside the GPU numbers.

1. The UNMODIFIED reference step loop (through oracle/refload.py; build container only) on
   examples/basic_example: wall time of `run_simulation_and_save_results`.
2. The faithful Python port (oracle/port.py: the same scipy pdist + Python pair loop) on synthetic
   single-floor crowds of N agents for a window of steps: agent-steps/s per N and the O(N^2) trend.
Writes a markdown table to stdout.  TEST/BENCH INFRASTRUCTURE (imports oracle/).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from citam_b200 import synthetic as S  # noqa: E402


def main():
    print("| what | N | steps | seconds | agent-steps/s |")
    print("|---|---|---|---|---|")
    from oracle import refload

    if refload.available():
        import shutil
        import tempfile

        refload.load()
        from citam.engine.io.input_parser import parse_input_file
        import citam.engine.main as ref_main
        from citam.engine.core.simulation import Simulation

        tmp = tempfile.mkdtemp(prefix="citam_ref_")
        shutil.copytree(os.path.join(refload.REFERENCE_ROOT, "examples/basic_example/foo_facility"), os.path.join(tmp, "foo_facility"))
        shutil.copy(os.path.join(refload.REFERENCE_ROOT, "examples/basic_example/example_sim_inputs.json"), tmp)
        cwd = os.getcwd()
        os.chdir(tmp)
        timing = {}
        orig = Simulation.run_simulation_and_save_results

        def spy(self, *a, **k):
            t0 = time.perf_counter()
            r = orig(self, *a, **k)
            timing["loop"] = time.perf_counter() - t0
            timing["n"], timing["T"] = self.n_agents, self.total_timesteps
            return r

        Simulation.run_simulation_and_save_results = spy
        try:
            np.random.seed(0)
            ref_main.run_simulation(parse_input_file("example_sim_inputs.json"))
        finally:
            Simulation.run_simulation_and_save_results = orig
            os.chdir(cwd)
        print("| unmodified reference, examples/basic_example, step loop | %d | %d | %.2f | %.3g |"
              % (timing["n"], timing["T"], timing["loop"], timing["n"] * timing["T"] / timing["loop"]))
    fac = S.SyntheticFacility(1)
    for n in (1000, 2000, 5000, 10000):
        seg_off, segs = S.make_itineraries(fac, n, 28_800, n_shifts=1, seed=1)
        p = dict(n_agents=n, total_timesteps=28_800, n_shifts=1)
        steps = 12 if n <= 5000 else 4
        r = bench.time_faithful_port(fac, seg_off, segs, p, 6.0, n_sub=n, steps=steps)
        dt = n * steps / r["value"]
        print("| reference formulation (oracle/port.py, pdist + Python pair loop, LUT locations), synthetic 1 floor | %d | %d | %.2f | %.3g |"
              % (n, steps, dt, r["value"]))


if __name__ == "__main__":
    main()
