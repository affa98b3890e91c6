"""`citam engine run <inputs.json>` on the GPU engine.

The reference's entry point (citam/cli.py:241-249 -> citam/engine/__init__.py:23-37 ->
citam/engine/main.py:321-453) parses the input file, loads the facility, builds the office
scheduler and calls `Simulation.run_serial`.  Parsing, facility loading and schedule generation
are upstream components that stay upstream: `engine run` imports them from an installed `citam`
package and swaps in this package's `Simulation` / `CloseContactsCalculator`.

`engine run-precomputed` needs no upstream package: it takes the hot path's inputs directly
(`inputs.npz` with packed itineraries + `geometry.json[.gz]`, the layout of tests/golden/).
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import sys
import time

import numpy as np


def _run(sim, workdir, sim_name, run_name, policy=None):
    t0 = time.time()
    os.makedirs(workdir, exist_ok=True)
    sim.run_serial(workdir, sim_name, run_name)
    if policy is not None:
        with open(os.path.join(workdir, "policy.json"), "w") as f:  # main.py:420-433
            json.dump(policy, f, indent=4)
    with open(os.path.join(workdir, "timing.txt"), "w") as f:  # main.py:446-449
        f.write(str(time.time() - t0))


def engine_run(input_file: str, **kwargs) -> None:
    try:
        from citam.engine.facility.indoor_facility import Facility
        from citam.engine.io.input_parser import parse_input_file
        from citam.engine.main import load_floorplans
        from citam.engine.schedulers.office_scheduler import OfficeScheduler
    except ImportError as e:  # pragma: no cover - depends on the deployment
        raise SystemExit(
            "citam_b200 engine run needs the upstream `citam` package for input parsing, facility loading and "
            "schedule generation (%s). Use `engine run-precomputed` for pre-built itineraries." % e
        )
    from .calculator import CloseContactsCalculator
    from .simulation import Simulation

    inputs = parse_input_file(input_file)
    floorplans = load_floorplans(inputs["floors"], inputs["facility_name"], inputs.get("user_scale"))
    facility = Facility(floorplans, inputs["entrances"], inputs["facility_name"], traffic_policy=inputs.get("traffic_policy"))
    scheduler = OfficeScheduler(
        facility, inputs["daylength"], shifts=inputs.get("shifts"), buffer=inputs.get("buffer", 300),
        preassigned_offices=inputs.get("preassigned_offices"), scheduling_policy=inputs.get("scheduling_policy"),
        meetings_policy_params=inputs.get("meetings_policy_params"), create_meetings=inputs.get("create_meetings", True),
    ) if "shifts" in inputs else OfficeScheduler(facility, inputs["daylength"])
    # upstream builds the calculator with the default distance; the parsed one is not consumed (main.py:405)
    sim = Simulation(facility, inputs["daylength"], inputs["n_agents"],
                     calculators=[CloseContactsCalculator(facility)], scheduler=scheduler)
    _run(sim, inputs.get("output_directory", os.getcwd() + "/"), inputs["simulation_name"], inputs["run_name"])


def load_precomputed(inputs_npz: str, geometry_file: str):
    z = np.load(inputs_npz)
    opener = gzip.open if geometry_file.endswith(".gz") else open
    with opener(geometry_file, "rt") as f:
        geometry = json.load(f)
    return z, geometry


def engine_run_precomputed(inputs_npz: str, geometry_file: str, workdir: str, sim_name="sim", run_name="run",
                           full_fidelity=True) -> None:
    from . import itinerary as itin
    from .calculator import CloseContactsCalculator
    from .facility import GeometryFacility
    from .scheduler import SegmentScheduler
    from .simulation import Simulation

    z, geometry = load_precomputed(inputs_npz, geometry_file)
    seg_off, segs = itin.encode_packed(z["itin_off"], z["itin_x"], z["itin_y"], z["itin_floor"])
    facility = GeometryFacility(geometry, "precomputed")
    T, D = int(z["total_timesteps"]), float(z["contact_distance"])
    n = len(seg_off) - 1
    sim = Simulation(facility, T, n, calculators=[CloseContactsCalculator(facility, D, full_fidelity=full_fidelity)],
                     scheduler=SegmentScheduler(facility, T, seg_off, segs))
    _run(sim, workdir, sim_name, run_name)


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(prog="citam_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    eng = sub.add_parser("engine").add_subparsers(dest="verb", required=True)
    r = eng.add_parser("run", help="run a simulation from a CITAM input file")
    r.add_argument("input_file")
    p = eng.add_parser("run-precomputed", help="run the step loop on pre-built itineraries")
    p.add_argument("inputs_npz")
    p.add_argument("geometry")
    p.add_argument("workdir")
    p.add_argument("--summary-only", action="store_true", help="skip the per-step files")
    a = ap.parse_args(argv)
    if a.verb == "run":
        engine_run(a.input_file)
    else:
        engine_run_precomputed(a.inputs_npz, a.geometry, a.workdir, full_fidelity=not a.summary_only)


if __name__ == "__main__":
    main()
