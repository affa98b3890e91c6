"""`citam engine run <inputs.json>` on the GPU engine.

The reference's entry point (citam/cli.py:241-249 -> citam/engine/__init__.py:23-37 ->
citam/engine/main.py:321-453) parses the input file, loads the facility, builds the office
scheduler and calls `Simulation.run_serial`, then writes policy.json and timing.txt.  Parsing,
facility loading and schedule generation are upstream components that stay upstream: `engine run`
imports them from an installed `citam` package and swaps in this package's `Simulation` /
`CloseContactsCalculator` (`build_simulation` + `run_simulation` below follow main.py step by step).

`engine run-precomputed` needs no upstream package: it takes the hot path's inputs directly
(`inputs.npz` with packed itineraries + `geometry.json[.gz]`, the layout of tests/golden/).
"""
from __future__ import annotations

import argparse
import gzip
import json
import logging
import os
import sys
import time

import numpy as np

LOG = logging.getLogger(__name__)


def _upstream():
    """The upstream components that stay upstream: input parsing, facility loading, schedule
    generation."""
    try:
        from citam.engine.constants import CAFETERIA_VISIT, DEFAULT_MEETINGS_POLICY, DEFAULT_SCHEDULING_RULES
        from citam.engine.facility.indoor_facility import Facility
        from citam.engine.io.input_parser import parse_input_file
        from citam.engine.main import load_floorplans
        from citam.engine.schedulers.office_scheduler import OfficeScheduler
    except ImportError as e:  # pragma: no cover - depends on the deployment
        raise SystemExit(
            "citam_b200 engine run needs the upstream `citam` package for input parsing, facility loading and "
            "schedule generation (%s). Use `engine run-precomputed` for pre-built itineraries." % e
        )
    return dict(CAFETERIA_VISIT=CAFETERIA_VISIT, DEFAULT_MEETINGS_POLICY=DEFAULT_MEETINGS_POLICY,
                DEFAULT_SCHEDULING_RULES=DEFAULT_SCHEDULING_RULES, Facility=Facility, parse_input_file=parse_input_file,
                load_floorplans=load_floorplans, OfficeScheduler=OfficeScheduler)


def build_simulation(inputs: dict, **calculator_options):
    """`run_simulation` up to the `Simulation(...)` call (citam/engine/main.py:321-408), with this
    package's `Simulation` and `CloseContactsCalculator` in place of upstream's.  `inputs` is the
    dictionary `parse_input_file` returns and is updated in place exactly as upstream does
    (defaults for close_dining / preassigned_offices / dry_run :345-351, n_agents from the
    occupancy rate :353-364, default policies :366-388)."""
    up = _upstream()
    from .calculator import CloseContactsCalculator
    from .simulation import Simulation

    LOG.info("Loading floorplans...")
    floorplans = up["load_floorplans"](inputs["facility_name"], inputs["floors"])  # main.py:331
    if len(floorplans) == 0:
        raise ValueError("At least one floorplan must be provided.")  # main.py:409-410
    facility = up["Facility"](floorplans, inputs["entrances"], inputs["facility_name"],
                              traffic_policy=inputs["traffic_policy"])
    if "close_dining" not in inputs:
        inputs["close_dining"] = False
    if "preassigned_offices" not in inputs:
        inputs["preassigned_offices"] = None
    if "dry_run" not in inputs:
        inputs["dry_run"] = False
    if inputs["n_agents"] is None and inputs["occupancy_rate"] is None:
        raise ValueError("One of 'n_agents' or 'occupancy_rate' must be specified")
    if inputs["n_agents"] is None:
        if inputs["occupancy_rate"] < 0.0 or inputs["occupancy_rate"] > 1.0:
            raise ValueError("Invalid occupancy rate (must be > 0 & <=1")
        inputs["n_agents"] = round(inputs["occupancy_rate"] * facility.total_offices)
    if inputs["meetings_policy_params"] is None and inputs["create_meetings"]:
        LOG.info("No meetings policy provided. The default policy defined in constants.py will be used.")
        inputs["meetings_policy_params"] = up["DEFAULT_MEETINGS_POLICY"]
    if inputs["scheduling_policy"] is None:
        LOG.info("No scheduling policy provided. The default policy defined in constants.py will be used.")
        inputs["scheduling_policy"] = up["DEFAULT_SCHEDULING_RULES"]
    if inputs["close_dining"] and up["CAFETERIA_VISIT"] in inputs["scheduling_policy"]:
        del inputs["scheduling_policy"][up["CAFETERIA_VISIT"]]
    scheduler = up["OfficeScheduler"](
        facility,
        inputs["total_timesteps"],
        timestep=inputs["timestep"],
        scheduling_rules=inputs["scheduling_policy"],
        meeting_policy=inputs["meetings_policy_params"],
        entry_exit_window=inputs["buffer"],
        preassigned_offices=inputs["preassigned_offices"],
        shifts=inputs["shifts"],
    )
    # upstream builds the calculator with the DEFAULT distance: the parsed
    # inputs["contact_distance"] is never consumed (main.py:405)
    return Simulation(
        facility,
        inputs["total_timesteps"],
        inputs["n_agents"],
        calculators=[CloseContactsCalculator(facility, **calculator_options)],
        scheduler=scheduler,
        dry_run=inputs["dry_run"],
    )


def write_policy(inputs: dict, work_directory: str) -> None:
    """policy.json as upstream writes it (main.py:420-433; consumed by citam/api/parser.py:263-270)."""
    policy = {}
    if "output_directory" in inputs:
        del inputs["output_directory"]
    policy["facility_name"] = inputs["facility_name"]
    policy["general"] = {key: val for key, val in inputs.items() if key in ["total_timesteps", "n_agents", "dry_run"]}
    policy["meetings"] = inputs["meetings_policy_params"]
    policy["scheduling"] = inputs["scheduling_policy"]
    policy["traffic"] = inputs["traffic_policy"]
    with open(os.path.join(work_directory, "policy.json"), "w") as outfile:
        json.dump(policy, outfile)


def run_simulation(inputs: dict, **calculator_options) -> None:
    """citam/engine/main.py:321-453 on the GPU engine: same inputs, same files."""
    start_time = time.time()
    upload_results = inputs["upload_results"]
    upload_location = inputs["upload_location"]
    my_model = build_simulation(inputs, **calculator_options)
    work_directory = inputs["output_directory"]
    LOG.info("Running simulation...")
    my_model.run_serial(workdir=work_directory, sim_name=inputs["simulation_name"], run_name=inputs["run_name"])
    write_policy(inputs, work_directory)
    if upload_results and upload_location is not None:  # main.py:435-444
        LOG.info("Uploading results to server...")
        cwd = os.getcwd()
        foldername = os.path.basename(cwd)
        os.chdir("..")
        try:
            os.system("s3cmd sync " + foldername + " " + upload_location)
        except Exception as e:  # pragma: no cover
            LOG.error("Failed to upload results to S3. Error: %s", e)
        os.chdir(foldername)
    total_time = time.time() - start_time
    with open(work_directory + "timing.txt", "w") as outfile:  # main.py:446-449
        outfile.write(str(total_time))
    LOG.info("Results were saved here: %s", work_directory)
    LOG.info("Elapsed time: %s sec", total_time)


def engine_run(**kwargs) -> None:
    """`citam engine run <input_file>` (citam/cli.py:241-249 -> citam/engine/__init__.py:23-37)."""
    LOG.info("Loading input file: %s", kwargs["input_file"])
    input_dict = _upstream()["parse_input_file"](kwargs["input_file"])
    LOG.info("Starting simulation")
    run_simulation(input_dict)
    LOG.info("Simulation has completed")


def _run(sim, workdir, sim_name, run_name):
    t0 = time.time()
    os.makedirs(workdir, exist_ok=True)
    sim.run_serial(workdir, sim_name, run_name)
    with open(os.path.join(workdir, "timing.txt"), "w") as f:
        f.write(str(time.time() - t0))


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
        engine_run(input_file=a.input_file)
    else:
        engine_run_precomputed(a.inputs_npz, a.geometry, a.workdir, full_fidelity=not a.summary_only)


if __name__ == "__main__":
    main()
