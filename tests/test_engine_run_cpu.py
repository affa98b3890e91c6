"""`citam engine run <inputs.json>` up to the step loop, on the CPU box, against the LIVE reference.

`citam_b200.cli.build_simulation` follows `run_simulation` (citam/engine/main.py:321-408) with the
upstream parser / facility / scheduler and this package's Simulation + CloseContactsCalculator.
Here the upstream package is the unmodified tree under /root/reference (oracle/refload.py; skipped
where it is absent, i.e. on the GPU box): under np.random.seed(0) the agents it creates must carry
exactly the itineraries the reference's own run produced (tests/golden/basic_seed0/inputs.npz, written
by oracle/make_golden.py from `run_simulation`), and policy.json must be byte-identical to the
reference's.  The step loop on those very inputs is then the GPU parity tests' business
(tests/test_gpu_surface.py compares all seven output files byte for byte)."""
import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from golden_util import expected, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, os, shutil, sys
import numpy as np
sys.path.insert(0, %(root)r)
from oracle import refload
refload.load()
REF = refload.REFERENCE_ROOT
tmp = %(tmp)r
shutil.copytree(os.path.join(REF, "examples/basic_example/foo_facility"), os.path.join(tmp, "foo_facility"))
shutil.copy(os.path.join(REF, "examples/basic_example/example_sim_inputs.json"), tmp)
os.chdir(tmp)
from citam.engine.io.input_parser import parse_input_file
from citam_b200 import cli, itinerary
from citam_b200.facility import geometry_from_facility
np.random.seed(0)
inputs = parse_input_file("example_sim_inputs.json")
sim = cli.build_simulation(inputs)
sim.create_sim_hash()
sim.save_manifest(tmp, inputs["simulation_name"], inputs["run_name"])
sim.create_agents()
off, x, y, fl = itinerary.pack([a.schedule.itinerary for a in sim.agents.values()])
np.savez(os.path.join(tmp, "got.npz"), off=off, x=x, y=y, fl=fl, T=sim.total_timesteps, n=sim.n_agents,
         D=sim.calculators[0].contact_distance, dry=sim.dry_run, ids=np.array(list(sim.agents.keys())))
with open(os.path.join(tmp, "geometry.json"), "w") as f:
    json.dump(geometry_from_facility(sim.facility), f)
cli.write_policy(inputs, tmp)
"""


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    from oracle import refload

    if not refload.available():
        pytest.skip("reference tree not present (GPU box)")
    tmp = str(tmp_path_factory.mktemp("engine_run"))
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "tmp": tmp}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return tmp


def test_build_simulation_creates_the_reference_runs_agents(built):
    case = load_case("basic_seed0")
    z = np.load(os.path.join(built, "got.npz"))
    off, x, y, fl = case["packed"]
    assert int(z["n"]) == len(off) - 1 == 20 and int(z["T"]) == case["T"] == 1800
    assert z["ids"].tolist() == list(range(20))
    assert (z["off"] == off).all()
    assert (z["fl"] == fl).all()
    active = fl >= 0
    assert (z["x"][active] == x[active]).all() and (z["y"][active] == y[active]).all()
    # the calculator keeps the DEFAULT distance: the input file's contact_distance is not consumed (main.py:405)
    assert float(z["D"]) == 6.0 == case["D"] and not bool(z["dry"])


def test_build_simulation_loads_the_same_facility(built):
    case = load_case("basic_seed0")
    with open(os.path.join(built, "geometry.json")) as f:
        got = json.load(f)
    assert len(got["floors"]) == len(case["geometry"]["floors"]) == 1
    g, w = got["floors"][0], case["geometry"]["floors"][0]
    assert g["spaces"] == w["spaces"]  # 214 spaces in the reference's order, after unreachable rooms were removed
    assert sorted(sorted(e) for e in g["hallway_edges"]) == sorted(sorted(e) for e in w["hallway_edges"])


def test_policy_and_manifest_files(built):
    case = load_case("basic_seed0")
    with open(os.path.join(built, "policy.json"), "rb") as f:
        assert f.read() == expected(case, "policy.json")
    with open(os.path.join(built, "manifest.json")) as f:
        m = json.load(f)
    for k, v in {"NumberOfAgents": 20, "TotalTimesteps": 1800, "FacilityName": "foo_facility", "SimulationName": "test_sim_1",
                 "RunName": "test_run_1", "NumberOfFloors": 1, "TrajectoryFile": "trajectory.txt",
                 "Floors": [{"name": "0", "directory": "floor_0/"}]}.items():
        assert m[k] == v, k


def test_cli_wires_engine_run_to_the_input_file(monkeypatch):
    """`citam_b200 engine run <file>` reaches engine_run(input_file=...), as citam/cli.py:241-249 does."""
    from citam_b200 import cli

    seen = {}
    monkeypatch.setattr(cli, "engine_run", lambda **kw: seen.update(kw))
    cli.main(["engine", "run", "some_inputs.json"])
    assert seen == {"input_file": "some_inputs.json"}
