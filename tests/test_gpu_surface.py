"""GPU: the reference-facing surface (Simulation / CloseContactsCalculator / ContactEvents and the
output files) against the files the UNMODIFIED reference wrote for the same itineraries."""
import os
import re

import numpy as np
import pytest

from expected_util import *
from golden_util import CASES, expected, load_case

pytestmark = pytest.mark.gpu

FILES = ["pair_contact.csv", "contact_dist_per_agent.csv", "statistics.json", "raw_contact_data.ccd", "trajectory.txt"]


def _files(case):
    out = list(FILES)
    for f in range(len(case["geometry"]["floors"])):
        out += ["floor_%d/contacts.txt" % f, "floor_%d/contact_dist_per_coord.csv" % f]
    return out


def _compare(case, workdir):
    for fn in _files(case):
        with open(os.path.join(workdir, fn), "rb") as fh:
            got = fh.read()
        want = expected(case, fn)
        if fn == "raw_contact_data.ccd":  # numpy>=2 scalar reprs in the reference's dump (values equal)
            want = re.sub(rb"np\.int64\((-?\d+)\)", rb"\1", want)
        assert got == want, fn


def _simulation(case):
    from citam_b200.calculator import CloseContactsCalculator
    from citam_b200.facility import GeometryFacility
    from citam_b200.scheduler import PrecomputedScheduler
    from citam_b200.simulation import Simulation

    fac = GeometryFacility(case["geometry"], case["name"])
    n = len(case["itineraries"])
    return Simulation(
        fac, case["T"], n, calculators=[CloseContactsCalculator(fac, case["D"])],
        scheduler=PrecomputedScheduler(fac, case["T"], case["itineraries"]),
    )


@pytest.mark.parametrize("name", CASES)
def test_run_serial_block_path_files_byte_identical(name, tmp_path):
    case = load_case(name)
    sim = _simulation(case)
    sim.block_steps = 700  # several blocks
    sim.run_serial(str(tmp_path), "s", "r")
    assert sim.current_step == case["T"]
    _compare(case, str(tmp_path))
    for fn in ("manifest.json", "schedules.txt"):
        assert os.path.isfile(os.path.join(str(tmp_path), fn))
    import json

    m = json.load(open(os.path.join(str(tmp_path), "manifest.json")))
    for k in ("RunID", "SimulationHash", "RunName", "SimulationName", "FacilityName", "NumberOfAgents", "Floors",
              "TotalTimesteps", "TrajectoryFile", "ScaleMultiplier"):
        assert k in m
    want_dur = parse_agent_csv(expected(case, "contact_dist_per_agent.csv"))
    assert [a.cumulative_properties["contact_duration"] for a in sim.agents.values()] == want_dur
    # dashboard svgs: one circle per contact coordinate on top of the floor map
    import xml.etree.ElementTree as ET

    for f in range(len(case["geometry"]["floors"])):
        d = os.path.join(str(tmp_path), "floor_%d" % f)
        assert ET.parse(os.path.join(d, "map.svg")).getroot()[0].attrib["id"] == "root"
        g = ET.parse(os.path.join(d, "heatmap.svg")).getroot()[0]
        n_coords = len(parse_coord_csv(expected(case, "floor_%d/contact_dist_per_coord.csv" % f)))
        assert sum(1 for e in g if e.tag.endswith("circle")) == n_coords


@pytest.mark.parametrize("name", ["twofloor", "crafted"])
def test_step_by_step_plugin_path_files_byte_identical(name, tmp_path):
    """Simulation.step -> Calculator.run(current_step, agents): one device call per step."""
    case = load_case(name)
    sim = _simulation(case)
    wd = str(tmp_path)
    sim.create_sim_hash()
    sim.save_manifest(wd, "s", "r")
    sim.create_agents()
    for c in sim.calculators:
        c.initialize(sim.agents, wd)
    with open(os.path.join(wd, "trajectory.txt"), "w") as traj:
        for _ in range(case["T"]):
            sim.step(traj_outfile=traj)
    for c in sim.calculators:
        c.finalize(sim.agents, work_directory=wd)
    assert sim.current_step == case["T"]
    _compare(case, wd)
    ce = sim.calculators[0].contact_events
    assert ce.count() == sum(r[2] for r in parse_pair_csv(expected(case, "pair_contact.csv")))


def test_calculator_protocol_errors():
    from citam_b200.agent import Agent
    from citam_b200.calculator import CloseContactsCalculator
    from citam_b200.facility import GeometryFacility
    from citam_b200.scheduler import PrecomputedScheduler, Schedule
    from citam_b200.simulation import Simulation
    from collections import OrderedDict

    case = load_case("twofloor")
    fac = GeometryFacility(case["geometry"])
    calc = CloseContactsCalculator(fac)
    a = Agent(0, Schedule([(None, None)]))
    a.cumulative_properties["contact_duration"] = 3
    with pytest.raises(ValueError):
        calc.initialize(OrderedDict([(0, a)]))
    with pytest.raises(ValueError):
        Simulation(fac, 10, 2, calculators="nope", scheduler=PrecomputedScheduler(fac, 10, []))
    with pytest.raises(ValueError):
        Simulation(fac, 10, 2, calculators=[object()], scheduler=PrecomputedScheduler(fac, 10, []))
    with pytest.raises(ValueError):
        Simulation(fac, 10, 2, calculators=calc, scheduler="nope")


def test_cli_run_precomputed(tmp_path):
    from citam_b200.cli import main
    from golden_util import GOLD

    d = os.path.join(GOLD, "basic_seed0")
    main(["engine", "run-precomputed", os.path.join(d, "inputs.npz"), os.path.join(d, "geometry.json.gz"), str(tmp_path)])
    case = load_case("basic_seed0")
    _compare(case, str(tmp_path))
    assert os.path.isfile(os.path.join(str(tmp_path), "timing.txt"))


def test_summary_mode_large_run_tables_only(tmp_path):
    """full_fidelity=False: no per-step files, same end-of-run tables."""
    from citam_b200.calculator import CloseContactsCalculator
    from citam_b200.facility import GeometryFacility
    from citam_b200.scheduler import PrecomputedScheduler
    from citam_b200.simulation import Simulation

    case = load_case("crafted")
    fac = GeometryFacility(case["geometry"])
    n = len(case["itineraries"])
    sim = Simulation(fac, case["T"], n, calculators=[CloseContactsCalculator(fac, case["D"], full_fidelity=False)],
                     scheduler=PrecomputedScheduler(fac, case["T"], case["itineraries"]))
    sim.run_serial(str(tmp_path), "s", "r")
    for fn in ("pair_contact.csv", "contact_dist_per_agent.csv", "statistics.json"):
        with open(os.path.join(str(tmp_path), fn), "rb") as fh:
            assert fh.read() == expected(case, fn), fn
    got = parse_coord_csv(open(os.path.join(str(tmp_path), "floor_0/contact_dist_per_coord.csv"), "rb").read())
    assert got == parse_coord_csv(expected(case, "floor_0/contact_dist_per_coord.csv"))
    assert not os.path.exists(os.path.join(str(tmp_path), "trajectory.txt"))
    with pytest.raises(RuntimeError):
        sim.calculators[0].contact_events.contact_data
