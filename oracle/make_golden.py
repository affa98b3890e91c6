#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (absent on the GPU box); the
fixtures it writes are committed so nothing reads the reference at test time.

Cases
  basic_seed0   examples/basic_example through `run_simulation` (the
                `citam engine run example_sim_inputs.json` code path) under
                np.random.seed(0): 20 agents, 1800 steps, 1 floor, 214 spaces.
  twofloor      tests/engine/data_navigation test_simple_facility duplicated to
                two floors with stairs: 8 agents, 3600
                steps; negative coordinates; floor changes.
  crafted       hand-made itineraries fed to the reference's own
                Simulation/Agent/CloseContactsCalculator through
                Precomputed{Schedule,Scheduler} subclasses of the reference
                ABCs: followers in hallways (hallway-adjacency contacts, events
                that break and resume), jittered off-network points, agents
                that never appear, itineraries shorter and longer than T.
Each case stores the hot path's INPUTS (itineraries, facility geometry, hallway
graph, contact distance) and the reference's OUTPUT files, gzip-compressed.
`loc_samples` pins the location rule on every visited point plus random lattice
points.
"""
import gzip
import json
import os
import shutil
import sys
import tempfile
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import refload  # noqa: E402

refload.load()
REF = refload.REFERENCE_ROOT
GOLD = os.path.join(REPO, "tests", "golden")

from citam.engine.calculators.close_contacts_calculator import CloseContactsCalculator  # noqa: E402
from citam.engine.core.simulation import Simulation  # noqa: E402
from citam.engine.facility.indoor_facility import Facility  # noqa: E402
from citam.engine.facility.navigation import Navigation  # noqa: E402
from citam.engine.io.input_parser import parse_input_file  # noqa: E402
from citam.engine.io.serializer import serializer  # noqa: E402
from citam.engine.schedulers.office_scheduler import OfficeScheduler  # noqa: E402
from citam.engine.schedulers.schedule import Schedule  # noqa: E402
from citam.engine.schedulers.scheduler import Scheduler  # noqa: E402
import citam.engine.main as ref_main  # noqa: E402

OUTPUT_FILES = [
    "pair_contact.csv",
    "contact_dist_per_agent.csv",
    "statistics.json",
    "raw_contact_data.ccd",
    "trajectory.txt",
]
FLOOR_FILES = ["contacts.txt", "contact_dist_per_coord.csv"]


# ---------------------------------------------------------------- extraction
def geometry_of(facility) -> dict:
    floors = []
    for f, fp in enumerate(facility.floorplans):
        spaces = []
        for sp in fp.spaces:
            spaces.append(
                {
                    "name": sp.unique_name,
                    "function": sp.space_function,
                    "path_bbox": [float(v) for v in sp.path.bbox()],
                    "boundaries": [
                        [l.start.real, l.start.imag, l.end.real, l.end.imag] for l in sp.boundaries
                    ],
                    "doors": [
                        [d.path.start.real, d.path.start.imag, d.path.end.real, d.path.end.imag]
                        for d in sp.doors
                    ],
                }
            )
        hg = facility.navigation.hallways_graph_per_floor[f]
        edges = None if hg is None else sorted([sorted([str(a), str(b)]) for a, b in hg.edges()])
        floors.append(
            {
                "spaces": spaces,
                "hallway_edges": edges,
                "minx": fp.minx, "maxx": fp.maxx, "miny": fp.miny, "maxy": fp.maxy,
            }
        )
    return {"floors": floors}


def pack_itineraries(itins):
    """Ragged (xy|None, floor|None) lists -> flat int32 arrays; floor=-1 encodes None."""
    off = np.zeros(len(itins) + 1, np.int64)
    for i, it in enumerate(itins):
        off[i + 1] = off[i] + len(it)
    x = np.zeros(off[-1], np.int32)
    y = np.zeros(off[-1], np.int32)
    fl = np.full(off[-1], -1, np.int16)
    k = 0
    for it in itins:
        for xy, f in it:
            if xy is not None:
                assert int(xy[0]) == xy[0] and int(xy[1]) == xy[1], xy
                x[k], y[k], fl[k] = int(xy[0]), int(xy[1]), int(f)
            k += 1
    return off, x, y, fl


def gz_copy(src, dst):
    with open(src, "rb") as a, gzip.GzipFile(dst, "wb", mtime=0) as b:
        shutil.copyfileobj(a, b)


def save_case(name, sim, workdir, total_timesteps, extra_loc_points=0, seed=0):
    out = os.path.join(GOLD, name)
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(os.path.join(out, "expected"))
    facility = sim.facility
    geom = geometry_of(facility)
    with gzip.GzipFile(os.path.join(out, "geometry.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(geom).encode())
    itins = [a.schedule.itinerary for a in sim.agents.values()]
    off, x, y, fl = pack_itineraries(itins)
    # location samples: every visited point + random lattice points per floor
    pts = set()
    for it in itins:
        for xy, f in it:
            if xy is not None:
                pts.add((int(f), int(xy[0]), int(xy[1])))
    rng = np.random.RandomState(seed + 12345)
    for f, fg in enumerate(geom["floors"]):
        bbs = np.array([s["path_bbox"] for s in fg["spaces"]])
        lo_x, hi_x = int(bbs[:, 0].min()) - 3, int(bbs[:, 1].max()) + 3
        lo_y, hi_y = int(bbs[:, 2].min()) - 3, int(bbs[:, 3].max()) + 3
        for _ in range(extra_loc_points):
            pts.add((f, int(rng.randint(lo_x, hi_x + 1)), int(rng.randint(lo_y, hi_y + 1))))
        # corners / door ends / wall points: the rule's special cases
        for s in fg["spaces"][:: max(1, len(fg["spaces"]) // 40)]:
            for seg in s["boundaries"] + s["doors"]:
                for px, py in ((seg[0], seg[1]), (seg[2], seg[3]), ((seg[0] + seg[2]) / 2, (seg[1] + seg[3]) / 2)):
                    for ddx in (-1, 0, 1):
                        for ddy in (-1, 0, 1):
                            pts.add((f, int(round(px)) + ddx, int(round(py)) + ddy))
    pts = sorted(pts)
    locs = []
    for f, px, py in pts:
        r = facility.floorplans[f].identify_this_location(px, py, include_boundaries=True)
        locs.append(-1 if r is None else r)
    np.savez_compressed(
        os.path.join(out, "inputs.npz"),
        itin_off=off, itin_x=x, itin_y=y, itin_floor=fl,
        total_timesteps=np.int64(total_timesteps),
        contact_distance=np.float64(sim.calculators[0].contact_distance),
        loc_points=np.array(pts, np.int32).reshape(-1, 3),
        loc_values=np.array(locs, np.int32),
    )
    for fn in OUTPUT_FILES:
        gz_copy(os.path.join(workdir, fn), os.path.join(out, "expected", fn + ".gz"))
    for f in range(facility.number_of_floors):
        os.makedirs(os.path.join(out, "expected", "floor_%d" % f))
        for fn in FLOOR_FILES:
            gz_copy(os.path.join(workdir, "floor_%d" % f, fn), os.path.join(out, "expected", "floor_%d" % f, fn + ".gz"))
    print("saved", name, "agents", len(itins), "loc samples", len(pts))


# ------------------------------------------------------------------- cases
def case_basic(seed=0):
    tmp = tempfile.mkdtemp(prefix="citam_gold_")
    shutil.copytree(os.path.join(REF, "examples/basic_example/foo_facility"), os.path.join(tmp, "foo_facility"))
    shutil.copy(os.path.join(REF, "examples/basic_example/example_sim_inputs.json"), tmp)
    cwd = os.getcwd()
    os.chdir(tmp)
    captured = {}
    orig = Simulation.run_serial

    def spy(self, *a, **k):
        captured["sim"] = self
        return orig(self, *a, **k)

    Simulation.run_serial = spy
    try:
        np.random.seed(seed)
        inputs = parse_input_file("example_sim_inputs.json")
        ref_main.run_simulation(inputs)
    finally:
        Simulation.run_serial = orig
        os.chdir(cwd)
    sim = captured["sim"]
    save_case("basic_seed%d" % seed, sim, tmp, sim.total_timesteps, extra_loc_points=6000, seed=seed)
    # the entry point's own files (main.py:420-433): pins `citam engine run` beyond the step loop
    gz_copy(os.path.join(tmp, "policy.json"), os.path.join(GOLD, "basic_seed%d" % seed, "expected", "policy.json.gz"))
    return sim, tmp


def two_floor_facility():
    cachedir = os.path.join(REF, "citam/tests/engine/data_navigation")
    os.environ["CITAM_CACHE_DIRECTORY"] = cachedir
    datadir = os.path.join(cachedir, "floorplans_and_nav_data")
    with open(os.path.join(datadir, "test_simple_facility", "floor_0", "updated_floorplan.json")) as f:
        fp = json.load(f, object_hook=serializer.decoder_hook)
    fp2 = deepcopy(fp)
    for sp in fp2.spaces:
        sp.unique_name = sp.unique_name + "_2"
    fp2.floor_name = "1"
    fps = [fp, fp2]
    fps[0].spaces[4].space_function = "stairs"
    fps[1].spaces[4].space_function = "stairs"
    nav = Navigation(fps, "test_simple_facility", None, multifloor_type="xy")
    return Facility(fps, [{"name": "1", "floor": "0"}], "test_simple_facility", navigation=nav)


def case_twofloor(seed=3):
    np.random.seed(seed)
    fac = two_floor_facility()
    T = 3600
    sim = Simulation(
        fac, T, 8, calculators=[CloseContactsCalculator(fac)],
        scheduler=OfficeScheduler(fac, T, entry_exit_window=100),
    )
    tmp = tempfile.mkdtemp(prefix="citam_gold_")
    sim.run_serial(tmp, "s", "r")
    save_case("twofloor", sim, tmp, T, extra_loc_points=4000, seed=seed)
    return sim, tmp


class PrecomputedSchedule(Schedule):
    """A reference `Schedule` whose itinerary is given (hot-path input format)."""

    def __init__(self, itinerary, navigation):
        super().__init__(1, 0, len(itinerary))
        self.itinerary = itinerary
        self.navigation = navigation
        self.current_standing = 0

    def build(self):
        pass

    def get_next_position(self):  # same contract as office_schedule.py:632-638
        position = self.itinerary[self.current_standing]
        if self.current_standing < len(self.itinerary) - 1:
            self.current_standing += 1
        return position

    def __str__(self):
        return " precomputed"


class PrecomputedScheduler(Scheduler):
    def __init__(self, facility, total_timesteps, itineraries):
        super().__init__(facility, 1, total_timesteps)
        self.itineraries = itineraries

    def run(self, n_agents):
        return [PrecomputedSchedule(it, self.facility.navigation) for it in self.itineraries[:n_agents]]

    def save_to_files(self, work_directory):
        pass


def case_crafted(base_sim, seed=7):
    """Followers and jitter on top of the seed-0 itineraries of the example facility."""
    rng = np.random.RandomState(seed)
    fac = base_sim.facility
    base = [a.schedule.itinerary for a in base_sim.agents.values()]
    T = 1500
    itins = []
    for k in range(44):
        src = base[k % len(base)]
        delay = int(rng.randint(0, 4)) if k >= len(base) else 0
        it = [(None, None)] * delay + [e for e in src]
        if k >= 30:  # jitter: off-network points, some outside every space
            it2 = []
            jx, jy = int(rng.randint(-4, 5)), int(rng.randint(-4, 5))
            for xy, f in it:
                it2.append(((xy[0] + jx, xy[1] + jy), f) if xy is not None else (None, None))
            it = it2
        if k % 7 == 3:  # occasional drop-outs: None in the middle is sticky
            it = [e if rng.rand() > 0.02 else (None, None) for e in it]
        itins.append(it)
    itins[5] = itins[5][:700]          # shorter than T: parks at its last entry
    itins[6] = itins[6][:400] + [(None, None)] * 5
    itins.append([(None, None)] * 50)  # never appears
    itins.append([(None, None)])       # single-entry itinerary
    itins.append(base[2] + base[2][::-1])  # longer than T
    far = (100000, 100000)
    itins.append([(far, 0)] * 30)      # active but outside every space (location None)
    n = len(itins)
    sim = Simulation(
        fac, T, n, calculators=[CloseContactsCalculator(fac, contact_distance=9.5)],
        scheduler=PrecomputedScheduler(fac, T, itins),
    )
    tmp = tempfile.mkdtemp(prefix="citam_gold_")
    for f in range(fac.number_of_floors):
        os.makedirs(os.path.join(tmp, "floor_%d" % f), exist_ok=True)
    # run_serial minus save_maps (map.svg is not on the path); same calls otherwise
    sim.create_agents()
    for c in sim.calculators:
        c.initialize(sim.agents, tmp)
    # heatmap needs map.svg: give it a minimal one
    for f in range(fac.number_of_floors):
        with open(os.path.join(tmp, "floor_%d" % f, "map.svg"), "w") as fh:
            fh.write('<svg xmlns="http://www.w3.org/2000/svg"><g id="root"></g></svg>')
    sim.run_simulation_and_save_results(tmp)
    save_case("crafted", sim, tmp, T, extra_loc_points=0, seed=seed)
    return sim, tmp


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ["basic", "twofloor", "crafted"]
    base = None
    if "basic" in which or "crafted" in which:
        base, _ = case_basic(0)
    if "twofloor" in which:
        case_twofloor()
    if "crafted" in which:
        case_crafted(base)
