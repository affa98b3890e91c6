"""CPU restatement of CITAM's per-timestep loop ("the oracle port").

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.
The product package `citam_b200` never does; its hot path is the CUDA library.

It restates, step for step and with the same third-party calls, what the
upstream reference does on the path (citations relative to /root/reference):

  position advance     citam/engine/core/agent.py:67-109,
                       citam/engine/schedulers/office_schedule.py:632-638
  location rule        citam/engine/map/floorplan.py:215-241,
                       citam/engine/map/space.py:179-280,300-312,
                       citam/engine/map/geometry.py:31-127,558-588
  step loop            citam/engine/core/simulation.py:291-361
  contact calculator   citam/engine/calculators/close_contacts_calculator.py:64-231
  event bookkeeping    citam/engine/calculators/contacts.py:92-142
  end-of-run outputs   citam/engine/calculators/contacts.py:155-368,
                       citam/engine/calculators/close_contacts_calculator.py:233-318

Parity pinning: `tests/test_oracle_golden.py` checks this port byte-for-byte
against output files produced by the real reference (run in the build
container through `oracle/refload.py`; fixtures + generator script committed
under `tests/golden/` and `oracle/make_golden.py`), and against the known-answer
values of the reference's own tests (test_close_contacts_calculator.py,
test_contact_events.py).

Data model (plain Python / numpy, no reference objects):
  geometry = {"floors": [ {"spaces": [space...], "hallway_edges": [[a,b]...]|None} ]}
  space    = {"name", "function", "path_bbox": [xmin,xmax,ymin,ymax],
              "boundaries": [[sx,sy,ex,ey]...], "doors": [[sx,sy,ex,ey]...]}
  itinerary (per agent) = list of ((x, y) | None, floor | None)
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.spatial.distance

HALLWAY_WORDS = ("circulation", "aisle", "lobby", "vestibule")


# --------------------------------------------------------------------------
# location rule
# --------------------------------------------------------------------------
def is_hallway(space_function: str) -> bool:
    """space.py:300-312 — substring test on the lower-cased function."""
    f = space_function.lower()
    return any(w in f for w in HALLWAY_WORDS)


def _in_box(px, py, qx, qy, rx, ry) -> bool:
    """geometry.py:31-53 `on_segment(p, q, r)`: q inside the box spanned by p, r."""
    return (
        qx <= max(px, rx) and qx >= min(px, rx) and qy <= max(py, ry) and qy >= min(py, ry)
    )


def _orient(px, py, qx, qy, rx, ry) -> int:
    """geometry.py:56-82: sign class of the fp64 cross product (0 colinear, 1 cw, 2 ccw)."""
    val = (qy - py) * (rx - qx) - (qx - px) * (ry - qy)
    if val == 0:
        return 0
    return 1 if val > 0 else 2


def _segments_cross(p1, q1, p2, q2) -> bool:
    """geometry.py:85-127 `do_intersect`."""
    o1 = _orient(*p1, *q1, *p2)
    o2 = _orient(*p1, *q1, *q2)
    o3 = _orient(*p2, *q2, *p1)
    o4 = _orient(*p2, *q2, *q1)
    if o1 != o2 and o3 != o4:
        return True
    if o1 == 0 and _in_box(*p1, *p2, *q1):
        return True
    if o2 == 0 and _in_box(*p1, *q2, *q1):
        return True
    if o3 == 0 and _in_box(*p2, *p1, *q2):
        return True
    if o4 == 0 and _in_box(*p2, *q1, *q2):
        return True
    return False


def _unit_normal(sx, sy, ex, ey) -> complex:
    """svgpathtools Line.normal: -1j * (end-start)/|end-start| (complex arithmetic)."""
    d = complex(ex, ey) - complex(sx, sy)
    return -1j * (d / abs(d))


def _point_on_line(seg, x, y, tol=1e-3) -> bool:
    """geometry.py:558-588 `is_point_on_line`."""
    sx, sy, ex, ey = seg
    if (x == sx and y == sy) or (x == ex and y == ey):
        return True
    if _in_box(sx, sy, x, y, ex, ey):
        n1 = _unit_normal(sx, sy, ex, ey)
        n2 = _unit_normal(sx, sy, x, y)
        if abs(n1.real - n2.real) < tol and abs(n1.imag - n2.imag) < tol:
            return True
    return False


_INF = 1e10


def point_in_space(space: dict, x, y) -> bool:
    """space.py:194-280 with include_boundaries=True (agent.py:102-106)."""
    xmin, xmax, ymin, ymax = space["path_bbox"]
    if x < round(xmin) or x > round(xmax) or y < round(ymin) or y > round(ymax):
        return False
    if not is_hallway(space["function"]) and len(space["doors"]) > 0:
        for dsx, dsy, dex, dey in space["doors"]:
            if _in_box(dsx, dsy, x, y, dex, dey):
                return True
    for seg in space["boundaries"]:
        sx, sy, ex, ey = seg
        if abs(complex(ex, ey) - complex(sx, sy)) > 1.0 and _point_on_line(seg, x, y):
            return True
    ends = (
        (_INF, y + 5.0),
        (x - 5.0, _INF),
        (-_INF, y + 5.0),
        (x - 5.0, -_INF),
        (_INF, _INF),
        (_INF, -_INF),
        (-_INF, _INF),
        (-_INF, -_INF),
    )
    for ext in ends:
        count = 0
        for sx, sy, ex, ey in space["boundaries"]:
            if _segments_cross((sx, sy), (ex, ey), (x, y), ext):
                count += 1
        if count % 2 == 1:
            return True
    return False


def identify_location(floor_geom: dict, x, y) -> Optional[int]:
    """floorplan.py:215-241: first space (list order) that accepts the point."""
    for i, sp in enumerate(floor_geom["spaces"]):
        if point_in_space(sp, x, y):
            return i
    return None


class LocationCache:
    """Memoised identify_location — LOC is a pure function of (floor, x, y)."""

    def __init__(self, geometry: dict):
        self.geometry = geometry
        self.memo: Dict[Tuple[int, int, int], Optional[int]] = {}

    def __call__(self, floor: int, x, y) -> Optional[int]:
        k = (floor, x, y)
        if k not in self.memo:
            self.memo[k] = identify_location(self.geometry["floors"][floor], x, y)
        return self.memo[k]


# --------------------------------------------------------------------------
# agents and the step loop
# --------------------------------------------------------------------------
class PortAgent:
    """State the reference keeps per agent (agent.py:25-65)."""

    __slots__ = ("unique_id", "itin", "cursor", "pos", "floor", "loc", "dur")

    def __init__(self, unique_id: int, itinerary: Sequence):
        self.unique_id = unique_id
        self.itin = itinerary
        self.cursor = 0
        self.pos = None
        self.floor = None
        self.loc = None
        self.dur = 0

    def step(self, locate) -> None:
        """agent.py:67-109 + office_schedule.py:632-638 (sticky on None)."""
        xy, fl = self.itin[self.cursor]
        if self.cursor < len(self.itin) - 1:
            self.cursor += 1
        if xy is not None:
            self.pos = (xy[0], xy[1])
            self.loc = locate(fl, xy[0], xy[1])
            self.floor = fl


class PortContactEvents:
    """contacts.py:73-142: key "lo-hi" -> list of events (dicts)."""

    def __init__(self):
        self.contact_data: Dict[str, List[dict]] = {}

    def add_contact(self, a1: PortAgent, a2: PortAgent, step: int, position) -> None:
        if a1.unique_id == a2.unique_id:
            raise ValueError("Agents must be different.")
        key = "%s-%s" % (a1.unique_id, a2.unique_id)
        ev = self.contact_data.get(key)
        if ev is not None:
            last = ev[-1]
            if last["start_step"] + last["duration"] == step:
                last["duration"] += 1
                last["positions"].append(position)
                last["locations"].append(a1.loc)
                last["floor_numbers"].append(a1.floor)
                return
            ev.append(_new_event(a1, position, step))
        else:
            self.contact_data[key] = [_new_event(a1, position, step)]

    def count(self) -> int:
        return sum(len(v) for v in self.contact_data.values())


def _new_event(a1: PortAgent, position, step: int) -> dict:
    # attribute order follows ContactEvent.__init__ (contacts.py:66-70) so that
    # str(__dict__) in the raw dump matches
    return {
        "locations": [a1.loc],
        "floor_numbers": [a1.floor],
        "positions": [position],
        "start_step": step,
        "duration": 1,
    }


def proximity_indices(positions: np.ndarray, contact_distance: float) -> np.ndarray:
    """close_contacts_calculator.py:152-171 — the same scipy/numpy calls."""
    dists = scipy.spatial.distance.pdist(positions)
    return np.argwhere(scipy.spatial.distance.squareform(dists) < contact_distance)


class PortSimulation:
    """The T-step loop with the close-contacts calculator inlined.

    `adjacency[f]` is None or a set of frozenset({name1, name2}) (undirected
    hallway graph, close_contacts_calculator.py:190-203).
    """

    def __init__(
        self,
        itineraries: Sequence[Sequence],
        geometry: dict,
        contact_distance: float = 6.0,
        locate=None,
    ):
        self.geometry = geometry
        self.n_floors = len(geometry["floors"])
        self.contact_distance = contact_distance
        self.locate = locate if locate is not None else LocationCache(geometry)
        self.agents = [PortAgent(i, it) for i, it in enumerate(itineraries)]
        self.events = PortContactEvents()
        self.current_step = 0
        self.adjacency = []
        for fl in geometry["floors"]:
            e = fl.get("hallway_edges")
            self.adjacency.append(None if e is None else {frozenset(p) for p in e})
        self.hallway = [[is_hallway(s["function"]) for s in fl["spaces"]] for fl in geometry["floors"]]
        self.names = [[s["name"] for s in fl["spaces"]] for fl in geometry["floors"]]

    # simulation.py:338-361
    def move_agents(self) -> List[PortAgent]:
        active = []
        for a in self.agents:
            a.step(self.locate)
            if a.pos is not None:
                active.append(a)
        return active

    # close_contacts_calculator.py:64-137
    def contacts(self, step: int, active: List[PortAgent], contact_out=None) -> None:
        step_locs: Dict[Tuple[float, float], int] = {}  # ONE dict shared by all floors (:86-88)
        if len(active) <= 1:
            return
        agents = [a for a in active if a.loc is not None]
        if len(agents) == 0:
            return
        pos = np.array([a.pos for a in agents])
        for i, j in proximity_indices(pos, self.contact_distance):
            if i >= j:
                continue
            a1, a2 = agents[i], agents[j]
            if a1.floor != a2.floor:
                continue
            if a1.loc != a2.loc:
                f = a1.floor
                adj = self.adjacency[f]
                if adj is None or not self.hallway[f][a1.loc] or not self.hallway[f][a2.loc]:
                    continue
                n1, n2 = self.names[f][a1.loc], self.names[f][a2.loc]
                if frozenset((n1, n2)) not in adj:
                    continue
            # add_contact_event (:205-231)
            cp = (a1.pos[0] + (a2.pos[0] - a1.pos[0]) / 2.0, a1.pos[1] + (a2.pos[1] - a1.pos[1]) / 2.0)
            a1.dur += 1
            a2.dur += 1
            self.events.add_contact(a1, a2, step, cp)
            step_locs[cp] = step_locs.get(cp, 0) + 1
        if contact_out is not None:
            for f in range(self.n_floors):
                contact_out[f].write(str(len(step_locs)) + "\nstep :" + str(step) + "\n")
                for cl, nc in step_locs.items():
                    contact_out[f].write(str(cl[0]) + "\t" + str(cl[1]) + "\t" + str(nc) + "\n")

    # simulation.py:291-336
    def step(self, traj_out=None, contact_out=None) -> None:
        active = self.move_agents()
        self.contacts(self.current_step, active, contact_out)
        if traj_out is not None:
            traj_out.write(str(len(active)) + "\nstep: " + str(self.current_step) + "\n")
            for a in active:
                traj_out.write(
                    str(a.unique_id) + "\t" + str(a.pos[0]) + "\t" + str(a.pos[1]) + "\t"
                    + str(a.floor) + "\t" + str(a.dur) + "\n"
                )
        self.current_step += 1

    def run(self, total_timesteps: int, workdir: Optional[str] = None) -> "PortSimulation":
        traj = None
        cfiles = None
        if workdir is not None:
            cfiles = []
            for f in range(self.n_floors):
                d = os.path.join(workdir, "floor_" + str(f))
                os.makedirs(d, exist_ok=True)
                cfiles.append(open(os.path.join(d, "contacts.txt"), "w"))
            traj = open(os.path.join(workdir, "trajectory.txt"), "w")
        for _ in range(total_timesteps):
            self.step(traj, cfiles)
        if workdir is not None:
            traj.close()
            for c in cfiles:
                c.close()
            write_outputs(self.events.contact_data, [a.dur for a in self.agents], self.n_floors, workdir)
        return self


# --------------------------------------------------------------------------
# end-of-run outputs
# --------------------------------------------------------------------------
def pair_rows(contact_data: Dict[str, List[dict]]) -> List[Tuple[str, str, int, int]]:
    """contacts.py:155-183: rows in dict insertion (= first contact) order."""
    rows = []
    for key, evs in contact_data.items():
        a1, a2 = key.split("-")[0], key.split("-")[1]
        rows.append((a1, a2, len(evs), sum(e["duration"] for e in evs)))
    return rows


def statistics_from_pairs(rows: Sequence[Tuple[str, str, int, int]]) -> List[dict]:
    """contacts.py:185-294 — six statistics from (agent1, agent2, n_events, duration) rows."""
    n_others: Dict[str, int] = {}
    n_contacts: Dict[str, int] = {}
    dur: Dict[str, int] = {}
    total = 0
    for a1, a2, n, d in rows:
        total += d
        for a in (a1, a2):
            n_contacts[a] = n_contacts.get(a, 0) + n
            dur[a] = dur.get(a, 0) + d
        n_others[a1] = n_others.get(a1, 0) + 1  # only the first agent (:228)
    n_with = len(n_contacts)
    avg_n = sum(n_contacts.values()) / n_with if n_with > 0 else 0
    avg_d = sum(dur.values()) / (n_with * 2.0) if n_with > 0 else 0
    avg_p = sum(n_others.values()) / n_with if n_with > 0 else 0
    mx = max(n_contacts.values()) if n_contacts else 0
    return [
        {"name": "overall_total_contact_duration", "value": round(total / 60.0, 2), "unit": "min"},
        {"name": "avg_n_contacts_per_agent", "value": round(avg_n, 2), "unit": ""},
        {"name": "avg_contact_duration_per_agent", "value": round(avg_d / 60.0, 2), "unit": "min"},
        {"name": "n_agents_with_contacts", "value": n_with, "unit": ""},
        {"name": "avg_number_of_people_per_agent", "value": round(avg_p, 2), "unit": ""},
        {"name": "max_contacts", "value": mx, "unit": ""},
    ]


def contacts_per_coordinates(contact_data: Dict[str, List[dict]], floor: int) -> Dict[Tuple[float, float], int]:
    """contacts.py:296-346 (including the per-pair `set` iteration order)."""
    out: Dict[Tuple[float, float], int] = {}
    for evs in contact_data.values():
        fpos = []
        for e in evs:
            for i, p in enumerate(e["positions"]):
                if e["floor_numbers"][i] == floor:
                    fpos.append(p)
        for p in list(set(fpos)):
            out[p] = out.get(p, 0) + fpos.count(p)
    return out


def write_outputs(contact_data, durations: Sequence[int], n_floors: int, workdir: str) -> None:
    """close_contacts_calculator.py:250-318 minus the heatmap."""
    with open(os.path.join(workdir, "contact_dist_per_agent.csv"), "w") as f:
        f.write("agent_ID,Number_of_Contacts\n")
        for i, d in enumerate(durations):
            f.write("ID_" + str(i + 1) + "," + str(d) + "\n")
    rows = pair_rows(contact_data)
    with open(os.path.join(workdir, "pair_contact.csv"), "w") as f:
        f.write("Agent1,Agent2,N_Contacts,TotalContactDuration\n")
        for a1, a2, n, d in rows:
            f.write(a1 + "," + a2 + "," + str(n) + "," + str(d) + "\n")
    with open(os.path.join(workdir, "raw_contact_data.ccd"), "w") as f:
        f.write(str(contact_data))
    with open(os.path.join(workdir, "statistics.json"), "w") as f:
        json.dump({"data": statistics_from_pairs(rows)}, f)
    for fl in range(n_floors):
        d = os.path.join(workdir, "floor_" + str(fl))
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "contact_dist_per_coord.csv"), "a") as f:  # append mode (:309)
            f.write("x,y,n_contacts\n")
            for (x, y), c in contacts_per_coordinates(contact_data, fl).items():
                f.write(str(x) + "," + str(y) + "," + str(c) + "\n")


# --------------------------------------------------------------------------
# order-independent summary (what the device tables hold) from the event dict
# --------------------------------------------------------------------------
def summary_tables(contact_data, durations, n_floors):
    """Reduce the event dict to the commutative tables the CUDA path accumulates.

    pairs: (i, j, first_step, n_events, duration) sorted by (first_step, i, j);
    coords: {(floor, x1+x2, y1+y2): count}; durations: per-agent contact seconds.
    """
    pairs = []
    coords: Dict[Tuple[int, int, int], int] = {}
    for key, evs in contact_data.items():
        i, j = (int(s) for s in key.split("-"))
        pairs.append((i, j, evs[0]["start_step"], len(evs), sum(e["duration"] for e in evs)))
        for e in evs:
            for p, fl in zip(e["positions"], e["floor_numbers"]):
                k = (int(fl), int(round(p[0] * 2)), int(round(p[1] * 2)))
                coords[k] = coords.get(k, 0) + 1
    pairs.sort(key=lambda r: (r[2], r[0], r[1]))
    return pairs, coords, list(durations)
