"""Synthetic facilities and itineraries of the BASELINE.json shapes (configs 2-5).

The reference builds itineraries with its scheduler + A* router
(citam/engine/schedulers/office_schedule.py:423-593, citam/engine/facility/navigation.py:608-949);
that is upstream of the hot path and out of scope here.  This module produces inputs of the
same *shape* directly in the hot path's input format, vectorised so that 10^6 agents take
seconds: a facility geometry dict (same schema as tests/golden/*/geometry.json.gz, consumed by
`DeviceEngine.load_geometry`, the oracle port and the C grid oracle alike) and a CSR store of
constant-velocity `citam_segment` runs (include/citam_b200.h).

Floor plan (integer lattice, default 1200 x 1200 like the example facility):
a vertical "circulation" spine, five bands of [room row | corridor | room row] left and right
of it, 48 x 96 rooms with a 6-unit door on the corridor wall, a few rooms per band designated
conference room / restroom / cafeteria.  The hallway graph links every corridor to the spine.

Agents (office_schedule.py:108-110,499-593 in spirit): not in the facility until their shift's
entry time; enter at the floor entrance (spine foot), walk axis-aligned integer steps at the
reference's pace (2-6 ft/s = 24-72 drawing units per step at this facility's scale) to their desk, alternate stays and walks over `n_stops` destinations (desk,
conference room, restroom, cafeteria, another office; a few per cent of the shared destinations
are on another floor, reached by the stairs at the spine head), keeping to a personal lane in
corridors and spine, then walk out to a parking point *outside every space* (location None), so that
finished agents do not pile up in contact at the door (SURVEY.md §7 "parked-at-exit blow-up").
Everything is a pure function of the seed (numpy PCG64).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from ._lib import SEGMENT

ROOM_W, ROOM_H, CORR_H, SPINE_W = 48, 96, 24, 24
N_BANDS, ROOMS_PER_ROW = 5, 12
BAND_H = 2 * ROOM_H + CORR_H          # 216
Y_BASE = 60
FLOOR_W = 2 * (12 + ROOMS_PER_ROW * ROOM_W) - 0  # 1176 + spine = 1200
SPINE_X0 = 12 + ROOMS_PER_ROW * ROOM_W           # 588
SPINE_X1 = SPINE_X0 + SPINE_W                     # 612
SPINE_XC = (SPINE_X0 + SPINE_X1) // 2             # 600
FLOOR_H = 1200
FEET = 12  # drawing units per foot (example facility: floorplan_scale = 1/12 ft per unit)
STAIRS_Y = FLOOR_H - 10
PARK_Y = -30

KIND_OFFICE, KIND_CONF, KIND_REST, KIND_CAFE = 0, 1, 2, 3
_FUNCTION = {KIND_OFFICE: "office", KIND_CONF: "conference room", KIND_REST: "restroom", KIND_CAFE: "cafeteria"}


def _rect_space(name: str, function: str, x0: int, y0: int, x1: int, y1: int, doors) -> Dict[str, Any]:
    return {
        "name": name,
        "function": function,
        "path_bbox": [float(x0), float(x1), float(y0), float(y1)],
        "boundaries": [
            [float(x0), float(y0), float(x1), float(y0)],
            [float(x1), float(y0), float(x1), float(y1)],
            [float(x1), float(y1), float(x0), float(y1)],
            [float(x0), float(y1), float(x0), float(y0)],
        ],
        "doors": [[float(a), float(b), float(c), float(d)] for a, b, c, d in doors],
    }


class SyntheticFacility:
    """Geometry dict + the room table the itinerary generator draws from."""

    def __init__(self, n_floors: int = 1):
        self.n_floors = int(n_floors)
        floors = []
        # room table (identical on every floor): rect, door x, corridor centre y, kind
        rx0, ry0, rx1, ry1, rdoor, rcorr, rkind = [], [], [], [], [], [], []
        for f in range(self.n_floors):
            spaces = [_rect_space("spine_%d" % f, "circulation", SPINE_X0, 0, SPINE_X1, FLOOR_H, [])]
            edges = []
            for b in range(N_BANDS):
                yb = Y_BASE + BAND_H * b
                yc0, yc1 = yb + ROOM_H, yb + ROOM_H + CORR_H
                for side, (cx0, cx1) in enumerate(((0, SPINE_X0), (SPINE_X1, FLOOR_W + SPINE_W))):
                    cname = "corr_%d_%d_%d" % (f, b, side)
                    spaces.append(_rect_space(cname, "circulation", cx0, yc0, cx1, yc1, []))
                    edges.append(sorted([cname, "spine_%d" % f]))
            for b in range(N_BANDS):
                yb = Y_BASE + BAND_H * b
                yc0, yc1 = yb + ROOM_H, yb + ROOM_H + CORR_H
                for side in range(2):
                    xb = 12 if side == 0 else SPINE_X1
                    for row in range(2):  # 0: below the corridor, 1: above
                        y0, y1 = (yb, yc0) if row == 0 else (yc1, yb + BAND_H)
                        ydoor = yc0 if row == 0 else yc1
                        for i in range(ROOMS_PER_ROW):
                            x0 = xb + ROOM_W * i
                            x1 = x0 + ROOM_W
                            xc = x0 + ROOM_W // 2
                            kind = KIND_OFFICE
                            if row == 0 and i == 5:
                                kind = KIND_CONF
                            elif row == 1 and i == 0 and side == 0:
                                kind = KIND_REST
                            elif b == 2 and side == 1 and row == 1 and i >= 8:
                                kind = KIND_CAFE
                            name = "room_%d_%d_%d_%d_%d" % (f, b, side, row, i)
                            spaces.append(
                                _rect_space(name, _FUNCTION[kind], x0, y0, x1, y1, [(xc - 3, ydoor, xc + 3, ydoor)])
                            )
                            if f == 0:
                                rx0.append(x0); ry0.append(y0); rx1.append(x1); ry1.append(y1)
                                rdoor.append(xc); rcorr.append((yc0 + yc1) // 2); rkind.append(kind)
            floors.append({"spaces": spaces, "hallway_edges": edges, "minx": 0.0, "maxx": float(FLOOR_W + SPINE_W),
                           "miny": 0.0, "maxy": float(FLOOR_H)})
        self.geometry = {"floors": floors}
        self.room_x0 = np.array(rx0, np.int32); self.room_y0 = np.array(ry0, np.int32)
        self.room_x1 = np.array(rx1, np.int32); self.room_y1 = np.array(ry1, np.int32)
        self.room_door_x = np.array(rdoor, np.int32); self.room_corr_y = np.array(rcorr, np.int32)
        kinds0 = np.array(rkind, np.int32)
        self.room_kind_floor0 = kinds0
        self.room_kind_other = kinds0  # every floor has the same room kinds
        self.n_rooms = len(rx0)

    def rooms_of_kind(self, kind: int, floor0: bool) -> np.ndarray:
        k = self.room_kind_floor0 if floor0 else self.room_kind_other
        return np.nonzero(k == kind)[0].astype(np.int32)

    def lut_expected(self, floor: int) -> Tuple[int, int, np.ndarray]:
        """First-match rectangle rule evaluated directly (tests: must equal the device-built table)."""
        sp = self.geometry["floors"][floor]["spaces"]
        bb = np.array([s["path_bbox"] for s in sp]).astype(np.int64)
        minx, maxx, miny, maxy = bb[:, 0].min(), bb[:, 1].max(), bb[:, 2].min(), bb[:, 3].max()
        lut = np.full((maxy - miny + 1, maxx - minx + 1), -1, np.int16)
        for i in range(len(sp) - 1, -1, -1):
            x0, x1, y0, y1 = bb[i]
            lut[y0 - miny : y1 - miny + 1, x0 - minx : x1 - minx + 1] = i
        return int(minx), int(miny), lut


def _leg(segs: List, t: np.ndarray, x: np.ndarray, y: np.ndarray, fl: np.ndarray, tx: np.ndarray, ty: np.ndarray,
         pace: np.ndarray):
    """Axis-aligned walk from (x,y) to (tx,ty) at `pace` units/step: one segment, whose first
    step is the short one so that the last step lands exactly on the target.  Returns the time
    after the leg."""
    ddx = (tx - x).astype(np.int64)
    ddy = (ty - y).astype(np.int64)
    L = np.abs(ddx) + np.abs(ddy)  # one of them is zero
    n = (L + pace - 1) // pace
    rem = L - (n - 1) * pace  # first step length (== pace when divisible), 0 when n == 0
    sx, sy = np.sign(ddx), np.sign(ddy)
    x0 = x + sx * rem
    y0 = y + sy * rem
    segs.append((t.copy(), n.astype(np.int32), x0.astype(np.int32), y0.astype(np.int32),
                 (sx * pace).astype(np.int16), (sy * pace).astype(np.int16), fl.astype(np.int16)))
    return t + n


def _walk(segs, t, x, y, fl, dax, yca, vx, vy, fl2, ycb, dbx, qx, qy, pace):
    """P -> own door x -> own corridor -> via point -> (floor switch) -> target corridor ->
    target door x -> target y -> target x.  Zero-length legs vanish later."""
    t = _leg(segs, t, x, y, fl, dax, y, pace)
    t = _leg(segs, t, dax, y, fl, dax, yca, pace)
    t = _leg(segs, t, dax, yca, fl, vx, yca, pace)
    t = _leg(segs, t, vx, yca, fl, vx, vy, pace)
    t = _leg(segs, t, vx, vy, fl2, vx, ycb, pace)
    t = _leg(segs, t, vx, ycb, fl2, dbx, ycb, pace)
    t = _leg(segs, t, dbx, ycb, fl2, dbx, qy, pace)
    t = _leg(segs, t, dbx, qy, fl2, qx, qy, pace)
    return t


def _stay(segs, t, x, y, fl, dur):
    z = np.zeros(len(t), np.int16)
    segs.append((t.copy(), dur.astype(np.int32), x.astype(np.int32), y.astype(np.int32), z, z, fl.astype(np.int16)))
    return t + dur


def make_itineraries(
    fac: SyntheticFacility,
    n_agents: int,
    total_timesteps: int,
    n_shifts: int = 1,
    seed: int = 0,
    n_stops: int = 8,
    entry_window: int = 600,
    social: float = 0.6,
    interfloor: float = 0.03,
    batch: int = 1 << 18,
) -> Tuple[np.ndarray, np.ndarray]:
    """(seg_off[N+1] int64, segments[SEGMENT]) for `n_agents` agents over `total_timesteps`.

    `social` is the probability that a stop is a shared destination (conference room,
    restroom, cafeteria, someone else's office) rather than the agent's own desk: it sets the
    pair density (config 5 raises it)."""
    offs = [np.zeros(1, np.int64)]
    chunks = []
    base = 0
    for a0 in range(0, n_agents, batch):
        a1 = min(n_agents, a0 + batch)
        so, sg = _make_batch(fac, a0, a1, n_agents, total_timesteps, n_shifts, seed, n_stops, entry_window, social, interfloor)
        offs.append(so[1:] + base)
        base += int(so[-1])
        chunks.append(sg)
    return np.concatenate(offs), (np.concatenate(chunks) if chunks else np.zeros(0, SEGMENT))


def _make_batch(fac, a0, a1, n_agents, T, n_shifts, seed, n_stops, entry_window, social, interfloor):
    n = a1 - a0
    rng = np.random.Generator(np.random.PCG64([seed, a0]))
    ids = np.arange(a0, a1, dtype=np.int64)
    F = fac.n_floors
    floor = (ids % F).astype(np.int32)
    # agents of one shift have consecutive ids, as in the reference, whose scheduler appends the
    # schedules shift by shift and numbers agents in that order
    # (office_scheduler.py:146-225, simulation.py:277-289)
    shift = np.minimum(ids * n_shifts // max(1, n_agents), n_shifts - 1).astype(np.int64)
    shift_len = T // n_shifts
    offices0, offices1 = fac.rooms_of_kind(KIND_OFFICE, True), fac.rooms_of_kind(KIND_OFFICE, False)

    def pick_room(kind, fl_arr):
        r0 = fac.rooms_of_kind(kind, True)
        r1 = fac.rooms_of_kind(kind, False)
        u = rng.random(n)
        a = r0[(u * len(r0)).astype(np.int64)] if len(r0) else np.zeros(n, np.int32)
        b = r1[(u * len(r1)).astype(np.int64)] if len(r1) else np.zeros(n, np.int32)
        return np.where(fl_arr == 0, a, b)

    def point_in(room):
        ux, uy = rng.random(n), rng.random(n)
        x = fac.room_x0[room] + 2 + (ux * (ROOM_W - 3)).astype(np.int32)
        y = fac.room_y0[room] + 2 + (uy * (ROOM_H - 3)).astype(np.int32)
        return x.astype(np.int32), y.astype(np.int32)

    office = pick_room(KIND_OFFICE, floor)
    desk_x, desk_y = point_in(office)
    # office_schedule.py:423-441: a pace of 2-6 ft/s (normal around 4), divided by the floorplan scale;
    # this facility is drawn like the example one, 12 drawing units to the foot: 24-72 units per step
    pace = (FEET * np.clip(np.rint(rng.normal(4.0, 1.0, n)), 2, 6)).astype(np.int64)
    lane = rng.integers(-9, 10, n).astype(np.int32)  # personal lane in corridors (24 wide) and spine
    spine_x = (SPINE_XC + lane).astype(np.int32)

    segs: List = []
    t = shift * shift_len + rng.integers(0, max(1, entry_window), n)
    t_enter = t.copy()
    # entrance door (spine foot) for one step, then to the desk
    x = spine_x.copy()
    y = np.zeros(n, np.int32)
    fl = floor.copy()
    t = _stay(segs, t, x, y, fl, np.ones(n, np.int64))
    yc = fac.room_corr_y[office] + lane
    t = _walk(segs, t, x, y, fl, x, yc, x, yc, fl, yc, fac.room_door_x[office], desk_x, desk_y, pace)
    x, y = desk_x.copy(), desk_y.copy()
    cur_room = office.copy()
    stay_slots = []
    # weights for distributing the stay budget
    w = rng.random((n_stops + 1, n)) + 0.25
    slot = len(segs)
    segs.append(None)  # first desk stay, filled in below
    stay_slots.append((slot, x.copy(), y.copy(), fl.copy()))
    walk_time = t - t_enter
    t_marks = [t.copy()]
    for k in range(n_stops):
        u = rng.random(n)
        s = social
        go_desk = u >= s
        kind = np.where(u < 0.45 * s, KIND_CONF, np.where(u < 0.65 * s, KIND_REST, np.where(u < 0.85 * s, KIND_CAFE, KIND_OFFICE)))
        # a few per cent of the shared destinations are on another floor (stairs at the spine head)
        other = (rng.random(n) < interfloor) & (F > 1)
        tfl = np.where(other, (fl + 1 + rng.integers(0, max(1, F - 1), n)) % F, fl).astype(np.int32)
        troom = np.empty(n, np.int32)
        for kd in (KIND_CONF, KIND_REST, KIND_CAFE, KIND_OFFICE):
            sel = pick_room(kd, tfl)
            troom = np.where(kind == kd, sel, troom)
        qx, qy = point_in(troom)
        # back to the own desk (own floor) when go_desk
        troom = np.where(go_desk, office, troom)
        tfl = np.where(go_desk, floor, tfl).astype(np.int32)
        qx = np.where(go_desk, desk_x, qx).astype(np.int32)
        qy = np.where(go_desk, desk_y, qy).astype(np.int32)
        dax, yca = fac.room_door_x[cur_room], fac.room_corr_y[cur_room] + lane
        dbx, ycb = fac.room_door_x[troom], fac.room_corr_y[troom] + lane
        same_room = (troom == cur_room) & (tfl == fl)
        same_corr = (yca == ycb) & ((dax < SPINE_XC) == (dbx < SPINE_XC)) & (tfl == fl)
        vx = np.where(same_corr, dbx, spine_x).astype(np.int32)
        vy = np.where(tfl == fl, ycb, STAIRS_Y + lane).astype(np.int32)
        # staying in the same room: walk straight (x then y) without leaving
        dax = np.where(same_room, x, dax).astype(np.int32)
        yca_e = np.where(same_room, y, yca).astype(np.int32)
        vx = np.where(same_room, x, vx).astype(np.int32)
        vy = np.where(same_room, y, vy).astype(np.int32)
        ycb_e = np.where(same_room, y, ycb).astype(np.int32)
        dbx_e = np.where(same_room, x, dbx).astype(np.int32)
        t0w = t.copy()
        t = _walk(segs, t, x, y, fl, dax, yca_e, vx, vy, tfl, ycb_e, dbx_e, qx, qy, pace)
        walk_time += t - t0w
        x, y, fl, cur_room = qx, qy, tfl, troom
        slot = len(segs)
        segs.append(None)
        stay_slots.append((slot, x.copy(), y.copy(), fl.copy()))
        t_marks.append(t.copy())
    # exit: to the spine foot of the current floor, then outside (PARK_Y < 0: no space there)
    dax, yca = fac.room_door_x[cur_room], fac.room_corr_y[cur_room] + lane
    sx = spine_x
    py = np.full(n, PARK_Y, np.int32)
    t0w = t.copy()
    t_exit = _walk(segs, t, x, y, fl, dax, yca, sx, py, fl, py, sx, sx, py, pace)
    walk_time += t_exit - t0w
    # stay budget: the shift minus all walking, spread over the stays by random weights
    budget = np.maximum(int(0.92 * shift_len) - entry_window - walk_time, 30 * (n_stops + 1))
    wsum = w.sum(axis=0)
    stays = np.maximum(30, (w / wsum * budget).astype(np.int64))
    # second pass: shift every later segment by the cumulative stays
    cum = np.zeros(n, np.int64)
    out = []
    k = 0
    for i, sg in enumerate(segs):
        if sg is None:
            _, sx_, sy_, sfl_ = stay_slots[k]
            z = np.zeros(n, np.int16)
            out.append((t_marks[k] + cum, stays[k].astype(np.int32), sx_, sy_, z, z, sfl_.astype(np.int16)))
            cum = cum + stays[k]
            k += 1
        else:
            out.append((sg[0] + cum,) + sg[1:])
    # terminal hold (sticky tail) at the parking point
    z = np.zeros(n, np.int16)
    out.append((t_exit + cum, np.ones(n, np.int32), sx, py, z, z, fl.astype(np.int16)))
    L = len(out)
    t0 = np.stack([o[0] for o in out], axis=1)          # (n, L)
    dur = np.stack([o[1] for o in out], axis=1)
    keep = dur > 0
    # merge a stay that follows a zero-velocity arrival at the same point? not needed: legs move
    rec = np.zeros((n, L), SEGMENT)
    rec["t0"] = t0
    rec["x0"] = np.stack([o[2] for o in out], axis=1)
    rec["y0"] = np.stack([o[3] for o in out], axis=1)
    rec["dx"] = np.stack([o[4] for o in out], axis=1)
    rec["dy"] = np.stack([o[5] for o in out], axis=1)
    rec["floor"] = np.stack([o[6] for o in out], axis=1)
    keep &= t0 < T  # segments starting at or after T are never read ...
    keep[:, L - 1] = True  # ... except the terminal hold: every agent's last segment has zero velocity (the sticky tail)
    cnt = keep.sum(axis=1)
    seg_off = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=seg_off[1:])
    return seg_off, rec[keep]


def expand_to_reference_format(seg_off: np.ndarray, segs: np.ndarray, total_timesteps: int):
    """Segments -> the reference's per-step itinerary lists `[((x, y) | None, floor | None), ...]`
    (office_schedule.py:108-110), for feeding the oracle / the real reference on small cases."""
    out = []
    T = int(total_timesteps)
    for a in range(len(seg_off) - 1):
        s = segs[seg_off[a] : seg_off[a + 1]]
        it = []
        if len(s) == 0:
            out.append([(None, None)])
            continue
        it.extend([(None, None)] * int(s["t0"][0]))
        for i in range(len(s)):
            t0 = int(s["t0"][i])
            t1 = int(s["t0"][i + 1]) if i + 1 < len(s) else max(t0 + 1, T)
            x0, y0, dx, dy, f = int(s["x0"][i]), int(s["y0"][i]), int(s["dx"][i]), int(s["dy"][i]), int(s["floor"][i])
            if i + 1 == len(s):
                # a zero-velocity tail is sticky (one entry is enough); a moving tail (later
                # segments start at or after T and were dropped) runs to the horizon
                t1 = t0 + 1 if (dx == 0 and dy == 0) else max(t0 + 1, T)
            for t in range(t0, t1):
                it.append(((x0 + dx * (t - t0), y0 + dy * (t - t0)), f))
        out.append(it)
    return out


CONFIGS = {
    # BASELINE.json configs[1..4]; config 4 (replica sweep) is built from `config2`-like replicas
    "config2": dict(n_floors=1, n_agents=10_000, total_timesteps=28_800, n_shifts=1, social=0.6),
    # three consecutive 2 h 40 min shifts (offsets 0, T/3, 2T/3).  Agents arrive within the
    # first 10 minutes of their shift and stay until its end (presence ~8 200 of 28 800 steps: a third of the
    # agents is in the building at any time); 4 destinations per shift is the 10-items-per-8-hours rate of the
    # survey's generator spec.  Walks are straight axis-aligned legs (6-8 per walk, ~45 segments per agent)
    # rather than the reference's node-by-node routes (~100 runs per agent).
    "config3": dict(n_floors=10, n_agents=1_000_000, total_timesteps=28_800, n_shifts=3, social=0.6, entry_window=600,
                    n_stops=4),
    "config5": dict(n_floors=1, n_agents=200_000, total_timesteps=28_800, n_shifts=1, social=0.85),
    "tiny": dict(n_floors=2, n_agents=300, total_timesteps=2400, n_shifts=2, social=0.7),
}


def activity(seg_off: np.ndarray, segs: np.ndarray, total_timesteps: int) -> Dict[str, float]:
    """Shape of a store: the share of the N x T agent-steps at which the agent is in the building
    (between its first segment and its terminal hold), the share at which it moves, and the segments
    per agent.  bench.py prints these beside the headline (which counts every agent-step)."""
    n = len(seg_off) - 1
    T = int(total_timesteps)
    if len(segs) == 0 or n == 0:
        return {"active_fraction": 0.0, "moving_fraction": 0.0, "segments_per_agent": 0.0}
    t0 = np.minimum(segs["t0"].astype(np.int64), T)
    nxt = np.empty_like(t0)
    nxt[:-1] = t0[1:]
    last = seg_off[1:][seg_off[1:] > seg_off[:-1]] - 1
    nxt[last] = T
    dur = np.maximum(nxt - t0, 0)
    is_last = np.zeros(len(segs), bool)
    is_last[last] = True
    moving = (segs["dx"] != 0) | (segs["dy"] != 0)
    present = dur[~is_last].sum()  # the terminal hold is the parking point outside every space
    return {"active_fraction": float(present) / (n * T), "moving_fraction": float(dur[moving].sum()) / (n * T),
            "segments_per_agent": len(segs) / n}


def make_config(name: str, seed: int = 0, **override):
    """(facility, seg_off, segs, params) for a named BASELINE shape."""
    p = dict(CONFIGS[name])
    p.update(override)
    fac = SyntheticFacility(p["n_floors"])
    seg_off, segs = make_itineraries(
        fac, p["n_agents"], p["total_timesteps"], n_shifts=p["n_shifts"], seed=seed, social=p["social"],
        n_stops=p.get("n_stops", 8), entry_window=p.get("entry_window", 600), interfloor=p.get("interfloor", 0.03),
    )
    return fac, seg_off, segs, p
