"""Itinerary store: the reference's per-agent itinerary lists as constant-velocity segments.

The reference keeps, per agent, a Python list with one entry per time step:
`((x, y) | None, floor | None)` (office_schedule.py:108-110, 467-593) and reads it with
`get_next_position` (office_schedule.py:632-638): entry `min(t, len-1)` at step t, where a
`None` entry leaves the agent where it was (agent.py:77-107).  More than 99 % of the entries
repeat the previous position, so the device store is a CSR list of `citam_segment` records
(include/citam_b200.h): position `(x0 + dx*(t-t0), y0 + dy*(t-t0))` on `floor` from `t0`
until the next segment's `t0`; the last segment of an agent has zero velocity and lasts for
ever (the sticky tail).  Encoding is lossless for any integer itinerary.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

from ._lib import SEGMENT

MAX_STEP = (1 << 24) - 1
MAX_FLOOR = 62


def pack(itineraries: Sequence[Sequence]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Reference-format lists -> flat arrays (off[N+1], x, y, floor) with floor=-1 for None.

    This is the way in for the reference's own `OfficeSchedule.itinerary` lists
    (office_schedule.py:108-110, 467-593).  No Python statement runs per entry: each list becomes an
    object array of its entries (`np.fromiter`), ONE elementwise comparison finds where an entry
    differs from the one before (more than 99 % of the entries repeat it), only the entries that
    start a run are unpacked and checked, and the dense columns are rebuilt with `np.repeat`
    (~55 ns per entry instead of ~400)."""
    n = len(itineraries)
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(it) for it in itineraries], out=off[1:])
    total = int(off[-1])
    x = np.zeros(total, np.int64)
    y = np.zeros(total, np.int64)
    fl = np.full(total, -1, np.int64)
    for i, it in enumerate(itineraries):
        L = len(it)
        if L == 0:
            continue
        lo = int(off[i])
        ent = np.fromiter(it, object, L)
        start = np.ones(L, bool)  # entry k starts a run: it differs from entry k-1
        if L > 1:
            start[1:] = ent[1:] != ent[:-1]
        idx = np.flatnonzero(start)
        runs = np.diff(np.append(idx, L))
        _, hx, hy, hf = pack_entrywise([ent[idx].tolist()])  # the run heads, checked like any entry
        x[lo : lo + L] = np.repeat(hx, runs)
        y[lo : lo + L] = np.repeat(hy, runs)
        fl[lo : lo + L] = np.repeat(hf, runs)
    return off, x, y, fl


def pack_entrywise(itineraries: Sequence[Sequence]) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """The same, one Python statement per entry: the plain statement of the format.  `pack` applies it
    to the first entry of every run; tests compare the two on whole itineraries."""
    n = len(itineraries)
    off = np.zeros(n + 1, np.int64)
    for i, it in enumerate(itineraries):
        off[i + 1] = off[i] + len(it)
    total = int(off[-1])
    x = np.zeros(total, np.int64)
    y = np.zeros(total, np.int64)
    fl = np.full(total, -1, np.int64)
    k = 0
    for it in itineraries:
        for xy, f in it:
            if xy is not None:
                xi, yi = xy[0], xy[1]
                if xi != int(xi) or yi != int(yi):
                    raise ValueError(
                        "itinerary coordinates must be integers (got %r); the device store is an "
                        "integer lattice like the reference's navigation network" % (xy,)
                    )
                if f is None or int(f) < 0:
                    raise ValueError("itinerary entry with a position but no floor: %r" % ((xy, f),))
                x[k], y[k], fl[k] = int(xi), int(yi), int(f)
            k += 1
    return off, x, y, fl


def encode_packed(off: np.ndarray, x: np.ndarray, y: np.ndarray, fl: np.ndarray):
    """Flat per-step arrays -> (seg_off[N+1] int64, segments[SEGMENT]).  Vectorised."""
    off = np.asarray(off, np.int64)
    n = len(off) - 1
    total = int(off[-1])
    if total == 0:
        return np.zeros(n + 1, np.int64), np.zeros(0, SEGMENT)
    x = np.asarray(x, np.int64)
    y = np.asarray(y, np.int64)
    fl = np.asarray(fl, np.int64)
    lens = np.diff(off)
    agent = np.repeat(np.arange(n, dtype=np.int64), lens)
    start = off[agent]
    k = np.arange(total, dtype=np.int64)
    valid = fl >= 0
    # sticky forward fill inside each agent
    src = np.maximum.accumulate(np.where(valid, k, -1))
    active = src >= start
    src = np.where(active, src, 0)
    ex, ey, ef = x[src], y[src], fl[src]
    is_last = k == (off[agent + 1] - 1)
    prev_ok = np.zeros(total, bool)  # k-1 belongs to the same agent and is active
    prev_ok[1:] = (agent[1:] == agent[:-1]) & active[:-1]
    prev_ok &= active
    dxk = np.zeros(total, np.int64)  # d_k = e_k - e_{k-1}
    dyk = np.zeros(total, np.int64)
    dxk[1:] = ex[1:] - ex[:-1]
    dyk[1:] = ey[1:] - ey[:-1]
    same_floor = np.zeros(total, bool)
    same_floor[1:] = ef[1:] == ef[:-1]
    cont = prev_ok & same_floor  # k continues k-1 geometrically
    # velocity leaving k (defined when k+1 continues k)
    nxt_cont = np.zeros(total, bool)
    nxt_cont[:-1] = cont[1:]
    vx = np.zeros(total, np.int64)
    vy = np.zeros(total, np.int64)
    vx[:-1] = np.where(nxt_cont[:-1], dxk[1:], 0)
    vy[:-1] = np.where(nxt_cont[:-1], dyk[1:], 0)
    # k starts a segment if it does not continue k-1, the velocity changes at k, or it is the
    # agent's last entry while moving (terminal hold)
    vel_change = cont & nxt_cont & ((vx != dxk) | (vy != dyk))
    stop = cont & ~nxt_cont & ((dxk != 0) | (dyk != 0))  # arrived moving, nothing follows
    brk = active & (~cont | vel_change | stop)
    idx = np.nonzero(brk)[0]
    seg_agent = agent[idx]
    t0 = idx - off[seg_agent]
    seg = np.zeros(len(idx), SEGMENT)
    sx, sy, sf = ex[idx], ey[idx], ef[idx]
    svx, svy = vx[idx], vy[idx]
    # absorb a segment into its predecessor when the predecessor's line already passes
    # through it with the same velocity (single points created by the break rule)
    if len(idx) > 1:
        same_agent = np.zeros(len(idx), bool)
        same_agent[1:] = seg_agent[1:] == seg_agent[:-1]
        pdt = np.zeros(len(idx), np.int64)
        pdt[1:] = idx[1:] - idx[:-1]
        on_line = np.zeros(len(idx), bool)
        on_line[1:] = (
            (sx[1:] == sx[:-1] + svx[:-1] * pdt[1:])
            & (sy[1:] == sy[:-1] + svy[:-1] * pdt[1:])
            & (sf[1:] == sf[:-1])
        )
        same_vel = np.zeros(len(idx), bool)
        same_vel[1:] = (svx[1:] == svx[:-1]) & (svy[1:] == svy[:-1])
        # a one-step segment that is not the agent's last one has no velocity of its own
        single = np.zeros(len(idx), bool)
        single[:-1] = (seg_agent[1:] == seg_agent[:-1]) & (idx[1:] == idx[:-1] + 1)
        absorb = same_agent & on_line & cont[idx] & (same_vel | single)
        # only absorb into a predecessor that is itself kept
        prev_absorbed = np.zeros(len(idx), bool)
        prev_absorbed[1:] = absorb[:-1]
        keep = ~(absorb & ~prev_absorbed)
        idx, seg_agent, t0 = idx[keep], seg_agent[keep], t0[keep]
        sx, sy, sf, svx, svy = sx[keep], sy[keep], sf[keep], svx[keep], svy[keep]
        seg = np.zeros(len(idx), SEGMENT)
    if len(idx):
        if t0.max() > MAX_STEP:
            raise ValueError("itinerary longer than %d steps" % MAX_STEP)
        if sf.max() > MAX_FLOOR:
            raise ValueError("floor index above %d" % MAX_FLOOR)
        if max(abs(svx).max(), abs(svy).max()) > 32767:
            raise ValueError("per-step displacement does not fit int16")
        if max(abs(sx).max(), abs(sy).max()) >= (1 << 26):
            raise ValueError("coordinates must stay below 2^26 in magnitude")
    seg["t0"], seg["x0"], seg["y0"] = t0, sx, sy
    seg["dx"], seg["dy"], seg["floor"] = svx, svy, sf
    seg_off = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(seg_agent, minlength=n), out=seg_off[1:])
    # the tail must hold still: a moving last segment gets a zero-velocity successor
    last = seg_off[1:] - 1
    has = seg_off[1:] > seg_off[:-1]
    moving_tail = np.zeros(n, bool)
    moving_tail[has] = (seg["dx"][last[has]] != 0) | (seg["dy"][last[has]] != 0)
    if moving_tail.any():  # cannot happen with the break rule above; kept as a guard
        raise AssertionError("encoder produced a moving tail")
    return seg_off, seg


def encode(itineraries: Sequence[Sequence]):
    """Reference-format itinerary lists -> (seg_off, segments)."""
    return encode_packed(*pack(itineraries))


def decode(seg_off: np.ndarray, seg: np.ndarray, agent: int, t: int):
    """Host evaluation of one agent's state at step t: (x, y, floor) or None (tests, writers)."""
    a, b = int(seg_off[agent]), int(seg_off[agent + 1])
    if a == b:
        return None
    t0 = seg["t0"][a:b]
    c = int(np.searchsorted(t0, t, side="right")) - 1
    if c < 0:
        return None
    s = seg[a + c]
    dt = t - int(s["t0"])
    return int(s["x0"]) + int(s["dx"]) * dt, int(s["y0"]) + int(s["dy"]) * dt, int(s["floor"])


def dense_states(seg_off: np.ndarray, seg: np.ndarray, t: int):
    """All agents at step t as arrays (x, y, floor) with floor=-1 for inactive (host, tests)."""
    n = len(seg_off) - 1
    x = np.zeros(n, np.int64)
    y = np.zeros(n, np.int64)
    f = np.full(n, -1, np.int64)
    for a in range(n):
        r = decode(seg_off, seg, a, t)
        if r is not None:
            x[a], y[a], f[a] = r
    return x, y, f


def slice_segments(seg_off: np.ndarray, seg: np.ndarray, t_lo: int, t_hi: int):
    """Sub-store that reproduces every agent's state for steps in [t_lo, t_hi): per agent the
    last segment starting at or before t_lo (if any) and all segments starting before t_hi.
    Used to stream one time block's itinerary slice to a device (time-block sharding)."""
    seg_off = np.asarray(seg_off, np.int64)
    n = len(seg_off) - 1
    if len(seg) == 0:
        return seg_off.copy(), seg.copy()
    t0 = seg["t0"].astype(np.int64)
    c_lo = np.concatenate(([0], np.cumsum(t0 <= t_lo)))  # segments with t0 <= t_lo before index i
    c_hi = np.concatenate(([0], np.cumsum(t0 < t_hi)))
    n_lo = c_lo[seg_off[1:]] - c_lo[seg_off[:-1]]
    n_hi = c_hi[seg_off[1:]] - c_hi[seg_off[:-1]]
    first = seg_off[:-1] + np.maximum(n_lo - 1, 0)
    last = seg_off[:-1] + n_hi  # exclusive
    cnt = np.maximum(last - first, 0)
    new_off = np.zeros(n + 1, np.int64)
    np.cumsum(cnt, out=new_off[1:])
    idx = np.repeat(first - new_off[:-1], cnt) + np.arange(int(new_off[-1]), dtype=np.int64)
    return new_off, np.ascontiguousarray(seg[idx])
