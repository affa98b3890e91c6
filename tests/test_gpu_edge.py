"""GPU edge cases against the faithful Python port (small, hand-made inputs)."""
import math

import numpy as np
import pytest

from oracle_util import run_port, tables_of, tables_of_port

pytestmark = pytest.mark.gpu


def _fac():
    from citam_b200 import synthetic as S

    return S.SyntheticFacility(1)


def _run_both(itins, T, D, **kw):
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    eng = DeviceEngine(len(itins), 1, D, **kw)
    eng.load_geometry(fac.geometry)
    eng.load_itineraries(itins)
    eng.run(0, T)
    got = tables_of(eng)
    stats = eng.statistics_sums()
    eng.close()
    sim = run_port(itins, fac.geometry, D, T)
    want = tables_of_port(sim, 1)
    return got, (want[0], want[1], want[2]), stats


def test_no_itineraries_at_all():
    itins = [[(None, None)] for _ in range(5)]
    got, want, stats = _run_both(itins, 50, 6.0)
    assert got == ([], {}, [0] * 5) and got == want
    assert stats["n_pairs"] == 0 and stats["n_agents_with_contacts"] == 0


def test_single_agent_and_zero_steps():
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    eng = DeviceEngine(1, 1, 6.0)
    eng.load_geometry(fac.geometry)
    eng.load_itineraries([[((30, 100), 0)] * 10])
    eng.run(0, 0)
    eng.run(0, 10)
    assert tables_of(eng) == ([], {}, [0])
    with pytest.raises(ValueError):
        eng.run(5, 3)
    eng.close()


def test_clique_of_colocated_agents_and_late_arrivals():
    """The reference's parked-together blow-up: K co-located agents make K(K-1)/2 pairs per step."""
    K, T = 40, 30
    itins = []
    for a in range(K):
        start = a % 7
        itins.append([(None, None)] * start + [((100, 100), 0)] * (T - start))
    itins.append([((103, 104), 0)] * T)  # distance exactly 5
    itins.append([((100, 106), 0)] * T)  # distance exactly 6: not a contact (strict <)
    got, want, _ = _run_both(itins, T, 6.0, steps_per_chunk=4)
    assert got == want
    assert len(got[0]) == K * (K - 1) // 2 + K + 1  # clique + (each, the 5-away agent) + the two outsiders


@pytest.mark.parametrize("D", [0.0, 0.5, 1.0, 6.0, 6.0000001, math.sqrt(72.0), np.nextafter(math.sqrt(72.0), 10.0), 9.5, 17.25])
def test_distance_threshold_is_the_float64_predicate(D):
    """Integer threshold s_max vs scipy's sqrt(dx^2+dy^2) < D on a cloud of static points."""
    rng = np.random.RandomState(3)
    n, T = 120, 6
    xs = rng.randint(20, 50, n)
    ys = rng.randint(70, 100, n)  # inside one room of the synthetic floor
    itins = [[((int(x), int(y)), 0)] * T for x, y in zip(xs, ys)]
    got, want, _ = _run_both(itins, T, float(D))
    assert got == want


def test_lattice_edges_outside_points_and_cell_borders():
    from citam_b200 import synthetic as S

    T = 12
    W, H = S.FLOOR_W + S.SPINE_W, S.FLOOR_H
    pts = [(600, 0), (600, 1), (612, 1200), (611, 1199), (588, 1200), (1200, 168), (1199, 170), (0, 168), (1, 166),
           (600, -30), (600, -1), (5000, 5000), (-7, 3), (594, 6), (594, 5), (600, 6), (606, 11), (606, 12)]
    itins = [[(p, 0)] * T for p in pts]
    # two walkers crossing cell borders along the spine, one step apart
    itins.append([((600, 2 * t), 0) for t in range(T)])
    itins.append([((601, 2 * t + 3), 0) for t in range(T)])
    got, want, _ = _run_both(itins, T, 6.0, steps_per_chunk=5)
    assert got == want
    assert len(got[0]) > 0


def test_events_break_and_resume_across_chunk_borders():
    """Stationary runs that start/stop at every offset relative to the chunk size."""
    T = 64
    a = [((40, 90), 0)] * T
    itins = [a]
    for k in range(9):  # partner k is near for steps [k, k+5) and [k+20, T), elsewhere far
        it = []
        for t in range(T):
            near = (k <= t < k + 5) or t >= k + 20
            it.append(((42, 91) if near else (20, 75), 0))
        itins.append(it)
    for chunk in (1, 3, 7, 64):
        got, want, _ = _run_both(itins, T, 6.0, steps_per_chunk=chunk)
        assert got == want, chunk
    ev = {(r[0], r[1]): r[3] for r in got[0]}
    assert all(ev[(0, k + 1)] == 2 for k in range(9))


def test_plugin_steps_with_a_gap_start_a_new_event():
    """Calculator.run semantics through citam_step_states: non-consecutive steps never continue an event."""
    from citam_b200 import _lib as L
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    eng = DeviceEngine(3, 1, 6.0)
    eng.load_geometry(fac.geometry)
    st = np.zeros(3, L.AGENT_STATE)
    st["x"], st["y"], st["floor"] = [40, 42, 400], [90, 91, 400], 0
    st["loc"] = [eng.locate(0, np.array([x]), np.array([y]))[0] for x, y in zip(st["x"], st["y"])]
    for t in (0, 1, 2, 5, 6, 9):
        eng.step_states(t, st)
    p = eng.pairs()
    assert [(int(r["a1"]), int(r["a2"]), int(r["first_step"]), int(r["n_events"]), int(r["duration"])) for r in p] == [(0, 1, 0, 3, 6)]
    assert eng.agent_durations().tolist() == [6, 6, 0]
    eng.close()


def test_capacity_errors_are_reported_not_silent():
    from citam_b200._lib import CitamCapacityError
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    K, T = 300, 4
    itins = [[((100 + (a % 3), 100), 0)] * T for a in range(K)]  # 44850 pairs
    eng = DeviceEngine(K, 1, 6.0, pair_capacity=1024)
    eng.load_geometry(fac.geometry)
    eng.load_itineraries(itins)
    eng.run(0, T)
    with pytest.raises(CitamCapacityError):
        eng.pairs()
    eng.close()
    # the library-sized table grows instead
    eng = DeviceEngine(K, 1, 6.0)
    eng.load_geometry(fac.geometry)
    eng.load_itineraries(itins)
    eng.run(0, T)
    assert len(eng.pairs()) == K * (K - 1) // 2
    eng.close()


def test_one_table_update_per_contact_run_whatever_the_chunking():
    """A stationary pair's contact run is ONE table update per citam_run range: the number of
    updates does not depend on how many chunks the range is cut into, and the results do not either."""
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    T = 900
    itins = []
    for a in range(12):  # six pairs sitting side by side in an office for most of the run
        x, y = 40 + 3 * (a // 2), 80 + 2 * (a % 2)
        itins.append([(None, None)] * (5 + a) + [((x, y), 0)] * (T - 20 - 2 * a) + [(None, None)])
    itins.append([((30 + (t % 40), 82), 0) for t in range(T)])  # one walker sweeping past them
    runs, tabs = [], []
    for chunk in (7, 64, 0):
        eng = DeviceEngine(len(itins), 1, 6.0, steps_per_chunk=chunk)
        eng.load_geometry(fac.geometry)
        eng.load_itineraries(itins)
        eng.run(0, T)
        runs.append(eng.counters()["contact_runs"])
        tabs.append(tables_of(eng))
        eng.close()
    assert runs[0] == runs[1] == runs[2], runs
    assert tabs[0] == tabs[1] == tabs[2]
    sim = run_port(itins, fac.geometry, 6.0, T)
    want = tables_of_port(sim, 1)
    assert tabs[0] == (want[0], want[1], want[2])
    # far fewer updates than contact-steps
    assert runs[0] * 5 < sum(r[4] for r in want[0])


def test_load_itineraries_rejects_stores_outside_the_documented_limits():
    """include/citam_b200.h: the limits are checked on the device copy (CITAM_ERR_INVALID)."""
    from citam_b200 import _lib as L
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    eng = DeviceEngine(3, 1, 6.0)
    eng.load_geometry(fac.geometry)

    def store(rows, off):
        s = np.zeros(len(rows), L.SEGMENT)
        for i, r in enumerate(rows):
            s[i] = r + (0,)
        return np.array(off, np.int64), s

    good = [(0, 30, 100, 0, 0, 0), (5, 30, 100, 1, 0, 0), (9, 34, 100, 0, 0, 0), (0, 50, 100, 0, 0, 0), (2, 60, 100, 0, 0, 0)]
    eng.load_segments(*store(good, [0, 3, 4, 5]))
    eng.run(0, 20)
    bad_cases = {
        "seg_off decreases": (good, [0, 4, 3, 5]),
        "t0 not increasing": ([(0, 30, 100, 0, 0, 0), (5, 30, 100, 1, 0, 0), (5, 34, 100, 0, 0, 0)] + good[3:], [0, 3, 4, 5]),
        "negative t0": ([(-1, 30, 100, 0, 0, 0)] + good[1:], [0, 3, 4, 5]),
        "t0 beyond 2^24": ([(0, 30, 100, 0, 0, 0), (5, 30, 100, 1, 0, 0), (1 << 24, 34, 100, 0, 0, 0)] + good[3:], [0, 3, 4, 5]),
        "floor 63": (good[:4] + [(2, 60, 100, 0, 0, 63)], [0, 3, 4, 5]),
        "negative floor": (good[:4] + [(2, 60, 100, 0, 0, -1)], [0, 3, 4, 5]),
        "moving last segment": (good[:2] + [(9, 34, 100, 0, 2, 0)] + good[3:], [0, 3, 4, 5]),
        "coordinate beyond 2^26": (good[:4] + [(2, 1 << 26, 100, 0, 0, 0)], [0, 3, 4, 5]),
    }
    for name, (rows, off) in bad_cases.items():
        with pytest.raises(ValueError):
            eng.load_segments(*store(rows, off))
        with pytest.raises(L.CitamDeviceError):
            eng.run(0, 5)  # a rejected store leaves the engine without itineraries
    eng.load_segments(*store(good, [0, 3, 4, 5]))
    with pytest.raises(ValueError):
        eng.run(0, 1 << 24)  # steps < 2^24 - 1
    eng.run(20, 30)
    eng.close()


def test_merge_grows_a_library_sized_table_and_rejects_malformed_records():
    from citam_b200 import _lib as L
    from citam_b200.engine import DeviceEngine

    fac = _fac()
    N = 6000
    eng = DeviceEngine(N, 1, 6.0)
    eng.load_geometry(fac.geometry)
    rng = np.random.default_rng(3)
    n = 5_000_000  # more than the library-sized table (2^22 slots) holds
    a1 = rng.integers(0, N - 1, n).astype(np.uint32)
    a2 = (a1 + 1 + rng.integers(0, N - 1 - a1)).astype(np.uint32)
    rec = np.zeros(n, L.PAIR)
    rec["a1"], rec["a2"], rec["first_step"], rec["n_events"], rec["duration"] = a1, a2, rng.integers(0, 1000, n), 1, 2
    eng.merge(rec, np.zeros(0, L.COORD))
    key = a1.astype(np.uint64) << np.uint64(32) | a2
    uk, first_idx, cnt = np.unique(key, return_index=True, return_counts=True)
    p = eng.pairs()
    assert len(p) == len(uk)
    assert int(p["duration"].sum()) == 2 * n and int(p["n_events"].sum()) == n
    got_key = p["a1"].astype(np.uint64) << np.uint64(32) | p["a2"]
    o = np.argsort(got_key)
    assert (got_key[o] == uk).all() and (p["duration"][o] == 2 * cnt).all()
    fmin = np.full(len(uk), 1 << 30)
    np.minimum.at(fmin, np.searchsorted(uk, key), rec["first_step"])
    assert (p["first_step"][o] == fmin).all()
    assert (np.lexsort((p["a2"], p["a1"], p["first_step"])) == np.arange(len(p))).all()  # device sort order
    assert eng.agent_durations().sum() == 0  # merging does not touch the per-agent counters
    bad = rec[:4].copy()
    bad["a2"][2] = bad["a1"][2]
    with pytest.raises(ValueError):
        eng.merge(bad, np.zeros(0, L.COORD))
    eng.close()
