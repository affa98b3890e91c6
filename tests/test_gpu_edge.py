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
