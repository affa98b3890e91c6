"""GPU parity at BASELINE.json's full sizes.

config 2 (10^4 agents, one floor, the whole 28 800-step day) is compared table for table with the C
oracle.  config 3 (10^6 agents, 10 floors) is compared with the C oracle on a 160-step window and on
2 400 contiguous steps of the day (many chunks, contact runs crossing chunk borders, both step
pipelines), config 5 (2 x 10^5 agents crowded on one floor) on a 96-step window, and config 3 is
checked through size-independent properties: conservation (every contact-step is in exactly one
pair, two agents' counters and one coordinate bucket), additivity over time blocks run on different
engines and merged, and independence of the chunking."""
import numpy as np
import pytest

from oracle_util import grid_oracle_for

pytestmark = pytest.mark.gpu


def _arrays(obj):
    p, c = obj.pairs(), obj.coords()
    return p, c, np.asarray(obj.agent_durations()).astype(np.int64)


def _same(a, b):
    for x, y in zip(a[:2], b[:2]):
        assert len(x) == len(y)
        for f in x.dtype.names:
            if f != "reserved":
                assert (x[f] == y[f]).all(), f
    assert (a[2] == b[2]).all()


def test_config2_whole_day_equals_oracle():
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    fac, seg_off, segs, p = S.make_config("config2", seed=0)
    T = p["total_timesteps"]
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(0, T, 16)
    eng = DeviceEngine(p["n_agents"], 1, 6.0, pair_capacity=1 << 25)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(0, T)
    got, want = _arrays(eng), _arrays(g)
    _same(got, want)
    c = eng.counters()
    assert c["contact_steps"] == g.contact_steps == int(want[0]["duration"].sum())
    eng.close()


def test_config3_million_agents_window_and_invariants():
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    fac, seg_off, segs, p = S.make_config("config3", seed=0)
    N, t0, t1 = p["n_agents"], 12_000, 12_160
    eng = DeviceEngine(N, p["n_floors"], 6.0, pair_capacity=1 << 27)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(t0, t1)
    got = _arrays(eng)
    c = eng.counters()
    # conservation
    assert int(got[0]["duration"].sum()) == c["contact_steps"]
    assert int(got[2].sum()) == 2 * c["contact_steps"]
    assert int(got[1]["count"].sum()) == c["contact_steps"]
    assert (got[0]["a1"] < got[0]["a2"]).all() and int(got[0]["a2"].max()) < N
    assert (got[0]["n_events"] <= got[0]["duration"]).all()
    # an event that was already running one step before the window starts no event inside it
    assert (got[0]["first_step"][got[0]["n_events"] == 0] == t0).all()
    assert ((got[0]["first_step"] >= t0) & (got[0]["first_step"] < t1)).all()
    order = np.lexsort((got[0]["a2"], got[0]["a1"], got[0]["first_step"]))
    assert (order == np.arange(len(order))).all()  # first-contact order
    # the C oracle on the same window
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(t0, t1, 16)
    _same(got, _arrays(g))
    # additivity over time blocks on different engines + independence of the chunking
    a = DeviceEngine(N, p["n_floors"], 6.0, pair_capacity=1 << 27, steps_per_chunk=23)
    a.load_geometry(fac.geometry)
    a.load_segments(seg_off, segs)
    mid = t0 + 71
    a.run(mid, t1)
    pa, ca, da = a.pairs(), a.coords(), a.agent_durations()
    a.reset()
    a.run(t0, mid)
    a.merge(pa, ca)  # pair table + histogram; per-agent seconds are reduced separately
    merged = _arrays(a)
    _same((merged[0], merged[1], merged[2] + da.astype(np.int64)), got)
    a.close()
    eng.close()


def test_config3_million_agents_2400_contiguous_steps_equal_oracle(monkeypatch):
    """10^6 agents through 2 400 contiguous steps in the middle of the first shift: ~19 chunks of the
    split pipeline, contact runs that span them, against the C oracle table for table."""
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    fac, seg_off, segs, p = S.make_config("config3", seed=0)
    N, t0, t1 = p["n_agents"], 3_000, 5_400
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(t0, t1, 16)
    want = _arrays(g)
    eng = DeviceEngine(N, p["n_floors"], 6.0, pair_capacity=1 << 28)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(t0, t1)
    got = _arrays(eng)
    _same(got, want)
    assert eng.counters()["contact_steps"] == g.contact_steps
    sums = eng.statistics_sums()
    assert sums["n_pairs"] == len(want[0]) and sums["total_duration"] == int(want[0]["duration"].sum())
    eng.close()


def test_config5_dense_floor_window_equals_oracle():
    """2 x 10^5 agents on one floor, 85 % of them bound for shared destinations (tens of neighbours
    within the contact distance each): a 96-step window against the C oracle."""
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    fac, seg_off, segs, p = S.make_config("config5", seed=0)
    N, t0, t1 = p["n_agents"], 9_000, 9_096
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(t0, t1, 16)
    want = _arrays(g)
    eng = DeviceEngine(N, p["n_floors"], 6.0, pair_capacity=1 << 28)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    eng.run(t0, t1)
    got = _arrays(eng)
    _same(got, want)
    assert eng.counters()["contact_steps"] == g.contact_steps
    eng.close()
