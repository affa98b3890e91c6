"""GPU parity on synthetic facilities (BASELINE shapes at reduced size) against the C grid oracle,
itself pinned to the reference by tests/test_oracle_golden.py.  Bit-exact integer tables."""
import numpy as np
import pytest

from oracle_util import grid_oracle_for, tables_of

pytestmark = pytest.mark.gpu

CASES = {
    "tiny": dict(name="tiny", over={}, T=2400),
    "config2_small": dict(name="config2", over=dict(n_agents=3000, total_timesteps=3000), T=3000),
    "config3_small": dict(name="config3", over=dict(n_agents=9000, total_timesteps=2400, n_floors=3, entry_window=300), T=2400),
    "config5_small": dict(name="config5", over=dict(n_agents=3000, total_timesteps=1500), T=1500),
}


@pytest.fixture(scope="module", params=list(CASES))
def synth(request):
    from citam_b200 import synthetic as S

    c = CASES[request.param]
    fac, seg_off, segs, p = S.make_config(c["name"], seed=11, **c["over"])
    g = grid_oracle_for(fac.geometry, 6.0, seg_off, segs)
    g.run(0, c["T"], 8)
    want = tables_of(g)
    return fac, seg_off, segs, p, c["T"], want, g.contact_steps


def _engine(fac, seg_off, segs, p, **kw):
    from citam_b200.engine import DeviceEngine

    eng = DeviceEngine(p["n_agents"], p["n_floors"], 6.0, **kw)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    return eng


def test_device_lut_equals_rectangle_rule(synth):
    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p)
    for f in range(p["n_floors"]):
        minx, miny, lut = fac.lut_expected(f)
        assert eng.floor_meta[f][:2] == (minx, miny)
        assert (eng.floor_lut(f) == lut).all()
    eng.close()


@pytest.mark.parametrize("chunk,no_split", [(0, 0), (37, 0), (5, 0), (249, 0), (495, 0), (37, 1)])
def test_tables_bit_exact_vs_grid_oracle(synth, chunk, no_split, monkeypatch):
    """Both step pipelines: the static/dynamic split (default without a contact log) and the one that
    expands and bins every agent at every step (CITAM_NO_SPLIT=1)."""
    fac, seg_off, segs, p, T, want, want_contacts = synth
    monkeypatch.setenv("CITAM_NO_SPLIT", str(no_split))
    eng = _engine(fac, seg_off, segs, p, steps_per_chunk=chunk)
    eng.run(0, T)
    got = tables_of(eng)
    assert eng.counters()["contact_steps"] == want_contacts
    assert got[2] == want[2]
    assert got[0] == want[0]
    assert got[1] == want[1]
    eng.close()


def test_statistics_from_counter_words_equal_the_table_pass(synth, monkeypatch):
    """citam_statistics' O(n_agents) path (per-agent words: seconds | events << 40) against the pass
    over the pair table, and both against sums taken from the oracle's pair rows."""
    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p)
    eng.run(0, T)
    fast = eng.statistics_sums()
    assert fast == eng.statistics_counters()
    eng.statistics_begin()
    slow = eng.statistics_end()
    assert fast == slow
    nev = {}
    for a1, a2, _first, n_events, _dur in want[0]:
        nev[a1] = nev.get(a1, 0) + n_events
        nev[a2] = nev.get(a2, 0) + n_events
    assert fast["n_pairs"] == len(want[0])
    assert fast["total_duration"] == sum(r[4] for r in want[0])
    assert fast["n_agents_with_contacts"] == len(nev)
    assert fast["sum_events_per_agent"] == sum(nev.values())
    assert fast["max_events_per_agent"] == (max(nev.values()) if nev else 0)
    eng.close()
    monkeypatch.setenv("CITAM_NO_FAST_STATS", "1")
    eng = _engine(fac, seg_off, segs, p)
    eng.run(0, T)
    assert eng.statistics_sums() == fast
    eng.close()


def test_hashed_coordinate_table_equals_dense(synth):
    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p, hashed_coords=True, coord_capacity=1 << 24)
    eng.run(0, T)
    assert tables_of(eng)[1] == want[1]
    eng.close()


def test_time_blocks_in_any_order_and_merge(synth):
    """Time-block sharding: two engines each run half of the blocks; merged tables equal a full run."""
    fac, seg_off, segs, p, T, want, _ = synth
    a = _engine(fac, seg_off, segs, p)
    b = _engine(fac, seg_off, segs, p)
    cuts = list(range(0, T, 500)) + [T]
    for i, (t0, t1) in enumerate(zip(cuts[:-1], cuts[1:])):
        (a if i % 2 == 0 else b).run(t0, t1)
    dur_b = b.agent_durations()
    a.merge(b.pairs(), b.coords())  # pair table + histogram; per-agent seconds are reduced separately
    got = tables_of(a)
    assert got[0] == want[0] and got[1] == want[1]
    assert [int(x) + int(y) for x, y in zip(got[2], dur_b)] == want[2]
    a.close()
    b.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_random_time_ranges_in_random_order(synth, seed):
    """Ranges of random length (1 step to several chunks, nothing aligned) run in shuffled order on two
    engines and merged: the first step of every range owns what it finds, stays and contact runs are
    clipped at every range end, the cursors jump back and forth -- the tables still equal a straight run."""
    fac, seg_off, segs, p, T, want, _ = synth
    rng = np.random.default_rng(seed)
    cuts = [0]
    while cuts[-1] < T:
        cuts.append(min(T, cuts[-1] + int(rng.choice([1, 2, 7, 15, 16, 17, 40, 130, 260, 700]))))
    ranges = list(zip(cuts[:-1], cuts[1:]))
    order = rng.permutation(len(ranges))
    a = _engine(fac, seg_off, segs, p, steps_per_chunk=int(rng.choice([0, 23, 250])))
    b = _engine(fac, seg_off, segs, p)
    for k in order:
        (a if rng.random() < 0.5 else b).run(*ranges[k])
    dur_b = b.agent_durations()
    a.merge(b.pairs(), b.coords())
    got = tables_of(a)
    assert got[0] == want[0] and got[1] == want[1]
    assert [int(x) + int(y) for x, y in zip(got[2], dur_b)] == want[2]
    a.close()
    b.close()


def test_itinerary_store_from_a_device_buffer(synth):
    """citam_load_itineraries takes the segment records from a device pointer too (the multi-GPU store
    distribution all-gathers them over NVLink): same tables as from the host arrays."""
    import torch

    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p)
    so = np.ascontiguousarray(seg_off, np.int64)
    dev = torch.from_numpy(np.ascontiguousarray(segs).view(np.uint8).copy()).cuda()
    eng.load_segments_raw(so.ctypes.data, dev.data_ptr(), len(segs))
    eng.run(0, T)
    assert tables_of(eng)[0] == want[0]
    eng.close()


def test_auto_capacity_grows(synth):
    """Library-chosen pair capacity starts small and is doubled while more than a quarter full."""
    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p, steps_per_chunk=100)
    eng.run(0, T)
    assert tables_of(eng)[0] == want[0]
    eng.close()


def test_simulation_surface_with_device_resident_itineraries(synth, tmp_path):
    """Simulation + SegmentScheduler (no per-step Python lists), summary mode: the end-of-run files
    hold exactly the oracle's tables."""
    import json

    from citam_b200.calculator import CloseContactsCalculator
    from citam_b200.facility import GeometryFacility
    from citam_b200.scheduler import SegmentScheduler
    from citam_b200.simulation import Simulation
    from expected_util import parse_agent_csv, parse_coord_csv, parse_pair_csv

    fac, seg_off, segs, p, T, want, _ = synth
    f = GeometryFacility(fac.geometry, "synthetic")
    sim = Simulation(f, T, p["n_agents"], calculators=[CloseContactsCalculator(f, 6.0, full_fidelity=False)],
                     scheduler=SegmentScheduler(f, T, seg_off, segs))
    sim.block_steps = 1000
    sim.run_serial(str(tmp_path), "s", "r")
    pairs = parse_pair_csv(open(tmp_path / "pair_contact.csv", "rb").read())
    assert pairs == [(a, b, n, d) for a, b, _, n, d in want[0]]
    assert parse_agent_csv(open(tmp_path / "contact_dist_per_agent.csv", "rb").read()) == want[2]
    for fl in range(p["n_floors"]):
        got = parse_coord_csv(open(tmp_path / ("floor_%d" % fl) / "contact_dist_per_coord.csv", "rb").read())
        assert got == {(sx / 2.0, sy / 2.0): c for (f2, sx, sy), c in want[1].items() if f2 == fl}
    st = json.load(open(tmp_path / "statistics.json"))["data"]
    assert st[3]["value"] == len({a for r in want[0] for a in r[:2]})
    assert sim.current_step == T


def test_replica_sweep_matches_oracle_per_replica():
    """BASELINE config 4 in small: independent replicas (occupancy x shifts x distance x seed) run
    concurrently on one GPU; each replica's statistics equal the C oracle's."""
    from citam_b200 import sweep
    from citam_b200 import synthetic as S
    from oracle import port

    specs = [s for s in sweep.make_sweep(10, total_timesteps=900, offices=600)]
    assert len({(s.n_agents, s.n_shifts, s.contact_distance) for s in specs}) == 10
    res = sweep.run_sweep(specs, workers=3)
    assert [r["replica"] for r in res] == list(range(10))
    fac = S.SyntheticFacility(1)
    for spec, r in zip(specs, res):
        seg_off, segs = S.make_itineraries(fac, spec.n_agents, spec.total_timesteps, n_shifts=spec.n_shifts, seed=spec.seed)
        g = grid_oracle_for(fac.geometry, spec.contact_distance, seg_off, segs)
        g.run(0, spec.total_timesteps, 4)
        pairs = g.pairs()
        rows = [(str(a), str(b), int(n), int(d)) for a, b, n, d in zip(pairs["a1"], pairs["a2"], pairs["n_events"], pairs["duration"])]
        want = {d["name"]: d["value"] for d in port.statistics_from_pairs(rows)}
        assert r["statistics"] == want, spec
        assert r["contact_steps"] == g.contact_steps


def test_partition_clear_merge_round_trip_on_device(synth):
    """The multi-GPU exchange on one GPU: partition the pair table by hash(a1, a2) mod G into a
    device buffer, empty the table, merge the parts back from device memory -> the same table; the
    parts are where the host mirror of the hash says, and the device exports equal the host ones."""
    import torch

    from citam_b200 import _lib as L
    from citam_b200 import distributed as D

    fac, seg_off, segs, p, T, want, _ = synth
    eng = _engine(fac, seg_off, segs, p)
    eng.run(0, T)
    n = eng.pair_count()
    assert n == len(want[0])
    for G in (1, 3, 8):
        buf = torch.zeros((max(1, n), 3), dtype=torch.int64, device="cuda")
        counts = eng.partition_pairs_device(buf.data_ptr(), n, G)
        assert int(counts.sum()) == n
        rec = buf.cpu().numpy().view(L.PAIR).reshape(-1)[:n]
        part = D.part_of(rec["a1"], rec["a2"], G)
        assert (part == np.repeat(np.arange(G), counts.astype(np.int64))).all()
        eng.clear_pairs()
        assert eng.pair_count() == 0
        off = 0
        for k in range(G):  # one merge per "sending rank"
            eng.merge_pairs_device(buf.data_ptr() + 24 * off, int(counts[k]))
            off += int(counts[k])
        assert tables_of(eng) == want
    # device-resident exports (sorted on the device) equal the host exports
    out = torch.zeros((max(1, n), 3), dtype=torch.int64, device="cuda")
    assert eng.export_pairs_device(out.data_ptr(), n) == n
    assert out.cpu().numpy().view(L.PAIR).reshape(-1)[:n].tolist() == eng.pairs().tolist()
    co = eng.coords()
    outc = torch.zeros((max(1, len(co)), 3), dtype=torch.int64, device="cuda")
    assert eng.export_coords_device(outc.data_ptr(), len(co)) == len(co)
    assert outc.cpu().numpy().view(L.COORD).reshape(-1)[: len(co)].tolist() == co.tolist()
    # statistics in two halves == in one call
    s1 = eng.statistics_sums()
    eng.statistics_begin()
    assert eng.statistics_end() == s1
    with pytest.raises(L.CitamCapacityError):
        eng.partition_pairs_device(buf.data_ptr(), max(0, n - 1), 2) if n else (_ for _ in ()).throw(L.CitamCapacityError(-3, ""))
    eng.close()
