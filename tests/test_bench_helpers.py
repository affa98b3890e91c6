"""CPU: bench.py's host-side helpers and the replica-sweep specification."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from citam_b200 import sweep  # noqa: E402


def test_block_starts_spread_over_the_day_and_disjoint_across_ranks():
    T, block, K = 28_800, 240, 12
    for world in (1, 2, 8):
        starts = sorted(s for r in range(world) for s in bench.block_starts(T, block, K, r, world))
        assert len(starts) == K * world == len(set(starts))
        assert starts[0] >= 0 and starts[-1] + block <= T
        gaps = [b - a for a, b in zip(starts[:-1], starts[1:])]
        assert min(gaps) >= block or world * K * block > T  # blocks of one run do not overlap
        assert starts[-1] > T * 0.8 and starts[0] < T * 0.2  # the whole day is sampled
    # warm-up blocks use other offsets than the timed ones
    assert set(bench.block_starts(T, block, 3, 0, 1, offset=3)).isdisjoint(bench.block_starts(T, block, K, 0, 1))


def test_auto_block_and_profile_lookup():
    assert bench.auto_block(dict(n_agents=1_000_000)) == 240
    assert bench.auto_block(dict(n_agents=10_000)) == 960
    assert bench.auto_block(dict(n_agents=200_000, social=0.85)) == 48
    t = bench.profiled_traffic("accumulate")
    assert t is None or t > 1e9
    assert bench.profiled_traffic("no_such_kernel") is None
    peak, src = bench.peaks()
    assert 3000 < peak < 9000 and src


def test_sweep_specification_covers_the_grid():
    specs = sweep.make_sweep(1024)
    assert len(specs) == 1024 and [s.replica for s in specs] == list(range(1024))
    combos = {(s.n_agents, s.n_shifts, s.contact_distance) for s in specs}
    assert len(combos) == 64  # 4 occupancies x 4 shift counts x 4 distances
    assert {s.n_agents for s in specs} == {625, 1250, 1875, 2500}
    assert len({s.seed for s in specs}) == 16
    from citam_b200.distributed import plan_replicas

    mine = [plan_replicas(1024, 8, r) for r in range(8)]
    assert sorted(i for m in mine for i in m) == list(range(1024)) and all(len(m) == 128 for m in mine)


def test_default_pair_table_is_sized_per_rank():
    """config 3 makes ~2.5e8 pairs a day; a rank's table (its own range's pairs, then its hash share) stays
    a quarter full at 1-8 ranks and shrinks with the rank count (every pass over it costs its capacity)."""
    import types

    caps = {}
    for world in (1, 2, 4, 8):
        a = types.SimpleNamespace(pair_cap_log2=0, mode="day", workload="config3")
        caps[world] = bench.default_pair_cap(a, {"n_agents": 1_000_000}, world)
        assert caps[world] & (caps[world] - 1) == 0
        assert 0.1 < 2.5e8 / world / caps[world] < 0.5
    assert caps[1] == 2 * caps[2] == 4 * caps[4] == 8 * caps[8]
    a = types.SimpleNamespace(pair_cap_log2=27, mode="day", workload="config3")
    assert bench.default_pair_cap(a, {"n_agents": 1_000_000}, 4) == 1 << 27


def test_committed_bench_line_follows_the_contract():
    """The bench line kept under profiles/ (written by bench.py on a B200) carries every key the
    measurement contract names, and bench.py can read the committed ncu summary it cites."""
    import json

    with open(os.path.join(ROOT, "profiles", "r2", "bench_config3.json")) as f:
        d = json.load(f)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["data"] == "synthetic"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    # the committed ncu summary: DRAM bytes of one k_accumulate launch (the bench line was printed from the
    # capture taken one build earlier; the kernel is the same, the bytes agree to a percent)
    t = bench.profiled_traffic("accumulate")
    assert t is not None and abs(t - r["traffic"]) < 0.02 * t
