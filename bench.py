#!/usr/bin/env python
"""bench.py — agent-steps/s of the step loop (agent advance + contact detection + accumulation).

Contract (see the task's bench section):
  python bench.py --gpus N --steps K --warmup W            one JSON line from rank 0
  python bench.py --impl reference ...                     the CPU restatement of the reference
                                                           path on the host cores (same JSON)

Workload (default `config3`, the shape BASELINE.json's target is quoted on): synthetic 10-floor
facility, 10^6 agents, 3 shifts, T = 28 800 one-second steps (citam_b200/synthetic.py).

--mode day (default)  One bench "step" = ONE WHOLE SIMULATED DAY: the tables are cleared, all T
  steps run over all agents, and the end-of-run reduction and the statistics are computed -- all
  inside the timed region.  With N ranks the day is cut into N contiguous time ranges, one per rank
  (time sharding; STRONG scaling: the total work is fixed), and the final exchange
  -- NCCL all-reduce of the per-agent counters and the coordinate histogram, all-to-all of the
  hash-partitioned pair table, device merge -- is part of the timed day.  value = K * N_agents * T /
  max-over-ranks time.
--mode blocks  One bench step = one time block of `--block` simulation steps; every rank runs its
  own K blocks spread over the day (weak scaling, no data-path collective) -- the round-1 figure.

  value     inputs resident in HBM (itinerary store loaded before the timed region)
  e2e       the same through the C ABI from HOST buffers: every step uploads the itinerary store from
            pinned host memory (citam_load_itineraries), runs, and reads the per-agent result back
  roofline  dominant kernel (by summed CUDA-event time, measured in the timed region):
            algorithmic bytes = 12 B/agent-step + 64 B/contact-step
  finalize  (day mode, once, after the timed days) compaction + device radix sort of the pair table
            into first-contact order, and one device-to-host export of it (`e2e_with_results`)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PROFILE_CSV = "r2_ncu_full_top_kernels.csv"
METRIC = "agent-steps/sec incl. contact detection"
UNIT = "agent-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="day", choices=["day", "blocks"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config5", "tiny"])
    ap.add_argument("--no-finalize", action="store_true", help="skip the once-only sorted export after the timed days")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity check before timing (N > 1)")
    ap.add_argument("--agents", type=int, default=0, help="override the workload's agent count")
    ap.add_argument("--block", type=int, default=0, help="simulation steps per bench step (0 = auto)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="engine steps_per_chunk (0 = auto)")
    ap.add_argument("--pair-cap-log2", type=int, default=0, help="pair table slots = 2^k (0 = sized for the workload)")
    return ap.parse_args()


_T0 = time.time()
_OUT = sys.stdout  # replaced in __main__ by a private handle on the real stdout


def log(*a):
    print("[bench %7.1fs]" % (time.time() - _T0), *a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/), or None."""
    import csv

    path = os.path.join(ROOT, "profiles", PROFILE_CSV)
    names = {"accumulate": "k_accumulate", "pairs": "k_pairs_changed_split", "expand_bin": "k_expand_dyn", "scatter": "k_scatter_dyn"}
    if not os.path.isfile(path) or kernel not in names:
        return None
    rows = list(csv.reader(open(path)))
    if not rows or names[kernel] not in rows[0]:
        return None
    col = rows[0].index(names[kernel])
    tot = 0.0
    for r in rows[1:]:
        if r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(r[1], 1.0)
            tot += float(r[col]) * scale
    return tot or None


def block_starts(T: int, block: int, count: int, rank: int, world: int, offset: int = 0):
    """`count` block start times for this rank, spread evenly over [0, T - block]."""
    total = count * world
    span = max(1, T - block)
    out = []
    for k in range(count):
        g = k * world + rank
        out.append(int((g + 0.5 + offset) * span / (total + offset)) if total > 1 else 0)
    return out


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=self.f, stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for ln in self.f.read().splitlines():
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def build_workload(args):
    from citam_b200 import synthetic as S

    over = {}
    if args.agents:
        over["n_agents"] = args.agents
    fac, seg_off, segs, p = S.make_config(args.workload, seed=args.seed, **over)
    return fac, seg_off, segs, p


def auto_block(p):
    # ~2.4e8 agent-steps per bench step, at least 32 simulation steps; the dense-crowding shape has
    # ~50 contact-steps per agent-step, so its blocks are kept short
    if p.get("social", 0) > 0.8:
        return 48
    return int(max(32, min(960, 240_000_000 // max(1, p["n_agents"]))))


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle's C grid restatement on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------
def cpu_oracle(fac, seg_off, segs, p, D):
    from citam_b200 import facility as F
    from oracle import grid

    # every core this process may run on (torchrun exports OMP_NUM_THREADS=1; the oracle's time-block
    # threads are requested explicitly, so that setting does not bind it)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    threads = max(grid.max_threads(), cores)
    floors = []
    cache = {}
    for fg in fac.geometry["floors"]:
        tabs = F.floor_tables(fg)
        key = (tabs[4].tobytes(), tabs[5].tobytes(), tabs[6].tobytes())
        if key not in cache:
            cache[key] = grid.build_lut(tabs, threads)
        floors.append((tabs[0], tabs[1], cache[key], len(tabs[4]), tabs[7]))
    return grid.GridOracle(len(seg_off) - 1, D, floors, seg_off, segs), threads


def time_cpu_sample(g, threads, starts, steps_each):
    """Run [s, s+steps_each) for every start; returns (agent_steps, seconds)."""
    t0 = time.perf_counter()
    for s in starts:
        g.run(s, s + steps_each, threads)
    dt = time.perf_counter() - t0
    return len(starts) * steps_each * g.n_agents, dt


def time_faithful_port(fac, seg_off, segs, p, D, n_sub=1500, steps=12):
    """The reference's own formulation (oracle/port.py: scipy pdist over ALL active agents + the
    Python pair loop, 1 core) on a sub-population, since its N x N matrix cannot hold the workload:
    agent-steps/s at n_sub agents, for scale only (cost grows ~ N^2)."""
    from citam_b200 import itinerary as itin
    from oracle import port

    n = min(n_sub, p["n_agents"])
    T = p["total_timesteps"]
    t0s = (T // max(1, p["n_shifts"])) // 2  # the middle of the first shift
    so, sg = itin.slice_segments(seg_off[: n + 1], segs[: int(seg_off[n])], t0s - 1, t0s + steps)
    from citam_b200 import synthetic as S

    its = S.expand_to_reference_format(so, sg, t0s + steps)
    its = [it[t0s - 1:] if len(it) > t0s - 1 else it[-1:] for it in its]  # start the window at t0s-1
    minx, miny, lut = fac.lut_expected(0) if hasattr(fac, "lut_expected") else (0, 0, None)

    def locate(f, x, y):
        fx, fy = x - minx, y - miny
        if lut is None or fx < 0 or fy < 0 or fy >= lut.shape[0] or fx >= lut.shape[1]:
            return None
        v = int(lut[fy, fx])
        return None if v < 0 else v

    sim = port.PortSimulation(its, fac.geometry, D, locate=locate)
    sim.step()  # halo step
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step()
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": UNIT, "cores": 1, "kind": "port (reference formulation, pdist O(N^2))",
            "sample": "the first %d agents x %d steps in the middle of the first shift (%.2f s)" % (n, steps, dt)}


def workload_config(args, p, block):
    """The `config` object: identical for the B200 arm and the reference arm of one mode."""
    T = p["total_timesteps"]
    cfg = {
        "workload": "%s: synthetic %d-floor facility, %d agents, %d shift(s), T=%d 1-s steps, contact distance 6.0"
        % (args.workload, p["n_floors"], p["n_agents"], p["n_shifts"], T),
        "mode": args.mode,
        "n_agents": p["n_agents"],
    }
    if args.mode == "day":
        cfg.update({
            "bench_step": "one whole simulated day: %d steps over all agents from cleared tables, incl. the end-of-run "
                          "reduction and statistics" % T,
            "agent_steps_per_bench_step": p["n_agents"] * T,
            "sharding": "contiguous time ranges, one per rank; final exchange (all-reduce + all-to-all + merge) inside the step",
            "l2": "inputs larger than L2 (each launch group streams > 1 GB of per-step records)"
            if p["n_agents"] * 64 * 16 > 2e8 else "working set may fit L2 (small workload)",
        })
    else:
        cfg.update({
            "bench_step": "one time block of %d simulation steps over all agents; blocks spread evenly over the day" % block,
            "block_steps": block,
            "sharding": "time blocks per rank, no data-path collective",
            "l2": "inputs larger than L2 (each block streams > 1 GB of per-step records)" if p["n_agents"] * block * 12 > 2e8
            else "working set may fit L2 (small workload)",
        })
    return cfg


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU restatement (the upstream reference is pure
    Python + scipy and cannot travel to the GPU box; oracle/grid_oracle.c restates its path and
    is pinned to it by tests/test_oracle_golden.py).  Other ranks exit 0 without work.
    Every bench step is a bounded sample of the arm's workload: one short time block per host
    thread, blocks spread over the day (a whole day would take ~15 minutes of CPU time)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fac, seg_off, segs, p = build_workload(args)
    D = 6.0
    g, threads = cpu_oracle(fac, seg_off, segs, p, D)
    T = p["total_timesteps"]
    block = args.block or auto_block(p)
    # simulation steps per thread and bench step (+ a one-step halo each): 32, or as many (>= 4) as keep the
    # whole --steps K run within ~150 s of CPU time, judged from a short warm-up sample
    probe = max(2, 4 * threads)
    u0, d0 = time_cpu_sample(g, threads, block_starts(T, probe, 1, 0, 1, offset=7), probe)
    rate = u0 / max(d0, 1e-6)
    per = int(max(4, min(32, 150.0 * rate / (max(1, args.steps) * g.n_agents * threads))))
    steps_each = max(2, per * threads)
    starts = block_starts(T, steps_each, args.steps, 0, 1)
    units, dt = time_cpu_sample(g, threads, starts, steps_each)
    v = units / dt
    sample = "%d samples x %d simulation steps x %d agents (%d steps per thread + halo), spread over the day" % (
        len(starts), steps_each, g.n_agents, per)
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.mode == "day" else "weak", "vs_baseline": None,
        "dtype": "i32 positions, f64 distance predicate", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, p, block),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/grid_oracle.c (grid restatement of the reference path, OpenMP time blocks); the upstream "
                "reference itself is single-threaded O(N^2) Python/scipy and cannot hold this N; each bench step is a "
                "bounded sample of the workload, ms_per_step is the sample's time",
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def default_pair_cap(args, p, world):
    if args.pair_cap_log2:
        return 1 << args.pair_cap_log2
    N = p["n_agents"]
    if args.mode == "blocks":
        return int(min(1 << 31, max(1 << 22, 2048 * N)))
    # whole day: ~250 distinct pairs per agent and day on these crowded floors (measured, config 3: 2.5e8 pairs);
    # a rank's table holds its own range's pairs before the exchange and its share of all pairs after it
    # (pairs that meet in several ranks' ranges are held by each of them; measured: the ranks' tables together
    # hold ~1.0 x the day's pairs at 2-8 ranks; 1.3 x the even share is allowed for).  The table is at most
    # a third full at these sizes (every pass over it -- reset, partition, clear -- costs its capacity);
    # a table that overflows aborts the bench.
    est = {"config5": 4096}.get(args.workload, 320) * N
    per_rank = est if world == 1 else int(1.3 * est / world)
    cap = 1 << 22
    while cap < 2 * per_rank:
        cap <<= 1
    return int(min(cap, 1 << 32))


def rows_sorted(buf, n, window=1 << 25):
    """citam_pair rows {a1 | a2 << 32, first_step | n_events << 32, duration} as int64 triples: strictly increasing
    in (first_step, a1, a2)?  Checked on the device, window by window."""
    ok = True
    for lo in range(0, max(0, n - 1), window):
        hi = min(n, lo + window + 1)
        key = buf[lo:hi, 1] & 0xFFFFFFFF
        a1, a2 = buf[lo:hi, 0] & 0xFFFFFFFF, (buf[lo:hi, 0] >> 32) & 0xFFFFFFFF
        lt = (key[:-1] < key[1:]) | ((key[:-1] == key[1:]) & ((a1[:-1] < a1[1:]) | ((a1[:-1] == a1[1:]) & (a2[:-1] < a2[1:]))))
        ok = ok and bool(lt.all().item())
    return ok


def mgpu_parity(rank, world, local):
    """N > 1: the sharded run + device exchange against the C oracle, table for table, before any timing."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import mgpu_check

        mgpu_check.check(None, 4, world, rank, local)
        mgpu_check.check(173, 5, world, rank, local)
        return True
    except AssertionError as e:
        log("multi-GPU parity FAILED on rank %d: %s" % (rank, e))
        return False


def run_b200(args):
    import torch

    from citam_b200 import distributed as DD
    from citam_b200 import synthetic as S
    from citam_b200.engine import DeviceEngine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # first collectives on a cold communicator cost tens of ms: warm it with the message sizes used later
        w = torch.zeros(1 << 24, dtype=torch.int64, device="cuda")
        for _ in range(2):
            dist.all_reduce(w)
            dist.all_to_all_single(torch.empty_like(w), w)
        torch.cuda.synchronize()
        del w

    parity = None
    if world > 1 and not args.no_parity:
        ok = mgpu_parity(rank, world, local)
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity = bool(t.item())
        log("multi-GPU parity vs oracle/grid_oracle.c:", parity)

    fac, seg_off, segs, p = build_workload(args)
    shape = S.activity(seg_off, segs, p["total_timesteps"])
    log("workload built:", len(segs), "segments", shape)
    N, T, D = p["n_agents"], p["total_timesteps"], 6.0
    block = args.block or auto_block(p)
    K, W = args.steps, max(3, args.warmup)

    pair_cap = default_pair_cap(args, p, world)
    coord_cap = int(min(1 << 27, max(1 << 22, 64 * N)))
    eng = DeviceEngine(N, p["n_floors"], D, device=local, pair_capacity=pair_cap, coord_capacity=coord_cap,
                       steps_per_chunk=args.chunk, timing=True)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    log("engine loaded; mode", args.mode, "K", K, "W", W, "pair table 2^%d slots" % (pair_cap.bit_length() - 1))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t_all = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        return float(t_all.item())

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    sharded = DD.ShardedRun(eng, T) if dist is not None else None
    t_lo, t_hi = DD.plan_range(T, world, rank)
    phase = {"run_ms": 0.0, "exchange_ms": 0.0}
    last = {}

    # ---- one bench step ------------------------------------------------------------------------
    if args.mode == "day":
        def step(k, timed=False):
            eng.reset()
            if timed:
                ev[2].record()
            eng.run(t_lo, t_hi)
            if timed:
                ev[3].record()
            last["sums"] = sharded.reduce() if sharded is not None else eng.statistics_sums()
            if timed:
                torch.cuda.synchronize()
                phase["run_ms"] += ev[2].elapsed_time(ev[3])
        units_rank = K * N * (t_hi - t_lo)
        units_all = K * N * T
        steps_w = list(range(W))
        steps_k = list(range(K))
    else:
        starts_w = block_starts(T, block, W, rank, world, offset=3)
        starts = block_starts(T, block, K, rank, world)

        def step(s, timed=False):
            eng.run(s, s + block)
        units_rank = K * block * N
        units_all = units_rank * world
        steps_w, steps_k = starts_w, starts

    # ---- value: itineraries resident in HBM -------------------------------------------------
    for s in steps_w:
        step(s)
    torch.cuda.synchronize()
    c0 = eng.counters()
    tm0 = eng.timings()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev[0].record()
    for s in steps_k:
        step(s, timed=True)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    clk = clocks.stop()
    c1 = eng.counters()
    tm1 = eng.timings()
    ms_max = max_over_ranks(ms)
    value = units_all / (ms_max * 1e-3)
    if args.mode == "day":
        # counters are cleared by every day's reset: c1 is the last day's (this rank's range)
        contacts, pair_tests, runs = (K * c1[k] for k in ("contact_steps", "pair_tests", "contact_runs"))
        launches = c1["kernel_launches"] - c0["kernel_launches"]  # host-side count, not cleared by the resets
    else:
        contacts = c1["contact_steps"] - c0["contact_steps"]
        pair_tests = c1["pair_tests"] - c0["pair_tests"]
        launches = c1["kernel_launches"] - c0["kernel_launches"]
        runs = c1["contact_runs"] - c0["contact_runs"]
    log("timed region done: %.1f ms (max over ranks %.1f), contact-steps on this rank %d, counters %s" % (ms, ms_max, contacts, c1))
    if c1["overflow"]:
        raise SystemExit("bench.py: a device table overflowed (flags %d); the timed numbers are void - raise the capacities"
                         % c1["overflow"])
    day = None
    if args.mode == "day":
        sums = last["sums"]
        day = {
            "day_s": ms_max * 1e-3 / K,
            "run_ms_per_day": phase["run_ms"] / K,
            "exchange_and_statistics_ms_per_day": (ms - phase["run_ms"]) / K,
            "pairs_total": sums["n_pairs"], "total_contact_seconds": sums["total_duration"],
            "pairs_on_rank0_after_exchange": c1["pair_slots_used"],
            "pair_table_load_factor": c1["pair_slots_used"] / pair_cap,
            "reset_included": True,
        }
        if sharded is not None:
            day["exchange"] = dict(sharded.info)
            day["limiting_collective"] = ("all-to-all of pair records (24 B each) by hash(a1,a2) mod %d: %.2f GB sent by rank 0"
                                          % (world, sharded.info.get("a2a_bytes_sent", 0) / 1e9))
    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_src = peaks()
    kern = {}
    for k, v in tm1.items():
        b = tm0.get(k, {"ms": 0.0, "launches": 0})
        if v["launches"] > b["launches"]:
            kern[k] = {"ms": v["ms"] - b["ms"], "launches": v["launches"] - b["launches"]}
    step_kinds = [k for k in kern if k in ("classify", "expand_bin", "scan", "scatter", "pairs", "accumulate")]
    dom = max(step_kinds, key=lambda k: kern[k]["ms"]) if step_kinds else None
    # k_accumulate is the HBM-bound kernel of the step (random read-modify-writes); the detection kernel
    # that ties with it is bound by instruction issue.  The roofline line describes k_accumulate whenever
    # it is within 15 % of the longest kernel.
    if dom and "accumulate" in kern and kern["accumulate"]["ms"] >= 0.85 * kern[dom]["ms"]:
        dom = "accumulate"
    roof = None
    if dom:
        run_ms = phase["run_ms"] if args.mode == "day" else ms
        # Algorithmic bytes (12 B per agent-step + 64 B per contact-step), applied to the units
        # a launch really processes: with run aggregation k_accumulate applies ONE table update per contact
        # RUN (a pair standing together for an hour is one record), so it is credited 64 B per applied
        # update, not per contact-step covered; the survey's own rule is printed beside it.
        own = 64.0 * runs if dom == "accumulate" else 12.0 * units_rank
        own_survey = 64.0 * contacts if dom == "accumulate" else 12.0 * units_rank
        per_launch = own / kern[dom]["launches"]
        avg_ms = kern[dom]["ms"] / kern[dom]["launches"]
        achieved = per_launch / (avg_ms * 1e-3) / 1e9
        alg_updates = 12.0 * units_rank + 64.0 * runs
        alg_survey = 12.0 * units_rank + 64.0 * contacts
        traffic = profiled_traffic(dom) if args.workload == "config3" and not args.agents else None
        roof = {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic,
            "traffic_source": "profiles/%s (one launch of the same build, a 480-step chunk of 1e6 agents)" % PROFILE_CSV,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
            "algorithmic_bytes_rule": "k_accumulate: 64 B x table updates applied (one per contact run); position kernels: "
                                      "12 B x agent-steps; whole step: both",
            "survey_rule": {"what": "64 B x contact-steps covered (SURVEY.md 8d as written: run aggregation makes this exceed the "
                                    "peak, the updates were never made)",
                            "achieved": own_survey / kern[dom]["launches"] / (avg_ms * 1e-3) / 1e9,
                            "frac": own_survey / kern[dom]["launches"] / (avg_ms * 1e-3) / 1e9 / peak},
            "dram_throughput_of_launch": (traffic / (avg_ms * 1e-3) / 1e9) if traffic else None,
            "kernel_share_of_step": kern[dom]["ms"] / run_ms,
            "kernels_ms": {k: round(v["ms"], 3) for k, v in kern.items()},
            "kernels_note": "CUDA-event times per kernel kind; k_accumulate runs on a side stream next to the following "
                            "chunk's build, so the kinds need not add up to the step",
            "whole_step_achieved": alg_updates / (run_ms * 1e-3) / 1e9, "whole_step_frac": alg_updates / (run_ms * 1e-3) / 1e9 / peak,
            "whole_step_survey_rule_frac": alg_survey / (run_ms * 1e-3) / 1e9 / peak,
        }

    if roof and roof["kernel"] == "accumulate" and world == 1:
        # what the memory system sustains for this access pattern: random 16-byte probe + 8-byte reduction
        # over a table of the pair table's size, nothing else (citam_measure_random_updates)
        try:
            from citam_b200.engine import measure_random_updates

            mups = measure_random_updates(pair_cap * 16, 1 << 28, True, device=local)
            per_record = 1.0  # the pair-table update (17 GB, no locality); the histogram one is counted as free
            rec_rate = runs / (kern["accumulate"]["ms"] * 1e-3) / 1e6
            roof["random_update_ceiling"] = {
                "million_updates_per_s": mups, "table_bytes": int(pair_cap * 16),
                "k_accumulate_million_records_per_s": rec_rate,
                "k_accumulate_million_updates_per_s": rec_rate * per_record,
                "frac_of_ceiling": rec_rate * per_record / mups,
                "what": "2^28 random (16-byte load + 8-byte reduction) updates, four in flight per thread, over a zeroed table "
                        "of the pair table's size; k_accumulate makes one such update per record (pair slot), one more over the "
                        "0.46 GB histogram and two L2-resident ones: the fraction counts the pair-table update alone",
            }
        except Exception as e:  # measurement aid only
            roof["random_update_ceiling"] = {"error": str(e)[:200]}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    e2e = None
    e2e_res = None
    if not args.no_e2e:
        so_p = torch.from_numpy(np.ascontiguousarray(seg_off, np.int64)).pin_memory()
        sg_p = torch.from_numpy(np.ascontiguousarray(segs).view(np.uint8)).pin_memory()
        out_dur = torch.empty(N, dtype=torch.int64).pin_memory()
        h2d_step = so_p.numel() * 8 + sg_p.numel()
        h2d_rank = [h2d_step]

        def e2e_step(s):
            if sharded is not None:
                # every rank holds the store in pinned host memory; each uploads 1/G of it, NVLink all-gathers it
                h2d_rank[0] = sharded.load_store(so_p, sg_p, len(segs))
            else:
                eng.load_segments_raw(so_p.data_ptr(), sg_p.data_ptr(), len(segs))
            step(s)
            eng.agent_durations_into(out_dur.data_ptr())

        if args.mode == "blocks":
            eng.reset()
        for s in steps_w[:2]:
            e2e_step(s)
        barrier()
        ev[0].record()
        for s in steps_k:
            e2e_step(s)
        ev[1].record()
        barrier()
        ms_e = max_over_ranks(ev[0].elapsed_time(ev[1]))
        e2e = {
            "value": units_all / (ms_e * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(h2d_step if sharded is None else h2d_rank[0] * world), "d2h_bytes_per_step": N * 8 * world,
            "h2d_bytes_per_step_and_rank": int(h2d_rank[0]),
            "api": ("citam_load_itineraries(whole store, pinned host) + " if sharded is None else
                    "ShardedRun.load_store (1/G of the store from pinned host memory per rank + NCCL all-gather + "
                    "citam_load_itineraries from the device buffer) + ") + (
                "citam_reset + citam_run(this rank's range) + final exchange + citam_statistics" if args.mode == "day"
                else "citam_run(block)") + " + citam_export_agent_durations",
        }
        log("e2e done: %.1f ms" % ms_e)

    # ---- finalize: the pair table in first-contact order (once; day mode) ---------------------------
    fin = None
    if args.mode == "day" and not args.no_finalize:
        try:
            n_pairs = eng.pair_count()
            buf = torch.empty((max(1, n_pairs), 3), dtype=torch.int64, device="cuda")
            barrier()
            ev[0].record()
            got = eng.export_pairs_device(buf.data_ptr(), n_pairs)  # first call: allocates the sort's ping-pong buffers
            ev[1].record()
            barrier()
            first_ms = ev[0].elapsed_time(ev[1])
            ev[0].record()
            got = eng.export_pairs_device(buf.data_ptr(), n_pairs)
            ev[1].record()
            torch.cuda.synchronize()
            fin = {"finalize_ms": max_over_ranks(ev[0].elapsed_time(ev[1])), "finalize_first_call_ms": max_over_ranks(first_ms),
                   "pairs_sorted_on_this_rank": int(got),
                   "what": "compaction of the pair table + device radix sort by (first contact step, a1, a2) into a device buffer; "
                           "the first call also allocates the sort's buffers (cudaMalloc of 2 x 16 B x pairs)"}
            fin["sorted_ok"] = rows_sorted(buf, int(got))
            if world == 1 and not args.no_e2e and got * 24 < 48e9:
                host = torch.empty((max(1, got), 3), dtype=torch.int64).pin_memory()
                eng.reset()
                barrier()
                ev[0].record()
                e2e_step(0)
                got2 = eng.export_pairs_device(buf.data_ptr(), n_pairs)
                host[:got2].copy_(buf[:got2], non_blocking=True)
                ev[1].record()
                torch.cuda.synchronize()
                ms_r = ev[0].elapsed_time(ev[1])
                e2e_res = {"value": N * T / (ms_r * 1e-3), "unit": UNIT, "d2h_bytes": int(got2) * 24 + N * 8,
                           "what": "one e2e day + the sorted pair table copied to pinned host memory"}
                del host
            del buf
        except Exception as e:  # out of memory on the biggest tables: the timed numbers above stand
            fin = {"finalize_ms": None, "error": str(e)[:200]}
        log("finalize:", fin)

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        g, threads = cpu_oracle(fac, seg_off, segs, p, D)
        steps_each = (8 if N <= 100_000 else 2) * threads
        cs = block_starts(T, steps_each, 4 if N <= 100_000 else 2, 0, 1)
        time_cpu_sample(g, threads, cs[:1], steps_each)
        units, dt = time_cpu_sample(g, threads, cs, steps_each)
        cpu = {"value": units / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle/grid_oracle.c, %d blocks x %d simulation steps x %d agents spread over the day (%.1f s)"
               % (len(cs), steps_each, N, dt)}
        g.close()
        cpu["faithful_port"] = time_faithful_port(fac, seg_off, segs, p, D)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong" if args.mode == "day" else "weak",
            "vs_baseline": None, "dtype": "i32 positions, f64 distance predicate", "data": "synthetic",
            "config": workload_config(args, p, block),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "gpu_launches": int(launches),
            "active_fraction": shape["active_fraction"], "moving_fraction": shape["moving_fraction"],
            "active_agent_steps_per_s": value * shape["active_fraction"],
            "contacts_per_agent_step": contacts / max(1, units_rank),
            "pair_tests_per_s": pair_tests / ((phase["run_ms"] if args.mode == "day" else ms) * 1e-3),
            "table_updates_per_contact_step": runs / max(1, contacts),
            "binned_fraction_of_agent_steps": (K if args.mode == "day" else 1) * c1.get("binned_agent_steps", 0) / max(1, units_rank)
            if args.mode == "day" else (c1.get("binned_agent_steps", 0) - c0.get("binned_agent_steps", 0)) / max(1, units_rank),
            "day": day, "finalize": fin, "e2e_with_results": e2e_res, "mgpu_parity": parity,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr,
    # so that library banners (e.g. NCCL's version line) cannot land in front of the result
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
