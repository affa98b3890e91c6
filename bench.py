#!/usr/bin/env python
"""bench.py — agent-steps/s of the step loop (agent advance + contact detection + accumulation).

Contract (see the task's bench section):
  python bench.py --gpus N --steps K --warmup W            one JSON line from rank 0
  python bench.py --impl reference ...                     the CPU restatement of the reference
                                                           path on the host cores (same JSON)

Workload (default `config3`, the shape BASELINE.json's target is quoted on): synthetic 10-floor
facility, 10^6 agents, 3 shifts, T = 28 800 one-second steps (citam_b200/synthetic.py).  One
bench "step" = one *time block* of `--block` simulation steps over ALL agents; the K timed
blocks are spread evenly over the day so that the mix of empty / busy hours is the whole-day
mix.  With N ranks every rank runs its own K blocks (time-block sharding, SURVEY.md §8e: no
data-path collective; weak scaling) and the value is all ranks' agent-steps / max-over-ranks time.

  value     inputs resident in HBM (itinerary store loaded before the timed region)
  e2e       each block's itinerary slice copied from pinned host memory through the C ABI,
            the block run, and its per-agent result read back, all inside the timed region
  roofline  dominant kernel (by summed CUDA-event time, measured in the timed region):
            algorithmic bytes = 12 B/agent-step + 64 B/contact-step
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

RESET_EVERY = 40  # bench blocks between table resets (default runs never reach it)
METRIC = "agent-steps/sec incl. contact detection"
UNIT = "agent-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config3", choices=["config2", "config3", "config5", "tiny"])
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

    path = os.path.join(ROOT, "profiles", "r1_ncu_full_top_kernels.csv")
    names = {"accumulate": "k_accumulate", "pairs": "k_pairs_changed", "expand_bin": "k_expand_bin"}
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


def run_reference(args):
    """`--impl reference`: rank 0 times the CPU restatement (the upstream reference is pure
    Python + scipy and cannot travel to the GPU box; oracle/grid_oracle.c restates its path and
    is pinned to it by tests/test_oracle_golden.py).  Other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fac, seg_off, segs, p = build_workload(args)
    D = 6.0
    g, threads = cpu_oracle(fac, seg_off, segs, p, D)
    T = p["total_timesteps"]
    # bounded sample per bench step: one short time block per thread
    per = 4  # simulation steps per thread and bench step (plus the one-step halo each thread evaluates)
    steps_each = max(2, per * threads)
    starts_w = block_starts(T, steps_each, max(1, args.warmup), 0, 1, offset=7)
    starts = block_starts(T, steps_each, args.steps, 0, 1)
    time_cpu_sample(g, threads, starts_w[: max(1, min(args.warmup, 1))], steps_each)
    units, dt = time_cpu_sample(g, threads, starts, steps_each)
    v = units / dt
    sample = "%d blocks x %d simulation steps x %d agents, blocks spread over the day" % (len(starts), steps_each, g.n_agents)
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32 positions, f64 distance predicate", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, p, steps_each),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/grid_oracle.c (grid restatement of the reference path, OpenMP time blocks); the upstream "
                "reference itself is single-threaded O(N^2) Python/scipy and cannot hold this N",
    }
    print(json.dumps(line), file=_OUT, flush=True)


def workload_config(args, p, block):
    return {
        "workload": "%s: synthetic %d-floor facility, %d agents, %d shift(s), T=%d 1-s steps, contact distance 6.0"
        % (args.workload, p["n_floors"], p["n_agents"], p["n_shifts"], p["total_timesteps"]),
        "bench_step": "one time block of %d simulation steps over all agents; blocks spread evenly over the day" % block,
        "block_steps": block,
        "n_agents": p["n_agents"],
        "sharding": "time blocks per rank, no data-path collective",
        "l2": "inputs larger than L2 (each block streams > 1 GB of per-step records)" if p["n_agents"] * block * 12 > 2e8
        else "working set may fit L2 (small workload)",
    }


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    from citam_b200 import _lib as L
    from citam_b200 import itinerary as itin
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

    fac, seg_off, segs, p = build_workload(args)
    log("workload built:", len(segs), "segments")
    N, T, D = p["n_agents"], p["total_timesteps"], 6.0
    block = args.block or auto_block(p)
    K, W = args.steps, max(3, args.warmup)

    pair_cap = int(min(1 << 31, max(1 << 22, 2048 * N)))
    if args.pair_cap_log2:
        pair_cap = 1 << args.pair_cap_log2
    coord_cap = int(min(1 << 27, max(1 << 22, 64 * N)))
    eng = DeviceEngine(N, p["n_floors"], D, device=local, pair_capacity=pair_cap, coord_capacity=coord_cap,
                       steps_per_chunk=args.chunk, timing=True)
    eng.load_geometry(fac.geometry)
    eng.load_segments(seg_off, segs)
    log("engine loaded; block", block, "K", K, "W", W)

    starts_w = block_starts(T, block, W, rank, world, offset=3)
    starts = block_starts(T, block, K, rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: itineraries resident in HBM -------------------------------------------------
    for s in starts_w:
        eng.run(s, s + block)
    torch.cuda.synchronize()
    c0 = eng.counters()
    tm0 = eng.timings()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k, s in enumerate(starts):
        if k and k % RESET_EVERY == 0:
            eng.reset()  # keeps the pair table's load bounded on very long runs (inside the timed region)
        eng.run(s, s + block)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    c1 = eng.counters()
    tm1 = eng.timings()
    t_all = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms_max = float(t_all.item())
    units_rank = K * block * N
    value = units_rank * world / (ms_max * 1e-3)
    if K > RESET_EVERY:  # counters were cleared on the way: count from the last reset, scaled to K blocks
        tail = K - (K - 1) // RESET_EVERY * RESET_EVERY
        c0 = {k: (v if k == "kernel_launches" else 0) for k, v in c0.items()}
        c1 = {k: (v * K // tail if k in ("contact_steps", "pair_tests", "contact_runs") else v) for k, v in c1.items()}
    contacts = c1["contact_steps"] - c0["contact_steps"]
    pair_tests = c1["pair_tests"] - c0["pair_tests"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]
    runs = c1["contact_runs"] - c0["contact_runs"]

    log("timed region done: %.1f ms, contacts %d, counters %s" % (ms, contacts, c1))
    if c1["overflow"]:
        raise SystemExit("bench.py: a device table overflowed (flags %d); the timed numbers are void - raise the capacities"
                         % c1["overflow"])
    # ---- roofline of the dominant kernel -----------------------------------------------------
    peak, peak_src = peaks()
    kern = {}
    for k, v in tm1.items():
        b = tm0.get(k, {"ms": 0.0, "launches": 0})
        if v["launches"] > b["launches"]:
            kern[k] = {"ms": v["ms"] - b["ms"], "launches": v["launches"] - b["launches"]}
    dom = max(kern, key=lambda k: kern[k]["ms"]) if kern else None
    roof = None
    if dom:
        alg_bytes = 12.0 * units_rank + 64.0 * contacts  # whole step
        # the dominant kernel's own share of it: k_accumulate owns the 64 B per contact-step,
        # the position kernels the 12 B per agent-step
        own = 64.0 * contacts if dom == "accumulate" else 12.0 * units_rank
        per_launch = own / kern[dom]["launches"]
        avg_ms = kern[dom]["ms"] / kern[dom]["launches"]
        achieved = per_launch / (avg_ms * 1e-3) / 1e9
        roof = {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": profiled_traffic(dom) if args.workload == "config3" and not args.agents else None, "traffic_source": "profiles/r1_ncu_full_top_kernels.csv (one launch, 64 steps x 1e6 agents)",
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": per_launch, "avg_launch_ms": avg_ms,
            "algorithmic_bytes_rule": "k_accumulate: 64 B x contact-steps; expand/scatter/pairs: 12 B x agent-steps; whole step: both",
            "kernel_share_of_step": kern[dom]["ms"] / ms,
            "kernels_ms": {k: round(v["ms"], 3) for k, v in kern.items()},
            "whole_step_achieved": alg_bytes / (ms * 1e-3) / 1e9, "whole_step_frac": alg_bytes / (ms * 1e-3) / 1e9 / peak,
        }

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -------------------
    e2e = None
    if not args.no_e2e:
        slices = []
        for s in starts_w[:2] + starts:
            so, sg = itin.slice_segments(seg_off, segs, s - 1, s + block)
            so_p = torch.from_numpy(so).pin_memory()
            sg_p = torch.from_numpy(sg.view(np.uint8)).pin_memory()
            slices.append((s, so_p, sg_p, len(sg)))
        out_dur = torch.empty(N, dtype=torch.int64).pin_memory()
        eng.reset()

        def e2e_block(s, so_p, sg_p, nseg):
            eng.load_segments_raw(so_p.data_ptr(), sg_p.data_ptr(), nseg)
            eng.run(s, s + block)
            eng.agent_durations_into(out_dur.data_ptr())

        for sl in slices[:2]:
            e2e_block(*sl)
        barrier()
        ev0.record()
        h2d = 0
        for sl in slices[2:]:
            e2e_block(*sl)
            h2d += sl[1].numel() * 8 + sl[2].numel()
        ev1.record()
        barrier()
        ms_e = ev0.elapsed_time(ev1)
        t_all = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        e2e = {
            "value": units_rank * world / (float(t_all.item()) * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": N * 8,
            "api": "citam_load_itineraries(block slice, pinned host) + citam_run(block) + citam_export_agent_durations",
        }
        eng.load_segments(seg_off, segs)

    log("e2e done")
    # ---- end-of-run reduction (multi-GPU only; outside the per-step timing) ---------------------------
    final_reduce_ms = None
    if dist is not None:
        ptr, n_el = eng.agent_durations_device()

        class _Holder:
            __cuda_array_interface__ = {"shape": (n_el,), "typestr": "<i8", "data": (ptr, False), "version": 3}

        tdur = torch.as_tensor(_Holder(), device="cuda")
        sums = eng.statistics_sums()
        tsum = torch.tensor([sums["total_duration"], sums["n_pairs"]], dtype=torch.int64, device="cuda")
        barrier()
        ev0.record()
        dist.all_reduce(tdur)
        dist.all_reduce(tsum)
        ev1.record()
        torch.cuda.synchronize()
        final_reduce_ms = ev0.elapsed_time(ev1)

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
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32 positions, f64 distance predicate", "data": "synthetic",
            "config": workload_config(args, p, block),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clk, "gpu_launches": int(launches),
            "contact_steps_per_step": contacts / K, "contacts_per_agent_step": contacts / units_rank,
            "pair_tests_per_s": pair_tests / (ms * 1e-3), "table_updates_per_contact_step": runs / max(1, contacts),
            "final_reduce_ms": final_reduce_ms,
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
