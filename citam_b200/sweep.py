"""Scenario-replica sweeps (BASELINE config 4): many independent small scenarios, sharded over GPUs.

Replicas do not interact, so replica r simply runs on rank r mod G (partitioning 1) and
only the per-replica statistics travel at the end.  One scenario of a few thousand agents does not fill
a B200, so each rank runs several replicas concurrently: one engine and one CUDA stream per worker
thread (the C ABI releases the GIL; calls on different engines are independent).
"""
from __future__ import annotations

import itertools
from concurrent.futures import ThreadPoolExecutor
from dataclasses import asdict, dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import synthetic as S
from .distributed import plan_replicas
from .engine import DeviceEngine, statistics_from_sums


@dataclass(frozen=True)
class ReplicaSpec:
    replica: int
    n_agents: int
    n_shifts: int
    contact_distance: float
    seed: int
    total_timesteps: int = 28_800
    n_floors: int = 1


def make_sweep(n_replicas: int = 1024, total_timesteps: int = 28_800, offices: int = 2_500) -> List[ReplicaSpec]:
    """occupancy {0.25, 0.5, 0.75, 1.0} x shifts {1, 2, 3, 4} x contact distance {3, 4.5, 6, 9} x seeds;
    agents = round(occupancy x offices) (the reference's rule, citam/engine/main.py:356-364)."""
    grid = list(itertools.product((0.25, 0.5, 0.75, 1.0), (1, 2, 3, 4), (3.0, 4.5, 6.0, 9.0)))
    out = []
    for r in range(n_replicas):
        occ, sh, d = grid[r % len(grid)]
        out.append(ReplicaSpec(r, int(round(occ * offices)), sh, d, 4000 + r // len(grid), total_timesteps))
    return out


def run_replica(spec: ReplicaSpec, device: int = 0, stream: Optional[int] = None, facility=None) -> Dict:
    fac = facility if facility is not None else S.SyntheticFacility(spec.n_floors)
    seg_off, segs = S.make_itineraries(fac, spec.n_agents, spec.total_timesteps, n_shifts=spec.n_shifts, seed=spec.seed)
    eng = DeviceEngine(spec.n_agents, spec.n_floors, spec.contact_distance, device=device)
    try:
        if stream is not None:
            eng.set_stream(stream)
        eng.load_geometry(fac.geometry)
        eng.load_segments(seg_off, segs)
        eng.run(0, spec.total_timesteps)
        sums = eng.statistics_sums()
        c = eng.counters()
    finally:
        eng.close()
    res = asdict(spec)
    res.update(sums)
    res["statistics"] = {d["name"]: d["value"] for d in statistics_from_sums(sums)}
    res["contact_steps"] = c["contact_steps"]
    res["agent_steps"] = c["agent_steps"]
    return res


def run_sweep(specs: Sequence[ReplicaSpec], device: int = 0, workers: int = 4, world: int = 1, rank: int = 0) -> List[Dict]:
    """This rank's share of the sweep, `workers` replicas in flight at a time."""
    import torch

    mine = [specs[i] for i in plan_replicas(len(specs), world, rank)]
    facs: Dict[int, S.SyntheticFacility] = {}
    for s in mine:
        facs.setdefault(s.n_floors, S.SyntheticFacility(s.n_floors))
    streams = [torch.cuda.Stream(device=device) for _ in range(max(1, workers))]

    def work(args):
        k, spec = args
        with torch.cuda.device(device):
            return run_replica(spec, device, streams[k % len(streams)].cuda_stream, facs[spec.n_floors])

    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        return list(pool.map(work, enumerate(mine)))


def gather_results(local: List[Dict]) -> List[Dict]:
    """All ranks' replica results on every rank, in replica order (torch.distributed object gather)."""
    import torch.distributed as dist

    parts: List = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return sorted((r for p in parts for r in p), key=lambda r: r["replica"])
