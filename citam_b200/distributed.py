"""Multi-GPU: time-block sharding of one scenario (or replica sharding) + the final reduction.

The step loop needs no per-step communication: every accumulator is commutative,
so rank r runs time blocks r, r+G, r+2G, ... of the day for ALL agents (each block re-evaluates
the predicate one step before its first step) and the ranks meet once, at the end:

  per-agent contact seconds   all-reduce (sum)
  pair records                all-gather, merged (steps +, events +, first step min)
  coordinate records          all-gather, merged (count +)

`torch.distributed` is the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests); the merge
itself is plain integer work on the host or, through `citam_merge_pairs/coords`, on the device.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import COORD, PAIR


def plan_blocks(total_timesteps: int, block: int, world: int, rank: int) -> List[Tuple[int, int]]:
    """Round-robin time blocks [t0, t1) of this rank; the union over ranks tiles [0, T) exactly."""
    if block <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad block plan arguments")
    out = []
    for k, t0 in enumerate(range(0, total_timesteps, block)):
        if k % world == rank:
            out.append((t0, min(total_timesteps, t0 + block)))
    return out


def plan_replicas(n_replicas: int, world: int, rank: int) -> List[int]:
    """Replica r runs on rank r mod G (independent scenarios, BASELINE config 4)."""
    return [r for r in range(n_replicas) if r % world == rank]


def merge_pair_records(parts: Sequence[np.ndarray]) -> np.ndarray:
    """Sum steps and events, min first step per (a1, a2); rows in first-contact order
    (first_step, a1, a2) == the reference's dict insertion order (contacts.py:135-142)."""
    parts = [np.asarray(p, PAIR) for p in parts if len(p)]
    if not parts:
        return np.zeros(0, PAIR)
    allp = np.concatenate(parts)
    key = (allp["a1"].astype(np.uint64) << np.uint64(32)) | allp["a2"].astype(np.uint64)
    order = np.argsort(key, kind="stable")
    key, allp = key[order], allp[order]
    start = np.flatnonzero(np.concatenate(([True], key[1:] != key[:-1])))
    out = np.zeros(len(start), PAIR)
    out["a1"], out["a2"] = allp["a1"][start], allp["a2"][start]
    out["duration"] = np.add.reduceat(allp["duration"].astype(np.uint64), start)
    out["n_events"] = np.add.reduceat(allp["n_events"].astype(np.uint64), start)
    out["first_step"] = np.minimum.reduceat(allp["first_step"], start)
    return out[np.lexsort((out["a2"], out["a1"], out["first_step"]))]


def merge_coord_records(parts: Sequence[np.ndarray]) -> np.ndarray:
    parts = [np.asarray(p, COORD) for p in parts if len(p)]
    if not parts:
        return np.zeros(0, COORD)
    allc = np.concatenate(parts)
    order = np.lexsort((allc["sy"], allc["sx"], allc["floor"]))
    allc = allc[order]
    new = np.concatenate(([True], (allc["floor"][1:] != allc["floor"][:-1]) | (allc["sx"][1:] != allc["sx"][:-1])
                          | (allc["sy"][1:] != allc["sy"][:-1])))
    start = np.flatnonzero(new)
    out = allc[start].copy()
    out["count"] = np.add.reduceat(allc["count"].astype(np.uint64), start)
    return out


def _dist():
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def all_gather_records(arr: np.ndarray, device=None) -> List[np.ndarray]:
    """All-gather a ragged structured array (as bytes, padded to the longest)."""
    import torch

    dist = _dist()
    world = dist.get_world_size()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    n = torch.tensor([arr.nbytes], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if arr.nbytes:
        buf[: arr.nbytes] = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1)).to(dev)
    outs = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o[:s].cpu().numpy().view(arr.dtype) for o, s in zip(outs, sizes)]


def all_reduce_sum(values: np.ndarray, device=None) -> np.ndarray:
    import torch

    dist = _dist()
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.from_numpy(np.ascontiguousarray(values).astype(np.int64)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def reduce_tables(pairs: np.ndarray, coords: np.ndarray, durations: np.ndarray):
    """The final exchange: every rank ends with the whole run's tables."""
    return (
        merge_pair_records(all_gather_records(pairs)),
        merge_coord_records(all_gather_records(coords)),
        all_reduce_sum(durations),
    )


class ShardedRun:
    """One scenario over the ranks of the current process group, sharded by time blocks."""

    def __init__(self, engine, total_timesteps: int, block: int = 1024):
        dist = _dist()
        self.engine = engine
        self.T = int(total_timesteps)
        self.block = int(block)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def run(self):
        for t0, t1 in plan_blocks(self.T, self.block, self.world, self.rank):
            self.engine.run(t0, t1)
        return self

    def reduce(self, into_engine: bool = False):
        """All ranks get the merged tables; with `into_engine` the other ranks' records are also
        merged into this rank's device tables (`citam_merge_pairs/coords`)."""
        pairs, coords = self.engine.pairs(), self.engine.coords()
        gp, gc = all_gather_records(pairs), all_gather_records(coords)
        if into_engine:
            for r in range(self.world):
                if r != self.rank:
                    self.engine.merge(gp[r], gc[r])
            return self.engine.pairs(), self.engine.coords(), self.engine.agent_durations().astype(np.int64)
        return merge_pair_records(gp), merge_coord_records(gc), all_reduce_sum(self.engine.agent_durations())
