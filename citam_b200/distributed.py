"""Multi-GPU: one scenario sharded by time over the ranks (or replica sharding) + the final exchange.

The step loop needs no per-step communication: every accumulator is commutative,
so rank r runs its own part of the day for ALL agents -- a contiguous range of steps (the default:
contact runs then aggregate over the whole range) or round-robin time blocks -- each part
re-evaluating the predicate one step before its first step, and the ranks meet once, at the end,
on DEVICE buffers (NCCL over NVLink through torch.distributed):

  per-agent contact seconds   all-reduce (sum) of the engine's counter array
  coordinate histogram        all-reduce (sum) of the dense histogram
  pair records                all-to-all by hash(a1, a2) mod G of the partitioned table
                              (citam_partition_pairs_device), then a device merge
                              (citam_merge_pairs_device: steps +, events +, first step min):
                              every rank ends with the complete records of its share of the pairs
  statistics                  per-share integer sums; the per-agent event counts are all-reduced

This is the parallel runner over "timestep blocks" that the reference's `run_serial` docstring
asks for (citam/engine/core/simulation.py:222-227).  The host-side merges below do the same on
numpy records (gloo CPU tests, small runs that want the whole table on every rank).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._lib import COORD, PAIR, CitamCapacityError


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


def plan_range(total_timesteps: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous share [t0, t1) of rank `rank`; the shares tile [0, T) exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad range plan arguments")
    return (total_timesteps * rank) // world, (total_timesteps * (rank + 1)) // world


_M1, _M2 = np.uint64(0xFF51AFD7ED558CCD), np.uint64(0xC4CEB9FE1A85EC53)


def part_of(a1: np.ndarray, a2: np.ndarray, n_parts: int) -> np.ndarray:
    """Destination rank of a pair: the device's `part_of` (mix64(a1 << 32 | a2) >> 40) mod n_parts."""
    k = (np.asarray(a1, np.uint64) << np.uint64(32)) | np.asarray(a2, np.uint64)
    with np.errstate(over="ignore"):
        k = k ^ (k >> np.uint64(33))
        k = k * _M1
        k = k ^ (k >> np.uint64(33))
        k = k * _M2
        k = k ^ (k >> np.uint64(33))
    return ((k >> np.uint64(40)) % np.uint64(n_parts)).astype(np.int64)


def exchange_pairs_host(pairs: np.ndarray) -> np.ndarray:
    """Host mirror of the device exchange (gloo has no all-to-all): every rank ends with the merged
    records of its share of the pairs."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = all_gather_records(np.ascontiguousarray(pairs, PAIR))
    mine = [p[part_of(p["a1"], p["a2"], world) == rank] for p in parts]
    return merge_pair_records(mine)


class _DevView:
    """A torch tensor over engine-owned device memory (no copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_tensor(ptr: int, n: int, typestr: str = "<i8"):
    import torch

    return torch.as_tensor(_DevView(ptr, n, typestr), device="cuda")


class ShardedRun:
    """One scenario over the ranks of the current process group, sharded by time.

    `block=None`: rank r runs the contiguous range `plan_range(T, G, r)` in one `citam_run` (contact
    runs aggregate over the whole range); `block=k`: round-robin blocks of k steps (`plan_blocks`)."""

    def __init__(self, engine, total_timesteps: int, block: Optional[int] = None):
        dist = _dist()
        self.engine = engine
        self.T = int(total_timesteps)
        self.block = None if block is None else int(block)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.info: Dict[str, float] = {}

    def load_store(self, seg_off_host, segs_host_bytes, n_segs: int):
        """Load ONE itinerary store, held in (pinned) host memory on every rank, into all ranks' engines
        with 1/G of the host-to-device traffic per rank: rank r uploads the r-th of G equal byte ranges of
        the segment array, an NCCL all-gather over NVLink completes it on every GPU, and the engine loads it
        from the device buffer.  `seg_off_host`: torch int64 [n_agents + 1]; `segs_host_bytes`: torch uint8
        view of the citam_segment records.  Returns the bytes this rank copied from the host."""
        import torch

        dist = _dist()
        G, r = self.world, self.rank
        total = int(segs_host_bytes.numel())
        chunk = (total + G - 1) // G
        chunk = (chunk + 255) // 256 * 256
        if getattr(self, "_store", None) is None or self._store.numel() != G * chunk:
            self._store = torch.empty(G * chunk, dtype=torch.uint8, device="cuda")
            self._part = torch.empty(chunk, dtype=torch.uint8, device="cuda")
        lo, hi = min(total, r * chunk), min(total, (r + 1) * chunk)
        if hi > lo:
            self._part[: hi - lo].copy_(segs_host_bytes[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(self._store, self._part)
        torch.cuda.current_stream().synchronize()
        self.engine.load_segments_raw(seg_off_host.data_ptr(), self._store.data_ptr(), n_segs)
        return (hi - lo) + int(seg_off_host.numel()) * 8

    def plan(self) -> List[Tuple[int, int]]:
        if self.block is None:
            return [plan_range(self.T, self.world, self.rank)]
        return plan_blocks(self.T, self.block, self.world, self.rank)

    def run(self):
        for t0, t1 in self.plan():
            if t1 > t0:
                self.engine.run(t0, t1)
        return self

    # -- the final exchange on device buffers (NCCL) --------------------------------------------
    def reduce(self):
        """After this every rank's engine holds: the whole run's per-agent contact seconds and
        coordinate histogram, and the complete pair records of its share of the pairs
        (hash(a1, a2) mod G == rank).  Returns the whole run's statistics sums."""
        import torch

        dist = _dist()
        eng, G = self.engine, self.world
        if (eng.n_agents - 1) * self.T >= 1 << 40:
            raise ValueError("(n_agents - 1) x total_timesteps must stay below 2^40 (per-agent counter width)")
        import time

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record()
        # the two all-reduces run on NCCL's stream while the engine's stream partitions the pair table
        ptr, n = eng.agent_durations_device()
        dur_t = device_tensor(ptr, n)
        pending = [dist.all_reduce(dur_t, async_op=True)]
        hptr, hn = eng.hist_device()
        hist_t = device_tensor(hptr, hn) if hn else None
        if hn:
            pending.append(dist.all_reduce(hist_t, async_op=True))
        # pair table: partition -> all-to-all -> merge
        w0 = time.perf_counter()
        n_local = eng.pair_count()
        send = torch.empty((max(1, n_local), 3), dtype=torch.int64, device="cuda")
        counts = eng.partition_pairs_device(send.data_ptr(), n_local, G)
        w1 = time.perf_counter()
        ev[1].record()
        for w in pending:
            w.wait()
        if not hn and G > 1:
            # hashed histogram (huge lattices): records through the host
            gc = all_gather_records(eng.coords())
            for r in range(G):
                if r != self.rank:
                    eng.merge(np.zeros(0, PAIR), gc[r])
        ev[2].record()
        send_counts = torch.from_numpy(counts.astype(np.int64)).cuda()
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        in_split = counts.astype(np.int64).tolist()
        out_split = recv_counts.cpu().tolist()
        n_recv = int(sum(out_split))
        recv = torch.empty((max(1, n_recv), 3), dtype=torch.int64, device="cuda")
        dist.all_to_all_single(recv[:n_recv], send[:n_local], out_split, in_split)
        ev[3].record()
        torch.cuda.synchronize()
        eng.clear_pairs()
        eng.merge_pairs_device(recv.data_ptr(), n_recv)
        ev[4].record()
        del send, recv
        sums = self._statistics(reduced=True)
        ev[5].record()
        torch.cuda.synchronize()
        self.info = {
            "pairs_local_before": int(n_local), "pairs_received": n_recv, "pairs_after": eng.pair_count(),
            "a2a_bytes_sent": int(24 * (n_local - int(counts[self.rank]))),
            "allreduce_bytes": int(8 * n + 8 * hn),
            "partition_ms": ev[0].elapsed_time(ev[1]), "allreduce_tail_ms": ev[1].elapsed_time(ev[2]),
            "all_to_all_ms": ev[2].elapsed_time(ev[3]), "clear_and_merge_ms": ev[3].elapsed_time(ev[4]),
            "statistics_ms": ev[4].elapsed_time(ev[5]),
            "partition_host_wall_ms": 1e3 * (w1 - w0),
            "overlap": "all-reduce of the per-agent words and the histogram (NCCL stream) runs beside the partition of the "
                       "pair table (engine stream); allreduce_tail_ms is what was left of it after the partition",
        }
        return sums

    def statistics(self) -> Dict[str, int]:
        """The run's statistics sums from the rank-partitioned pair table."""
        return self._statistics(reduced=bool(self.info))

    def _statistics(self, reduced: bool) -> Dict[str, int]:
        import torch

        dist = _dist()
        eng = self.engine
        # after `reduce` the per-agent counter words are the whole run's (seconds | events << 40): the
        # sums need no pass over the pair table; n_pairs is this rank's share
        fast = None
        if reduced:
            try:
                fast = eng.statistics_counters()
            except CitamCapacityError:
                fast = None
        t = torch.tensor([1 if fast is not None else 0, fast["n_pairs"] if fast is not None else 0], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)  # [ranks whose counter path is exact, pairs over all shares]
        n_ok, n_pairs = (int(v) for v in t.tolist())
        if n_ok == self.world:
            fast["n_pairs"] = n_pairs
            return fast
        eng.statistics_begin()
        a, b, n = eng.stats_arrays_device()
        dist.all_reduce(device_tensor(a, n, "<i4"))
        dist.all_reduce(device_tensor(b, n, "<i4"))
        torch.cuda.synchronize()
        s = eng.statistics_end()
        t = torch.tensor([s["total_duration"], s["n_pairs"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        s["total_duration"], s["n_pairs"] = int(t[0].item()), int(t[1].item())
        return s

    # -- small runs: the whole tables on every rank, through the host --------------------------------
    def gather_tables(self):
        """(pairs, coords, durations) of the whole run on every rank, merged on the host from the
        ranks' tables.  Call instead of -- or after -- `reduce` (after it the per-rank tables are
        disjoint shares and the counters / histogram are already global)."""
        reduced = bool(self.info)
        pairs = merge_pair_records(all_gather_records(self.engine.pairs()))
        if reduced:
            return pairs, self.engine.coords(), self.engine.agent_durations().astype(np.int64)
        return (pairs, merge_coord_records(all_gather_records(self.engine.coords())),
                all_reduce_sum(self.engine.agent_durations()))
