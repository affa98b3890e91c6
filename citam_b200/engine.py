"""DeviceEngine — Python handle on one `citam_engine` (one GPU).

Thin layer over the C ABI (include/citam_b200.h): numpy in, numpy out.  All compute happens in
libcitam_b200.so; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence

import numpy as np

from . import _lib as L
from . import facility as fac
from . import itinerary as itin


class DeviceEngine:
    def __init__(
        self,
        n_agents: int,
        n_floors: int,
        contact_distance: float = 6.0,
        *,
        device: int = 0,
        cell_size: int = 0,
        pair_capacity: int = 0,
        coord_capacity: int = 0,
        log_capacity: int = 0,
        steps_per_chunk: int = 0,
        timing: bool = False,
        hashed_coords: bool = False,
    ):
        self.lib = L.load()
        self.n_agents = int(n_agents)
        self.n_floors = int(n_floors)
        self.contact_distance = float(contact_distance)
        cfg = L.Config(
            n_agents=self.n_agents, n_floors=self.n_floors, contact_distance=self.contact_distance,
            cell_size=cell_size, device=device, pair_capacity=pair_capacity, coord_capacity=coord_capacity,
            log_capacity=log_capacity, steps_per_chunk=steps_per_chunk,
            flags=(L.FLAG_TIMING if timing else 0) | (L.FLAG_HASHED_COORDS if hashed_coords else 0),
        )
        self._cfg = cfg
        h = C.c_void_p()
        L.check(self.lib.citam_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.floor_meta: Dict[int, tuple] = {}

    # -- life cycle ---------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.citam_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int) -> None:
        L.check(self.lib.citam_set_stream(self._h, C.c_void_p(cuda_stream)))

    # -- facility -----------------------------------------------------------
    def load_geometry(self, geometry: Dict[str, Any]) -> None:
        """Build every floor's location table on the device and install hallway adjacency."""
        geometry = fac.geometry_from_facility(geometry)
        if len(geometry["floors"]) != self.n_floors:
            raise ValueError("geometry has %d floors, engine %d" % (len(geometry["floors"]), self.n_floors))
        for f, fg in enumerate(geometry["floors"]):
            minx, miny, W, H, spaces, bnd, doors, adj = fac.floor_tables(fg)
            L.check(
                self.lib.citam_build_floor_lut(
                    self._h, f, minx, miny, W, H, len(spaces), L.ptr(spaces), len(bnd), L.ptr(bnd), len(doors), L.ptr(doors)
                )
            )
            self.floor_meta[f] = (minx, miny, W, H, len(spaces))
            self.set_adjacency(f, adj)

    def set_floor_lut(self, floor: int, minx: int, miny: int, lut: np.ndarray, n_spaces: int) -> None:
        lut = np.ascontiguousarray(lut, np.int16)
        H, W = lut.shape
        L.check(self.lib.citam_set_floor_lut(self._h, floor, minx, miny, W, H, n_spaces, L.ptr(lut)))
        self.floor_meta[floor] = (minx, miny, W, H, n_spaces)

    def set_adjacency(self, floor: int, adj: Optional[np.ndarray]) -> None:
        if adj is None:
            L.check(self.lib.citam_set_floor_adjacency(self._h, floor, 0, None))
        else:
            adj = np.ascontiguousarray(adj, np.uint8)
            L.check(self.lib.citam_set_floor_adjacency(self._h, floor, adj.shape[0], L.ptr(adj)))

    def floor_lut(self, floor: int) -> np.ndarray:
        minx, miny, W, H, _ = self.floor_meta[floor]
        out = np.empty((H, W), np.int16)
        L.check(self.lib.citam_get_floor_lut(self._h, floor, L.ptr(out)))
        return out

    def locate(self, floor: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        """Space index (or -1) of points, read from the device-built table."""
        minx, miny, W, H, _ = self.floor_meta[floor]
        lut = self.floor_lut(floor)
        fx, fy = np.asarray(x) - minx, np.asarray(y) - miny
        ok = (fx >= 0) & (fx < W) & (fy >= 0) & (fy < H)
        out = np.full(fx.shape, -1, np.int32)
        out[ok] = lut[fy[ok], fx[ok]]
        return out

    # -- itineraries ----------------------------------------------------------
    def load_itineraries(self, itineraries: Sequence[Sequence]) -> None:
        self.load_segments(*itin.encode(itineraries))

    def load_segments(self, seg_off: np.ndarray, segs: np.ndarray) -> None:
        seg_off = np.ascontiguousarray(seg_off, np.int64)
        segs = np.ascontiguousarray(segs, L.SEGMENT)
        if len(seg_off) != self.n_agents + 1:
            raise ValueError("seg_off must have n_agents+1 entries")
        L.check(self.lib.citam_load_itineraries(self._h, L.ptr(seg_off), L.ptr(segs), len(segs)))

    def load_segments_raw(self, seg_off_ptr: int, segs_ptr: int, n_segs: int) -> None:
        """Same as load_segments for caller-owned (e.g. pinned) host buffers given by address."""
        L.check(self.lib.citam_load_itineraries(self._h, C.c_void_p(seg_off_ptr), C.c_void_p(segs_ptr), int(n_segs)))

    def agent_durations_into(self, out_ptr: int) -> None:
        """Per-agent contact seconds into a caller-owned uint64[n_agents] host buffer."""
        L.check(self.lib.citam_export_agent_durations(self._h, C.c_void_p(out_ptr)))

    # -- stepping -------------------------------------------------------------
    def run(self, t_begin: int, t_end: int) -> None:
        L.check(self.lib.citam_run(self._h, int(t_begin), int(t_end)))

    def step_states(self, step: int, states: np.ndarray) -> None:
        states = np.ascontiguousarray(states, L.AGENT_STATE)
        if len(states) != self.n_agents:
            raise ValueError("states must have n_agents records")
        L.check(self.lib.citam_step_states(self._h, int(step), L.ptr(states)))

    def states_at(self, step: int) -> np.ndarray:
        out = np.empty(self.n_agents, L.AGENT_STATE)
        L.check(self.lib.citam_states_at(self._h, int(step), L.ptr(out)))
        return out

    def states_range(self, t_begin: int, t_end: int) -> np.ndarray:
        """Agent states for every step of [t_begin, t_end): array [t_end - t_begin, n_agents]."""
        out = np.empty((max(0, t_end - t_begin), self.n_agents), L.AGENT_STATE)
        if out.size:
            L.check(self.lib.citam_states_range(self._h, int(t_begin), int(t_end), L.ptr(out)))
        return out

    def reset(self) -> None:
        L.check(self.lib.citam_reset(self._h))

    # -- results ----------------------------------------------------------------
    def pairs(self) -> np.ndarray:
        n = C.c_uint64()
        L.check(self.lib.citam_pair_count(self._h, C.byref(n)))
        out = np.empty(n.value, L.PAIR)
        L.check(self.lib.citam_export_pairs(self._h, L.ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def coords(self) -> np.ndarray:
        n = C.c_uint64()
        L.check(self.lib.citam_coord_count(self._h, C.byref(n)))
        out = np.empty(n.value, L.COORD)
        L.check(self.lib.citam_export_coords(self._h, L.ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def agent_durations(self) -> np.ndarray:
        out = np.empty(self.n_agents, np.uint64)
        L.check(self.lib.citam_export_agent_durations(self._h, L.ptr(out)))
        return out

    def contact_log(self) -> np.ndarray:
        c = self.counters()
        out = np.empty(c["log_records"], L.CONTACT)
        n = C.c_uint64()
        L.check(self.lib.citam_export_contact_log(self._h, L.ptr(out), len(out), C.byref(n)))
        return out[: n.value]

    def clear_contact_log(self) -> None:
        L.check(self.lib.citam_clear_contact_log(self._h))

    @property
    def log_enabled(self) -> bool:
        return self._cfg.log_capacity > 0

    def statistics_sums(self) -> Dict[str, int]:
        s = L.Stats()
        L.check(self.lib.citam_statistics(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in L.Stats._fields_ if k != "reserved"}

    def counters(self) -> Dict[str, int]:
        c = L.Counters()
        L.check(self.lib.citam_get_counters(self._h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in L.Counters._fields_}

    def timings(self) -> Dict[str, Dict[str, float]]:
        t = L.Timings()
        L.check(self.lib.citam_get_timings(self._h, C.byref(t)))
        return {
            L.KERNEL_NAMES[k]: {"ms": float(t.ms[k]), "launches": int(t.launches[k])}
            for k in range(L.N_KERNELS)
            if t.launches[k]
        }

    def merge(self, pairs: np.ndarray, coords: np.ndarray) -> None:
        """Other ranks' records into this engine's tables (pair table and histogram only: the
        per-agent contact seconds are reduced separately)."""
        pairs = np.ascontiguousarray(pairs, L.PAIR)
        coords = np.ascontiguousarray(coords, L.COORD)
        L.check(self.lib.citam_merge_pairs(self._h, L.ptr(pairs), len(pairs)))
        L.check(self.lib.citam_merge_coords(self._h, L.ptr(coords), len(coords)))

    # -- device-resident results (multi-GPU exchange, large runs) ------------------------------
    def agent_durations_device(self):
        """(device pointer, n) of the live uint64 per-agent counter words (contact seconds in bits 0-39,
        contact events above), for the NCCL all-reduce; `agent_durations()` returns the seconds."""
        p = C.c_void_p()
        n = C.c_uint64()
        L.check(self.lib.citam_agent_durations_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def hist_device(self):
        """(device pointer, n) of the dense uint64 coordinate histogram; (None, 0) when hashed."""
        p = C.c_void_p()
        n = C.c_uint64()
        L.check(self.lib.citam_hist_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stats_arrays_device(self):
        """(events-per-agent pointer, has-contact pointer, n): uint32 arrays filled by statistics_begin."""
        a, b = C.c_void_p(), C.c_void_p()
        n = C.c_uint64()
        L.check(self.lib.citam_stats_arrays_device(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def pair_count(self) -> int:
        n = C.c_uint64()
        L.check(self.lib.citam_pair_count(self._h, C.byref(n)))
        return int(n.value)

    def export_pairs_device(self, dev_ptr: int, capacity: int) -> int:
        """Pair records in first-contact order into a caller-owned DEVICE buffer; returns the count."""
        n = C.c_uint64()
        L.check(self.lib.citam_export_pairs_device(self._h, C.c_void_p(dev_ptr), int(capacity), C.byref(n)))
        return int(n.value)

    def export_coords_device(self, dev_ptr: int, capacity: int) -> int:
        n = C.c_uint64()
        L.check(self.lib.citam_export_coords_device(self._h, C.c_void_p(dev_ptr), int(capacity), C.byref(n)))
        return int(n.value)

    def partition_pairs_device(self, dev_ptr: int, capacity: int, n_parts: int) -> np.ndarray:
        """All pair records grouped by hash(a1, a2) mod n_parts into a DEVICE buffer; returns the
        per-part counts (uint64[n_parts])."""
        counts = np.zeros(n_parts, np.uint64)
        L.check(self.lib.citam_partition_pairs_device(self._h, C.c_void_p(dev_ptr), int(capacity), int(n_parts),
                                                      counts.ctypes.data_as(C.POINTER(C.c_uint64))))
        return counts

    def clear_pairs(self) -> None:
        L.check(self.lib.citam_clear_pairs(self._h))

    def merge_pairs_device(self, dev_ptr: int, n: int) -> None:
        L.check(self.lib.citam_merge_pairs_device(self._h, C.c_void_p(dev_ptr), int(n)))

    def statistics_begin(self) -> None:
        L.check(self.lib.citam_statistics_begin(self._h))

    def statistics_counters(self) -> Dict[str, int]:
        """Statistics sums from the per-agent counter words + the table's pair count (O(n_agents));
        raises CitamDeviceError(ERR_CAPACITY) when an agent's event field is not provably exact."""
        s = L.Stats()
        L.check(self.lib.citam_statistics_counters(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in L.Stats._fields_ if k != "reserved"}

    def statistics_end(self) -> Dict[str, int]:
        s = L.Stats()
        L.check(self.lib.citam_statistics_end(self._h, C.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in L.Stats._fields_ if k != "reserved"}


def measure_random_updates(table_bytes: int, n_updates: int, probe_first: bool = True, device: int = 0) -> float:
    """Million random 8-byte read-modify-writes per second the device sustains over a table of `table_bytes`
    (the ceiling for k_accumulate's access pattern); measurement aid, see the header."""
    ms = C.c_double()
    L.check(L.load().citam_measure_random_updates(device, table_bytes, n_updates, 1 if probe_first else 0, C.byref(ms)))
    return n_updates / (ms.value * 1e-3) / 1e6


def statistics_from_sums(s: Dict[str, int]):
    """The six statistics of ContactEvents.extract_statistics (contacts.py:185-294) from the
    device's integer sums, with Python's own division and round()."""
    n_with = s["n_agents_with_contacts"]
    avg_n = s["sum_events_per_agent"] / n_with if n_with > 0 else 0
    avg_d = (2 * s["total_duration"]) / (n_with * 2.0) if n_with > 0 else 0
    avg_p = s["n_pairs"] / n_with if n_with > 0 else 0
    return [
        {"name": "overall_total_contact_duration", "value": round(s["total_duration"] / 60.0, 2), "unit": "min"},
        {"name": "avg_n_contacts_per_agent", "value": round(avg_n, 2), "unit": ""},
        {"name": "avg_contact_duration_per_agent", "value": round(avg_d / 60.0, 2), "unit": "min"},
        {"name": "n_agents_with_contacts", "value": n_with, "unit": ""},
        {"name": "avg_number_of_people_per_agent", "value": round(avg_p, 2), "unit": ""},
        {"name": "max_contacts", "value": s["max_events_per_agent"], "unit": ""},
    ]
