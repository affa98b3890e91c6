"""ctypes wrapper of oracle/grid_oracle.c (the C grid restatement of the step loop).

TEST INFRASTRUCTURE ONLY — see the header of grid_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libgrid_oracle.so")

PAIR = np.dtype([("a1", "<u4"), ("a2", "<u4"), ("first_step", "<u4"), ("n_events", "<u4"), ("duration", "<u8")])
COORD = np.dtype([("floor", "<i4"), ("sx", "<i4"), ("sy", "<i4"), ("reserved", "<i4"), ("count", "<u8")])
SEGMENT = np.dtype(
    [("t0", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("dx", "<i2"), ("dy", "<i2"), ("floor", "<i2"), ("reserved", "<i2")]
)


class _Floor(C.Structure):
    _fields_ = [("minx", C.c_int32), ("miny", C.c_int32), ("W", C.c_int32), ("H", C.c_int32),
                ("n_spaces", C.c_int32), ("lut", C.c_void_p), ("adj", C.c_void_p)]


_lib = None


def build() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "grid_oracle.c")
        if not os.path.isfile(LIB) or (os.path.isfile(src) and os.path.getmtime(LIB) < os.path.getmtime(src)):
            build()
        lib = C.CDLL(LIB)
        vp = C.c_void_p
        lib.go_create.restype = vp
        lib.go_create.argtypes = [C.c_int32, C.c_int32, C.c_double, C.c_int32, vp, vp, vp]
        lib.go_destroy.argtypes = [vp]
        lib.go_run.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32]
        for n in ("go_pair_count", "go_coord_count", "go_pair_tests", "go_contact_steps"):
            getattr(lib, n).restype = C.c_uint64
            getattr(lib, n).argtypes = [vp]
        lib.go_export_pairs.argtypes = [vp, vp]
        lib.go_export_coords.argtypes = [vp, vp]
        lib.go_export_durations.argtypes = [vp, vp]
        lib.go_build_lut.argtypes = [C.c_int32] * 5 + [vp, vp, vp, vp, C.c_int32]
        lib.go_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return int(load().go_max_threads())


def build_lut(floor_tables, threads: int = 0) -> np.ndarray:
    """The reference's location rule over the floor lattice, in C (floorplan.py:215-241)."""
    minx, miny, W, H, spaces, bnd, doors, _ = floor_tables
    out = np.empty((H, W), np.int16)
    spaces = np.ascontiguousarray(spaces)
    bnd = np.ascontiguousarray(bnd)
    doors = np.ascontiguousarray(doors)
    load().go_build_lut(minx, miny, W, H, len(spaces), spaces.ctypes.data, bnd.ctypes.data, doors.ctypes.data,
                        out.ctypes.data, threads or max_threads())
    return out


class GridOracle:
    """Step loop on the CPU.  `floors` is a list of (minx, miny, lut[H,W] int16, n_spaces, adj|None)."""

    def __init__(self, n_agents, contact_distance, floors, seg_off, segs, cell_size: int = 0):
        self.lib = load()
        self.n_agents = int(n_agents)
        self._keep = []
        arr = (_Floor * len(floors))()
        for i, (minx, miny, lut, n_spaces, adj) in enumerate(floors):
            lut = np.ascontiguousarray(lut, np.int16)
            self._keep.append(lut)
            a_ptr = None
            if adj is not None:
                adj = np.ascontiguousarray(adj, np.uint8)
                self._keep.append(adj)
                a_ptr = adj.ctypes.data
            arr[i] = _Floor(minx, miny, lut.shape[1], lut.shape[0], n_spaces, lut.ctypes.data, a_ptr)
        self.seg_off = np.ascontiguousarray(seg_off, np.int64)
        self.segs = np.ascontiguousarray(segs, SEGMENT)
        self._h = self.lib.go_create(self.n_agents, len(floors), float(contact_distance), int(cell_size),
                                     C.cast(arr, C.c_void_p), self.seg_off.ctypes.data, self.segs.ctypes.data)

    @classmethod
    def from_geometry(cls, geometry, contact_distance, seg_off, segs, floor_tables_fn, threads: int = 0):
        floors = []
        for fg in geometry["floors"]:
            tabs = floor_tables_fn(fg)
            lut = build_lut(tabs, threads)
            floors.append((tabs[0], tabs[1], lut, len(tabs[4]), tabs[7]))
        n = len(seg_off) - 1
        return cls(n, contact_distance, floors, seg_off, segs)

    def run(self, t_begin: int, t_end: int, threads: int = 1) -> None:
        self.lib.go_run(self._h, int(t_begin), int(t_end), int(threads))

    def pairs(self) -> np.ndarray:
        out = np.empty(self.lib.go_pair_count(self._h), PAIR)
        self.lib.go_export_pairs(self._h, out.ctypes.data)
        return out

    def coords(self) -> np.ndarray:
        out = np.empty(self.lib.go_coord_count(self._h), COORD)
        self.lib.go_export_coords(self._h, out.ctypes.data)
        return out

    def agent_durations(self) -> np.ndarray:
        out = np.empty(self.n_agents, np.uint64)
        self.lib.go_export_durations(self._h, out.ctypes.data)
        return out

    @property
    def pair_tests(self) -> int:
        return int(self.lib.go_pair_tests(self._h))

    @property
    def contact_steps(self) -> int:
        return int(self.lib.go_contact_steps(self._h))

    def close(self) -> None:
        if self._h:
            self.lib.go_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
