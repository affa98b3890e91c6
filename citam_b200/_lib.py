"""ctypes binding of libcitam_b200.so (the C ABI declared in include/citam_b200.h).

There is no CPU fallback: if the shared library is missing, or the machine has
no CUDA device, the calls raise `CitamDeviceError`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcitam_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_STATE, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
FLAG_TIMING = 1
FLAG_HASHED_COORDS = 2
N_KERNELS = 9
KERNEL_NAMES = ["expand_bin", "scan", "scatter", "pairs", "lut", "finalize", "misc", "accumulate", "classify"]


class CitamDeviceError(RuntimeError):
    """The CUDA library is missing or a device call failed."""

    def __init__(self, code, message):
        super().__init__("citam_b200 [%d]: %s" % (code, message))
        self.code = code


class CitamCapacityError(CitamDeviceError):
    """A device table overflowed; re-run with larger capacities."""


class Config(C.Structure):
    _fields_ = [
        ("n_agents", C.c_int32),
        ("n_floors", C.c_int32),
        ("contact_distance", C.c_double),
        ("cell_size", C.c_int32),
        ("device", C.c_int32),
        ("pair_capacity", C.c_uint64),
        ("coord_capacity", C.c_uint64),
        ("log_capacity", C.c_uint64),
        ("steps_per_chunk", C.c_int32),
        ("flags", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "agent_steps", "pair_tests", "contact_steps", "kernel_launches",
            "pair_slots_used", "coord_slots_used", "log_records", "overflow", "contact_runs", "binned_agent_steps",
        )
    ]


class Stats(C.Structure):
    _fields_ = [
        (n, C.c_uint64)
        for n in (
            "total_duration", "n_pairs", "n_agents_with_contacts",
            "sum_events_per_agent", "max_events_per_agent", "reserved",
        )
    ]


class Timings(C.Structure):
    _fields_ = [("ms", C.c_double * N_KERNELS), ("launches", C.c_uint64 * N_KERNELS)]


# numpy mirrors of the C records (sizes asserted against the header in tests)
SEGMENT = np.dtype(
    [("t0", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("dx", "<i2"), ("dy", "<i2"), ("floor", "<i2"), ("reserved", "<i2")]
)
AGENT_STATE = np.dtype([("x", "<i4"), ("y", "<i4"), ("floor", "<i2"), ("loc", "<i2")])
SPACE = np.dtype(
    [("bb_xmin", "<i4"), ("bb_xmax", "<i4"), ("bb_ymin", "<i4"), ("bb_ymax", "<i4"),
     ("seg_begin", "<i4"), ("seg_end", "<i4"), ("door_begin", "<i4"), ("door_end", "<i4")]
)
BOUNDARY = np.dtype(
    [("sx", "<f8"), ("sy", "<f8"), ("ex", "<f8"), ("ey", "<f8"), ("nx", "<f8"), ("ny", "<f8"),
     ("long_enough", "<i4"), ("reserved", "<i4")]
)
DOOR = np.dtype([("sx", "<f8"), ("sy", "<f8"), ("ex", "<f8"), ("ey", "<f8")])
PAIR = np.dtype([("a1", "<u4"), ("a2", "<u4"), ("first_step", "<u4"), ("n_events", "<u4"), ("duration", "<u8")])
COORD = np.dtype([("floor", "<i4"), ("sx", "<i4"), ("sy", "<i4"), ("reserved", "<i4"), ("count", "<u8")])
CONTACT = np.dtype([("step", "<u4"), ("a1", "<u4"), ("a2", "<u4")])

EXPORTS = [
    "citam_abi_version", "citam_last_error", "citam_create", "citam_destroy", "citam_set_stream",
    "citam_build_floor_lut", "citam_set_floor_lut", "citam_get_floor_lut", "citam_set_floor_adjacency",
    "citam_load_itineraries", "citam_run", "citam_step_states", "citam_states_at", "citam_states_range", "citam_reset",
    "citam_pair_count", "citam_export_pairs", "citam_export_agent_durations", "citam_coord_count",
    "citam_export_coords", "citam_export_contact_log", "citam_clear_contact_log", "citam_statistics", "citam_get_counters",
    "citam_get_timings", "citam_agent_durations_device", "citam_merge_pairs", "citam_merge_coords",
    "citam_hist_device", "citam_export_pairs_device", "citam_export_coords_device", "citam_partition_pairs_device",
    "citam_clear_pairs", "citam_merge_pairs_device", "citam_statistics_begin", "citam_stats_arrays_device",
    "citam_statistics_end", "citam_statistics_counters", "citam_measure_random_updates",
]
ABI_VERSION = 3

_lib = None


def load() -> C.CDLL:
    """Load the shared library or fail loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CitamDeviceError(
            ERR_NO_DEVICE,
            "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU fallback." % LIB_PATH,
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    pu64 = C.POINTER(C.c_uint64)
    sig = {
        "citam_abi_version": (C.c_int, []),
        "citam_last_error": (C.c_char_p, []),
        "citam_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "citam_destroy": (None, [vp]),
        "citam_set_stream": (C.c_int, [vp, vp]),
        "citam_build_floor_lut": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
        "citam_set_floor_lut": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp]),
        "citam_get_floor_lut": (C.c_int, [vp, i32, vp]),
        "citam_set_floor_adjacency": (C.c_int, [vp, i32, i32, vp]),
        "citam_load_itineraries": (C.c_int, [vp, vp, vp, C.c_int64]),
        "citam_run": (C.c_int, [vp, i32, i32]),
        "citam_step_states": (C.c_int, [vp, i32, vp]),
        "citam_states_at": (C.c_int, [vp, i32, vp]),
        "citam_states_range": (C.c_int, [vp, i32, i32, vp]),
        "citam_reset": (C.c_int, [vp]),
        "citam_pair_count": (C.c_int, [vp, pu64]),
        "citam_export_pairs": (C.c_int, [vp, vp, u64, pu64]),
        "citam_export_agent_durations": (C.c_int, [vp, vp]),
        "citam_coord_count": (C.c_int, [vp, pu64]),
        "citam_export_coords": (C.c_int, [vp, vp, u64, pu64]),
        "citam_export_contact_log": (C.c_int, [vp, vp, u64, pu64]),
        "citam_clear_contact_log": (C.c_int, [vp]),
        "citam_statistics": (C.c_int, [vp, C.POINTER(Stats)]),
        "citam_get_counters": (C.c_int, [vp, C.POINTER(Counters)]),
        "citam_get_timings": (C.c_int, [vp, C.POINTER(Timings)]),
        "citam_agent_durations_device": (C.c_int, [vp, C.POINTER(vp), pu64]),
        "citam_merge_pairs": (C.c_int, [vp, vp, u64]),
        "citam_merge_coords": (C.c_int, [vp, vp, u64]),
        "citam_hist_device": (C.c_int, [vp, C.POINTER(vp), pu64]),
        "citam_export_pairs_device": (C.c_int, [vp, vp, u64, pu64]),
        "citam_export_coords_device": (C.c_int, [vp, vp, u64, pu64]),
        "citam_partition_pairs_device": (C.c_int, [vp, vp, u64, i32, pu64]),
        "citam_clear_pairs": (C.c_int, [vp]),
        "citam_merge_pairs_device": (C.c_int, [vp, vp, u64]),
        "citam_statistics_begin": (C.c_int, [vp]),
        "citam_stats_arrays_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), pu64]),
        "citam_statistics_end": (C.c_int, [vp, C.POINTER(Stats)]),
        "citam_statistics_counters": (C.c_int, [vp, C.POINTER(Stats)]),
        "citam_measure_random_updates": (C.c_int, [C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.citam_abi_version() != ABI_VERSION:
        raise CitamDeviceError(ERR_STATE, "%s is ABI %d, this package expects %d: rebuild it" % (LIB_PATH, lib.citam_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = load().citam_last_error().decode("utf-8", "replace")
    if rc == ERR_CAPACITY:
        raise CitamCapacityError(rc, msg)
    if rc == ERR_INVALID:
        raise ValueError("citam_b200: " + msg)
    raise CitamDeviceError(rc, msg)


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)
