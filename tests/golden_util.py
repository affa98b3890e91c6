"""Helpers to read the committed golden fixtures (tests/golden/, made by oracle/make_golden.py)."""
import gzip
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["basic_seed0", "twofloor", "crafted"]


def load_case(name):
    d = os.path.join(GOLD, name)
    z = np.load(os.path.join(d, "inputs.npz"))
    with gzip.open(os.path.join(d, "geometry.json.gz"), "rt") as f:
        geom = json.load(f)
    off, x, y, fl = z["itin_off"], z["itin_x"], z["itin_y"], z["itin_floor"]
    itins = []
    for a in range(len(off) - 1):
        it = []
        for k in range(off[a], off[a + 1]):
            it.append(((int(x[k]), int(y[k])), int(fl[k])) if fl[k] >= 0 else (None, None))
        itins.append(it)
    return {
        "name": name,
        "dir": d,
        "geometry": geom,
        "itineraries": itins,
        "T": int(z["total_timesteps"]),
        "D": float(z["contact_distance"]),
        "loc_points": z["loc_points"],
        "loc_values": z["loc_values"],
        "packed": (off, x, y, fl),
    }


def expected(case, relpath):
    with gzip.open(os.path.join(case["dir"], "expected", relpath + ".gz"), "rb") as f:
        return f.read()
