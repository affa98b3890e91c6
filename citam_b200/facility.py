"""Facility geometry -> the arrays the device location-table builder consumes.

The hot path reads three things from a reference `Facility`
(citam/engine/facility/indoor_facility.py:32-81): `floorplans[f].spaces` (list order matters:
`identify_this_location` returns the first match, floorplan.py:233-241),
`number_of_floors`, and `navigation.hallways_graph_per_floor[f]`
(close_contacts_calculator.py:190-203).  `geometry_from_facility` takes them from any object
with that shape (the upstream classes, or the light stand-ins below) and produces a plain
dict (the same schema as tests/golden/*/geometry.json.gz):

  {"floors": [{"spaces": [{"name", "function", "path_bbox": [xmin,xmax,ymin,ymax],
                           "boundaries": [[sx,sy,ex,ey]...], "doors": [[sx,sy,ex,ey]...]}],
               "hallway_edges": [[name_a, name_b]...] | None}]}

`floor_tables` turns one floor of that dict into the C records of include/citam_b200.h.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from ._lib import BOUNDARY, DOOR, SPACE

HALLWAY_WORDS = ("circulation", "aisle", "lobby", "vestibule")


def is_hallway(space_function: str) -> bool:
    """Space.is_space_a_hallway (space.py:300-312)."""
    f = space_function.lower()
    return any(w in f for w in HALLWAY_WORDS)


def _seg(line) -> List[float]:
    return [line.start.real, line.start.imag, line.end.real, line.end.imag]


def geometry_from_facility(facility) -> Dict[str, Any]:
    """Duck-typed extraction from a reference-shaped Facility (or the stand-ins below)."""
    if isinstance(facility, dict) and "floors" in facility:
        return facility
    if hasattr(facility, "geometry"):
        return facility.geometry
    floors = []
    graphs = facility.navigation.hallways_graph_per_floor
    for f, fp in enumerate(facility.floorplans):
        spaces = []
        for sp in fp.spaces:
            spaces.append(
                {
                    "name": sp.unique_name,
                    "function": sp.space_function,
                    "path_bbox": [float(v) for v in sp.path.bbox()],
                    "boundaries": [_seg(l) for l in sp.boundaries],
                    "doors": [_seg(d.path) for d in sp.doors],
                }
            )
        hg = graphs[f] if f < len(graphs) else None
        edges = None if hg is None else [[a, b] for a, b in hg.edges()]
        directed = bool(hg is not None and hasattr(hg, "is_directed") and hg.is_directed())
        floors.append({"spaces": spaces, "hallway_edges": edges, "hallway_directed": directed})
    return {"floors": floors}


def floor_tables(floor_geom: Dict[str, Any]):
    """One floor -> (minx, miny, W, H, spaces[SPACE], boundaries[BOUNDARY], doors[DOOR], adjacency|None)."""
    sps = floor_geom["spaces"]
    n = len(sps)
    if n > 32767:
        raise ValueError("more than 32767 spaces on one floor")
    spaces = np.zeros(n, SPACE)
    bnd: List[tuple] = []
    doors: List[tuple] = []
    for i, sp in enumerate(sps):
        xmin, xmax, ymin, ymax = sp["path_bbox"]
        # Python's round() (banker's) exactly as space.py:210-215
        spaces["bb_xmin"][i], spaces["bb_xmax"][i] = round(xmin), round(xmax)
        spaces["bb_ymin"][i], spaces["bb_ymax"][i] = round(ymin), round(ymax)
        spaces["seg_begin"][i] = len(bnd)
        for sx, sy, ex, ey in sp["boundaries"]:
            d = complex(ex, ey) - complex(sx, sy)
            long_enough = abs(d) > 1.0  # Line.length() > 1.0 (space.py:189)
            nx = ny = 0.0
            if long_enough:
                nrm = -1j * (d / abs(d))  # svgpathtools Line.normal (geometry.py:580)
                nx, ny = nrm.real, nrm.imag
            bnd.append((sx, sy, ex, ey, nx, ny, 1 if long_enough else 0, 0))
        spaces["seg_end"][i] = len(bnd)
        spaces["door_begin"][i] = len(doors)
        if not is_hallway(sp["function"]) and len(sp["doors"]) > 0:  # space.py:218-224
            for d in sp["doors"]:
                doors.append(tuple(d))
        spaces["door_end"][i] = len(doors)
    boundaries = np.array(bnd, BOUNDARY) if bnd else np.zeros(0, BOUNDARY)
    door_arr = np.array(doors, DOOR) if doors else np.zeros(0, DOOR)
    if n:
        minx, maxx = int(spaces["bb_xmin"].min()), int(spaces["bb_xmax"].max())
        miny, maxy = int(spaces["bb_ymin"].min()), int(spaces["bb_ymax"].max())
    else:
        minx = maxx = miny = maxy = 0
    W, H = maxx - minx + 1, maxy - miny + 1
    edges = floor_geom.get("hallway_edges")
    adj = None
    if edges is not None:
        hall = np.array([is_hallway(sp["function"]) for sp in sps], bool)
        adj = np.zeros((n, n), np.uint8)
        by_name: Dict[str, List[int]] = {}
        for i, sp in enumerate(sps):
            if hall[i]:
                by_name.setdefault(str(sp["name"]), []).append(i)
        for a, b in edges:  # undirected has_edge(name_a, name_b)
            for ia in by_name.get(str(a), ()):
                for ib in by_name.get(str(b), ()):
                    adj[ia, ib] = 1
                    if not floor_geom.get("hallway_directed", False):
                        adj[ib, ia] = 1
    return minx, miny, W, H, spaces, boundaries, door_arr, adj


# --------------------------------------------------------------------------
# light stand-ins with the attribute names the hot path reads from upstream
# --------------------------------------------------------------------------
class GeometryFacility:
    """A Facility made of a geometry dict (tests, synthetic facilities, golden fixtures).

    Exposes what Simulation / CloseContactsCalculator read from the upstream Facility:
    `number_of_floors`, `floorplans` (len only), `facility_name`, `traffic_policy`.
    """

    def __init__(self, geometry: Dict[str, Any], facility_name: str = "synthetic", widths: Optional[List[float]] = None):
        self.geometry = geometry
        self.facility_name = facility_name
        self.number_of_floors = len(geometry["floors"])
        self.floorplans = [_FloorStub(f, fg) for f, fg in enumerate(geometry["floors"])]
        self.traffic_policy = None
        self.entrances: list = []


class _FloorStub:
    def __init__(self, index: int, fg: Dict[str, Any]):
        self.floor_name = str(index)
        bbs = np.array([s["path_bbox"] for s in fg["spaces"]], float).reshape(-1, 4)
        self.minx = fg.get("minx", float(bbs[:, 0].min()) if len(bbs) else 0.0)
        self.maxx = fg.get("maxx", float(bbs[:, 1].max()) if len(bbs) else 0.0)
        self.miny = fg.get("miny", float(bbs[:, 2].min()) if len(bbs) else 0.0)
        self.maxy = fg.get("maxy", float(bbs[:, 3].max()) if len(bbs) else 0.0)
        self.n_spaces = len(fg["spaces"])
