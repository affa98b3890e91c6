"""map.svg / heatmap.svg for the dashboard.

The reference draws `floor_k/map.svg` with svgpathtools' `wsvg` (`export_world_to_svg`,
citam/engine/io/visualization.py:298-443; every wall a `<path d="M x,y L x,y" fill="white"
stroke="black" stroke-width=3.0*round(xmax/1500, 1)>`), then `add_root_layer_to_svg` (:446-473)
moves all children (incl. svgwrite's `<defs />`) into `<g id="root">`; `floor_k/heatmap.svg` is that file
plus a blur filter and one circle per contact coordinate whose `fill` is matplotlib's RdYlGn colour
at `1 - v / max` -- written as the *string of the RGBA tuple* (close_contacts_calculator.py:320-377:
the tuple is handed to ElementTree unformatted).  Neither svgpathtools nor matplotlib is a dependency
here: the map is written directly, byte-compatible with what svgwrite + ElementTree produce (pinned
by tests against the reference's own example files), and the colour map is rebuilt from its eleven
ColorBrewer anchors with matplotlib's lookup-table arithmetic.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

_RDYLGN = [
    (165, 0, 38), (215, 48, 39), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
    (217, 239, 139), (166, 217, 106), (102, 189, 99), (26, 152, 80), (0, 104, 55),
]
_NS = "http://www.w3.org/2000/svg"
_LUT = None


def _lut() -> np.ndarray:
    """matplotlib's 256-entry RdYlGn table: `LinearSegmentedColormap.from_list` over the eleven
    anchors (`matplotlib/_cm.py` _RdYlGn_data = k/255) through `colors._create_lookup_table`."""
    global _LUT
    if _LUT is None:
        N = 256
        anchors = np.array(_RDYLGN, np.float64) / 255.0
        x = np.linspace(0, 1, len(anchors)) * (N - 1)
        xind = (N - 1) * np.linspace(0, 1, N)
        ind = np.searchsorted(x, xind)[1:-1]
        distance = (xind[1:-1] - x[ind - 1]) / (x[ind] - x[ind - 1])
        cols = []
        for c in range(3):
            y = anchors[:, c]
            cols.append(np.clip(np.concatenate([[y[0]], distance * (y[ind] - y[ind - 1]) + y[ind - 1], [y[-1]]]), 0.0, 1.0))
        _LUT = np.stack(cols + [np.ones(N)], axis=1)
    return _LUT


def rdylgn_rgba(v: float) -> Tuple[float, float, float, float]:
    """`cm.get_cmap("RdYlGn")(v)` for a float v: the RGBA tuple of table entry int(v * 256), v == 1 -> 255."""
    lut = _lut()
    v = float(v)
    if not v == v:  # NaN: matplotlib's "bad" colour
        return (0.0, 0.0, 0.0, 0.0)
    i = int(np.floor(v * 256))
    if v * 256 == 256:
        i = 255
    i = min(255, max(0, i))  # below 0 / above 1: the under / over colours are the end entries
    return tuple(float(c) for c in lut[i])


def rdylgn_hex(v: float) -> str:
    """The same colour as #rrggbb (`matplotlib.colors.to_hex`)."""
    r, g, b, _ = rdylgn_rgba(v)
    return "#%02x%02x%02x" % tuple(int(round(c * 255)) for c in (r, g, b))


def path_d(points: Sequence[Tuple[float, float]]) -> str:
    """svgpathtools `Path.d()` of a polyline: `M x0,y0 L x1,y1 L ...` with float formatting."""
    pts = [(float(x), float(y)) for x, y in points]
    return "M %s,%s" % pts[0] + "".join(" L %s,%s" % p for p in pts[1:])


def wall_paths_of_floorplan(floorplan) -> List[str]:
    """`floorplan.walls` (svgpathtools Paths of Lines, as upstream hands them to `wsvg`) as d strings."""
    out = []
    for wall in floorplan.walls:
        segs = list(wall) if not hasattr(wall, "start") or hasattr(wall, "__iter__") else [wall]
        pts, prev_end, d = [], None, ""
        for seg in segs:
            if prev_end is None or seg.start != prev_end:
                if pts:
                    d += (" " if d else "") + path_d(pts)
                pts = [(seg.start.real, seg.start.imag)]
            pts.append((seg.end.real, seg.end.imag))
            prev_end = seg.end
        if pts:
            d += (" " if d else "") + path_d(pts)
        out.append(d)
    return out


def floor_walls(floor_geometry: Dict) -> List[str]:
    """A geometry dict carries no wall list: every boundary segment of every space, once."""
    seen, out = set(), []
    for sp in floor_geometry["spaces"]:
        for seg in sp["boundaries"]:
            k = tuple(seg)
            kr = (seg[2], seg[3], seg[0], seg[1])
            if k not in seen and kr not in seen:
                seen.add(k)
                out.append(path_d([(seg[0], seg[1]), (seg[2], seg[3])]))
    return out


def write_map_svg(walls: Iterable[str], svg_file: str, viewbox: Sequence[float]) -> None:
    """The file `export_world_to_svg(walls, [], svg_file, viewbox=...)` leaves behind
    (visualization.py:298-473 as called from simulation.py:432-458): viewBox
    `minx-50 miny-50 maxx+50 maxy+50`, stroke width scaled by round(xmax / 1500, 1)."""
    xmin, ymin, xmax, ymax = viewbox
    multiplier = round(xmax / 1500, 1)
    viewbox_str = str(xmin) + " " + str(ymin) + " " + str(xmax) + " " + str(ymax)
    sw = str(3.0 * multiplier)
    parts = ['<svg xmlns="%s" baseProfile="full" height="100%%" version="1.1" viewBox="%s" width="100%%">\n\t' % (_NS, viewbox_str),
             '<g id="root"><defs />']
    for d in walls:
        parts.append('\n\t<path d="%s" fill="white" stroke="black" stroke-width="%s" />' % (d, sw))
    parts.append("\n</g></svg>")
    with open(svg_file, "w") as f:
        f.write("".join(parts))


def write_heatmap_svg(contacts_per_coord: Dict[Tuple[float, float], int], floor_directory: str, contact_distance: float) -> None:
    """close_contacts_calculator.py:320-377: map.svg + blur filter + one circle per coordinate."""
    map_file = os.path.join(floor_directory, "map.svg")
    if not os.path.isfile(map_file):
        raise FileNotFoundError(map_file)
    ET.register_namespace("", _NS)
    tree = ET.parse(map_file)
    root = tree.getroot()[0]
    filt = ET.Element("filter", {"id": "blurMe", "width": str(contact_distance * 2), "height": str(contact_distance * 2),
                                 "x": "-100%", "y": "-100%"})
    filt.append(ET.Element("feGaussianBlur", {"in": "SourceGraphic", "stdDeviation": str(contact_distance * 0.4),
                                              "edgeMode": "duplicate"}))
    root.append(filt)
    max_contacts = max(contacts_per_coord.values()) if contacts_per_coord else 0
    for (x, y), v in contacts_per_coord.items():
        root.append(ET.Element("circle", {
            "cx": str(x), "cy": str(y), "r": str(contact_distance),
            "fill": str(rdylgn_rgba(1.0 - v * 1.0 / max_contacts)),  # upstream writes the tuple itself
            "fill-opacity": str(0.2), "filter": "url(#blurMe)",
        }))
    tree.write(os.path.join(floor_directory, "heatmap.svg"))
