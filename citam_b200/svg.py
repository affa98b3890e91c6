"""map.svg / heatmap.svg for the dashboard.

The reference draws `floor_k/map.svg` with svgpathtools (`export_world_to_svg`,
citam/engine/io/visualization.py:298-443, then `add_root_layer_to_svg` :446-473 wraps everything in
`<g id="root">`) and `floor_k/heatmap.svg` by appending a blur filter and one circle per contact
coordinate, coloured by matplotlib's RdYlGn map (close_contacts_calculator.py:320-377).  Neither
library is a dependency here: the walls are written as plain `<path d="M x,y L x,y">` elements and the
colour map is the 11-class ColorBrewer RdYlGn table, linearly interpolated.  The files have the
structure the dashboard serves (root group, same viewBox rule, same circle attributes); they are not
byte-compared with the reference's, whose text depends on svgpathtools/matplotlib versions.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from typing import Dict, Iterable, Sequence, Tuple

_RDYLGN = [
    (165, 0, 38), (215, 48, 39), (244, 109, 67), (253, 174, 97), (254, 224, 139), (255, 255, 191),
    (217, 239, 139), (166, 217, 106), (102, 189, 99), (26, 152, 80), (0, 104, 55),
]
_NS = "http://www.w3.org/2000/svg"


def rdylgn_hex(v: float) -> str:
    """RdYlGn at v in [0, 1] (0 = red, 1 = green) as #rrggbb."""
    v = min(1.0, max(0.0, float(v)))
    p = v * (len(_RDYLGN) - 1)
    i = min(int(p), len(_RDYLGN) - 2)
    f = p - i
    a, b = _RDYLGN[i], _RDYLGN[i + 1]
    return "#%02x%02x%02x" % tuple(int(round(a[k] + (b[k] - a[k]) * f)) for k in range(3))


def write_map_svg(walls: Iterable[Sequence[float]], svg_file: str, viewbox: Sequence[float]) -> None:
    """Walls `[sx, sy, ex, ey]` as black paths inside `<g id="root">`; viewBox as upstream
    (`minx-50 miny-50 maxx+50 maxy+50`, simulation.py:452-457)."""
    ET.register_namespace("", _NS)
    svg = ET.Element("{%s}svg" % _NS, {"width": "100%", "height": "100%",
                                       "viewBox": " ".join(str(v) for v in viewbox), "version": "1.1"})
    root = ET.SubElement(svg, "{%s}g" % _NS, {"id": "root"})
    for sx, sy, ex, ey in walls:
        ET.SubElement(root, "{%s}path" % _NS, {"d": "M %s,%s L %s,%s" % (sx, sy, ex, ey), "fill": "none",
                                               "stroke": "black", "stroke-width": "2"})
    ET.ElementTree(svg).write(svg_file)


def floor_walls(floor_geometry: Dict) -> list:
    """Every boundary segment of every space of a geometry floor, once."""
    seen, out = set(), []
    for sp in floor_geometry["spaces"]:
        for seg in sp["boundaries"]:
            k = tuple(seg)
            kr = (seg[2], seg[3], seg[0], seg[1])
            if k not in seen and kr not in seen:
                seen.add(k)
                out.append(seg)
    return out


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
            "cx": str(x), "cy": str(y), "r": str(contact_distance), "fill": rdylgn_hex(1.0 - v * 1.0 / max_contacts),
            "fill-opacity": str(0.2), "filter": "url(#blurMe)",
        }))
    tree.write(os.path.join(floor_directory, "heatmap.svg"))
