"""CPU: host-side logic (itinerary store, facility tables, synthetic generator, C-ABI surface)."""
import ctypes
import os
import re

import numpy as np
import pytest

from golden_util import CASES, load_case

from citam_b200 import _lib as L
from citam_b200 import facility as fac
from citam_b200 import itinerary as itin
from citam_b200 import synthetic as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", CASES)
def test_segment_encoding_is_lossless(name):
    c = load_case(name)
    seg_off, segs = itin.encode(c["itineraries"])
    assert seg_off[0] == 0 and seg_off[-1] == len(segs)
    rng = np.random.RandomState(1)
    for a, it in enumerate(c["itineraries"]):
        t0 = segs["t0"][seg_off[a] : seg_off[a + 1]]
        assert (np.diff(t0) > 0).all()
        ts = set(rng.randint(0, c["T"] + 50, 40).tolist()) | {0, len(it) - 1, len(it), c["T"] - 1}
        for t in ts:
            k = min(t, len(it) - 1)
            while k >= 0 and it[k][0] is None:  # None is sticky (agent.py:77-107)
                k -= 1
            want = None if k < 0 else (it[k][0][0], it[k][0][1], it[k][1])
            assert itin.decode(seg_off, segs, a, t) == want, (a, t)
    # far fewer segments than entries: the store is a run-length code
    assert len(segs) < sum(len(it) for it in c["itineraries"]) / 5


def test_pack_rejects_non_integer_and_floorless_entries():
    with pytest.raises(ValueError):
        itin.pack([[((1.5, 2), 0)]])
    with pytest.raises(ValueError):
        itin.pack([[((1, 2), None)]])


def test_slice_segments_reproduces_states():
    f, seg_off, segs, p = S.make_config("tiny")
    so2, sg2 = itin.slice_segments(seg_off, segs, 699, 900)
    assert len(sg2) < len(segs) / 3
    for a in range(0, p["n_agents"], 7):
        for t in (699, 700, 777, 899):
            assert itin.decode(seg_off, segs, a, t) == itin.decode(so2, sg2, a, t)


def test_record_sizes_match_header():
    hdr = open(os.path.join(ROOT, "include", "citam_b200.h")).read()
    assert L.SEGMENT.itemsize == 20 and L.AGENT_STATE.itemsize == 12 and L.SPACE.itemsize == 32
    assert L.BOUNDARY.itemsize == 56 and L.DOOR.itemsize == 32 and L.PAIR.itemsize == 24
    assert L.COORD.itemsize == 24 and L.CONTACT.itemsize == 12
    assert ctypes.sizeof(L.Config) == 56 and ctypes.sizeof(L.Counters) == 80 and ctypes.sizeof(L.Stats) == 48
    assert "#define CITAM_ABI_VERSION %d" % L.ABI_VERSION in hdr


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports exactly what include/citam_b200.h declares."""
    hdr = open(os.path.join(ROOT, "include", "citam_b200.h")).read()
    declared = set(re.findall(r"^(?:int|void|const char\*)\s+(citam_\w+)\(", hdr, re.M))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.citam_abi_version() == L.ABI_VERSION == 3


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from citam_b200.engine import DeviceEngine

    with pytest.raises(L.CitamDeviceError):
        DeviceEngine(4, 1, 6.0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "citam_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "grid_oracle" not in src, fn


def test_floor_tables_follow_the_reference_conventions():
    c = load_case("twofloor")
    fg = c["geometry"]["floors"][0]
    minx, miny, W, H, spaces, bnd, doors, adj = fac.floor_tables(fg)
    assert (minx, miny) == (0, -120)  # negative coordinates: lattice origin comes from the spaces
    assert len(spaces) == len(fg["spaces"]) and spaces["seg_end"][-1] == len(bnd)
    # hallways never take the door shortcut (space.py:218-224)
    for i, sp in enumerate(fg["spaces"]):
        if fac.is_hallway(sp["function"]):
            assert spaces["door_begin"][i] == spaces["door_end"][i]
    assert fac.is_hallway("Circulation (main)") and fac.is_hallway("LOBBY") and not fac.is_hallway("office")
    n = np.hypot(bnd["nx"], bnd["ny"])
    assert np.allclose(n[bnd["long_enough"] == 1], 1.0)


def test_synthetic_generator_is_deterministic_and_well_formed():
    f1, so1, sg1, p = S.make_config("tiny", seed=5)
    f2, so2, sg2, _ = S.make_config("tiny", seed=5)
    assert (so1 == so2).all() and (sg1 == sg2).all()
    _, so3, sg3, _ = S.make_config("tiny", seed=6)
    assert len(sg3) != len(sg1) or (sg3 != sg1).any()
    # visited points lie in a space, except on the way out to the parking point below the floor
    minx, miny, lut = f1.lut_expected(0)
    for a in range(0, p["n_agents"], 3):
        for t in range(0, p["total_timesteps"], 37):
            d = itin.decode(so1, sg1, a, t)
            if d is None or d[1] < 0:
                continue
            x, y, fl = d
            assert 0 <= fl < p["n_floors"]
            assert lut[y - miny, x - minx] >= 0, (a, t, d)
    # the reference-format expansion agrees with the segment store
    its = S.expand_to_reference_format(so1, sg1, p["total_timesteps"])
    so4, sg4 = itin.encode(its)
    for a in range(0, p["n_agents"], 11):
        for t in range(0, p["total_timesteps"], 53):
            assert itin.decode(so1, sg1, a, t) == itin.decode(so4, sg4, a, t)


def test_statistics_from_sums_matches_python_rounding():
    from citam_b200.engine import statistics_from_sums

    s = dict(total_duration=46467, n_pairs=190, n_agents_with_contacts=20, sum_events_per_agent=436, max_events_per_agent=25)
    v = [d["value"] for d in statistics_from_sums(s)]
    assert v == [774.45, 21.8, 38.72, 20, 9.5, 25]  # SURVEY.md Appendix B.5 (reference, seed 0)
    z = dict(total_duration=0, n_pairs=0, n_agents_with_contacts=0, sum_events_per_agent=0, max_events_per_agent=0)
    assert [d["value"] for d in statistics_from_sums(z)] == [0.0, 0, 0.0, 0, 0, 0]


def test_svg_writers(tmp_path):
    """Structure of map.svg / heatmap.svg as visualization.py:298-473 and
    close_contacts_calculator.py:320-377 produce them."""
    import xml.etree.ElementTree as ET

    from citam_b200 import svg

    c = load_case("twofloor")
    walls = svg.floor_walls(c["geometry"]["floors"][0])
    assert len(walls) > 10 and len(set(walls)) == len(walls) and all(w.startswith("M ") and " L " in w for w in walls)
    d = str(tmp_path)
    with pytest.raises(FileNotFoundError):
        svg.write_heatmap_svg({(1.0, 2.0): 3}, d, 6.0)
    svg.write_map_svg(walls, os.path.join(d, "map.svg"), [-50, -170, 400, 250])
    root = ET.parse(os.path.join(d, "map.svg")).getroot()
    assert root.attrib == {"baseProfile": "full", "height": "100%", "version": "1.1", "viewBox": "-50 -170 400 250", "width": "100%"}
    g = root[0]
    assert len(root) == 1 and g.attrib == {"id": "root"} and g[0].tag.endswith("defs") and len(g) == 1 + len(walls)
    # stroke width 3.0 * round(xmax / 1500, 1) (visualization.py:340-344, 358-361)
    assert g[1].attrib == {"d": walls[0], "fill": "white", "stroke": "black", "stroke-width": str(3.0 * round(400 / 1500, 1))}
    svg.write_heatmap_svg({(1.0, 2.0): 3, (5.5, 7.0): 1}, d, 6.0)
    g = ET.parse(os.path.join(d, "heatmap.svg")).getroot()[0]
    circles = [e for e in g if e.tag.endswith("circle")]
    assert len(circles) == 2 and circles[0].attrib["filter"] == "url(#blurMe)"
    assert circles[0].attrib == {"cx": "1.0", "cy": "2.0", "r": "6.0", "fill": str(svg.rdylgn_rgba(0.0)), "fill-opacity": "0.2",
                                 "filter": "url(#blurMe)"}
    # upstream hands matplotlib's RGBA tuple to ElementTree as it is: the attribute is the tuple's string
    assert circles[0].attrib["fill"] == "(0.6470588235294118, 0.0, 0.14901960784313725, 1.0)"
    assert svg.rdylgn_hex(0.0) == "#a50026" and svg.rdylgn_hex(1.0) == "#006837" and svg.rdylgn_hex(0.5) == "#feffbe"
    f = [e for e in g if e.tag.endswith("filter")]
    assert len(f) == 1 and f[0].attrib == {"id": "blurMe", "width": "12.0", "height": "12.0", "x": "-100%", "y": "-100%"}
    assert f[0][0].attrib == {"in": "SourceGraphic", "stdDeviation": str(6.0 * 0.4), "edgeMode": "duplicate"}


def test_svg_files_against_the_references_own_example_files(tmp_path):
    """examples/basic_example/floor_0/{map,heatmap}.svg were written by the real svgpathtools / svgwrite /
    matplotlib stack: our map of the same facility is byte-identical up to the stroke width (the example
    predates the viewBox-scaled width), and every heatmap circle gets the same fill string."""
    import csv
    import re

    from oracle import refload

    ex = os.path.join(refload.REFERENCE_ROOT, "examples", "basic_example")
    if not refload.available() or not os.path.isfile(os.path.join(ex, "floor_0", "map.svg")):
        pytest.skip("reference tree not present (GPU box)")
    import subprocess
    import sys

    from citam_b200 import svg

    script = (
        "import sys, os\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import refload\n"
        "refload.load()\n"
        "os.chdir(%r)\n"
        "from citam.engine.main import load_floorplans\n"
        "from citam_b200 import svg\n"
        "fp = load_floorplans('foo_facility', ['foo_floor'])[0]\n"
        "svg.write_map_svg(svg.wall_paths_of_floorplan(fp), %r, [fp.minx - 50, fp.miny - 50, fp.maxx + 50, fp.maxy + 50])\n"
        % (ROOT, ex, str(tmp_path / "map.svg"))
    )
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ours = open(str(tmp_path / "map.svg")).read()
    theirs = open(os.path.join(ex, "floor_0", "map.svg")).read()
    assert 'stroke-width="%s"' % (3.0 * round(1250.0 / 1500, 1)) in ours
    assert ours.replace('stroke-width="%s"' % (3.0 * round(1250.0 / 1500, 1)), 'stroke-width="3.0"') == theirs
    # heatmap colours
    txt = open(os.path.join(ex, "floor_0", "heatmap.svg")).read()
    circ = re.findall(r'<circle cx="([^"]+)" cy="([^"]+)" r="[^"]+" fill="([^"]+)"', txt)
    rows = list(csv.reader(open(os.path.join(ex, "floor_0", "contact_dist_per_coord.csv"))))[1:]
    cnt = {(float(a), float(b)): int(c) for a, b, c in rows}
    assert len(circ) == len(cnt) > 100
    mx = max(cnt.values())
    for cx, cy, fill in circ:
        assert str(svg.rdylgn_rgba(1.0 - cnt[(float(cx), float(cy))] * 1.0 / mx)) == fill
