"""Straight-line subset of the svgpathtools API (TEST INFRASTRUCTURE ONLY).

The upstream CITAM reference imports svgpathtools, which is not installed in
this image and cannot be fetched.  CITAM floorplans on the simulation path hold
only M/L/Z commands, so `Line`, `Path`, `parse_path` and a dummy `wsvg` are all
the reference needs to run here.  Arithmetic follows svgpathtools' published
definitions (Line.point/length/unit_tangent/normal/bbox, Path.point/bbox/d).
This shim exists only so `oracle/refload.py` can import `/root/reference`; it is
never imported by the product package.
"""
import re


class Line:
    def __init__(self, start, end):
        self.start = complex(start)
        self.end = complex(end)

    def __repr__(self):
        return "Line(start=%s, end=%s)" % (self.start, self.end)

    def __eq__(self, other):
        if not isinstance(other, Line):
            return NotImplemented
        return self.start == other.start and self.end == other.end

    def __ne__(self, other):
        if not isinstance(other, Line):
            return NotImplemented
        return not self == other

    def __hash__(self):
        return hash((self.start, self.end))

    def __getitem__(self, item):
        return self.bpoints()[item]

    def __len__(self):
        return 2

    def bpoints(self):
        return self.start, self.end

    def point(self, t):
        return self.start + (self.end - self.start) * t

    def length(self, t0=0, t1=1, error=None, min_depth=None):
        return abs(self.end - self.start) * (t1 - t0)

    def derivative(self, t=None, n=1):
        if n == 1:
            return self.end - self.start
        return 0

    def unit_tangent(self, t=None):
        dseg = self.end - self.start
        return dseg / abs(dseg)

    def normal(self, t=None):
        return -1j * self.unit_tangent(t)

    def bbox(self):
        xmin = min(self.start.real, self.end.real)
        xmax = max(self.start.real, self.end.real)
        ymin = min(self.start.imag, self.end.imag)
        ymax = max(self.start.imag, self.end.imag)
        return xmin, xmax, ymin, ymax

    def reversed(self):
        return Line(self.end, self.start)

    def translated(self, z0):
        return Line(self.start + z0, self.end + z0)

    def rotated(self, degs, origin=None):
        import cmath, math

        if origin is None:
            origin = self.point(0.5)
        r = cmath.exp(1j * math.radians(degs))
        return Line(origin + r * (self.start - origin), origin + r * (self.end - origin))

    def scaled(self, sx, sy=None, origin=0j):
        if sy is None:
            sy = sx

        def f(z):
            z = z - origin
            return complex(sx * z.real, sy * z.imag) + origin

        return Line(f(self.start), f(self.end))


class CubicBezier:  # placeholder: never instantiated on the simulation path
    def __init__(self, *a, **k):
        raise NotImplementedError("shim: straight-line subset only")


class Arc(CubicBezier):
    pass


class QuadraticBezier(CubicBezier):
    pass


class Path(list):
    """A list of segments with the handful of Path methods CITAM calls."""

    def __init__(self, *segments):
        super().__init__(segments)

    def __repr__(self):
        return "Path(%s)" % (",\n     ".join(repr(x) for x in self))

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        if not isinstance(other, Path):
            return NotImplemented
        return len(self) == len(other) and all(a == b for a, b in zip(self, other))

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    @property
    def start(self):
        return self[0].start

    @property
    def end(self):
        return self[-1].end

    def bbox(self):
        bbs = [seg.bbox() for seg in self]
        xmins, xmaxs, ymins, ymaxs = list(zip(*bbs))
        return min(xmins), max(xmaxs), min(ymins), max(ymaxs)

    def length(self, T0=0, T1=1, error=None, min_depth=None):
        return sum(seg.length() for seg in self) * (T1 - T0)

    def point(self, pos):
        if pos == 0.0:
            return self[0].point(pos)
        if pos == 1.0:
            return self[-1].point(pos)
        lengths = [seg.length() for seg in self]
        total = sum(lengths)
        fractions = [x / total for x in lengths]
        segment_start = 0
        for index, segment in enumerate(self):
            segment_end = segment_start + fractions[index]
            if segment_end >= pos:
                segment_pos = (pos - segment_start) / (segment_end - segment_start)
                return segment.point(segment_pos)
            segment_start = segment_end
        return self[-1].point(1.0)

    def d(self, useSandT=False, use_closed_attrib=False, rel=False):
        parts = []
        current = None
        for seg in self:
            if current is None or seg.start != current:
                parts.append("M {},{}".format(seg.start.real, seg.start.imag))
            parts.append("L {},{}".format(seg.end.real, seg.end.imag))
            current = seg.end
        return " ".join(parts)

    def isclosed(self):
        return len(self) > 0 and self[0].start == self[-1].end

    def reversed(self):
        return Path(*[seg.reversed() for seg in reversed(self)])

    def translated(self, z0):
        return Path(*[seg.translated(z0) for seg in self])

    def rotated(self, degs, origin=None):
        if origin is None:
            origin = self.point(0.5)
        return Path(*[seg.rotated(degs, origin) for seg in self])

    def scaled(self, sx, sy=None, origin=0j):
        return Path(*[seg.scaled(sx, sy, origin) for seg in self])


_TOKEN = re.compile(r"([MmLlHhVvZz])|([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)")


def parse_path(pathdef, current_pos=0j, tree_element=None):
    tokens = [(m.group(1), m.group(2)) for m in _TOKEN.finditer(pathdef)]
    segs = Path()
    i = 0
    cmd = None
    start_pos = None
    pos = current_pos

    def num():
        nonlocal i
        v = float(tokens[i][1])
        i += 1
        return v

    while i < len(tokens):
        if tokens[i][0] is not None:
            cmd = tokens[i][0]
            i += 1
            if cmd in "Zz":
                if start_pos is not None and pos != start_pos:
                    segs.append(Line(pos, start_pos))
                pos = start_pos
                continue
        if cmd is None:
            raise ValueError("path data must start with a command")
        rel = cmd.islower()
        c = cmd.upper()
        if c == "M":
            x, y = num(), num()
            pos = (pos if rel else 0j) + complex(x, y)
            start_pos = pos
            cmd = "l" if rel else "L"
        elif c == "L":
            x, y = num(), num()
            new = (pos if rel else 0j) + complex(x, y)
            segs.append(Line(pos, new))
            pos = new
        elif c == "H":
            x = num()
            new = complex((pos.real if rel else 0) + x, pos.imag)
            segs.append(Line(pos, new))
            pos = new
        elif c == "V":
            y = num()
            new = complex(pos.real, (pos.imag if rel else 0) + y)
            segs.append(Line(pos, new))
            pos = new
        else:
            raise NotImplementedError("shim: command %r" % cmd)
    return segs


def svg2paths(*a, **k):
    raise NotImplementedError("shim: svg2paths is not on the simulation path")


def wsvg(paths=None, colors=None, filename=None, stroke_widths=None, nodes=None,
         node_colors=None, node_radii=None, openinbrowser=False, timestamp=False,
         margin_size=0.1, mindim=600, dimensions=None, viewbox=None, text=None,
         text_path=None, font_size=None, attributes=None, svg_attributes=None,
         svgwrite_debug=False, paths2Drawing=False, baseunit="px"):
    """Write a minimal but valid SVG holding the given paths (map.svg is not parity-relevant)."""
    paths = paths or []
    vb = viewbox if viewbox is not None else "0 0 100 100"
    if not isinstance(vb, str):
        vb = " ".join(str(v) for v in vb)
    out = ['<?xml version="1.0" encoding="utf-8" ?>',
           '<svg xmlns="http://www.w3.org/2000/svg" xmlns:xlink="http://www.w3.org/1999/xlink" '
           'baseProfile="full" version="1.1" viewBox="%s">' % vb, "<defs />"]
    for k, p in enumerate(paths):
        attrs = dict(attributes[k]) if attributes else {}
        d = p.d() if hasattr(p, "d") else Path(p).d()
        attr_s = " ".join('%s="%s"' % (a, v) for a, v in attrs.items())
        out.append('<path d="%s" %s/>' % (d, attr_s))
    out.append("</svg>")
    if paths2Drawing:
        class _D:
            def __init__(self, s):
                self._s = s
            def tostring(self):
                return self._s
        return _D("\n".join(out))
    with open(filename, "w") as f:
        f.write("\n".join(out))
