def get_cmap(name=None, lut=None):
    def _cmap(v):
        v = min(1.0, max(0.0, float(v)))
        r = int(round(255 * (1.0 - v)))
        g = int(round(255 * v))
        return "#%02x%02x00" % (r, g)
    return _cmap


class ScalarMappable:
    def __init__(self, norm=None, cmap=None):
        self.cmap = cmap if callable(cmap) else get_cmap(cmap)
        self.norm = norm

    def to_rgba(self, v, *a, **k):
        return self.cmap(v)
