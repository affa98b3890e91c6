def to_hex(c, keep_alpha=False):
    return c if isinstance(c, str) else "#000000"


class Normalize:
    def __init__(self, vmin=None, vmax=None, clip=False):
        self.vmin, self.vmax = vmin, vmax

    def __call__(self, v):
        if self.vmax in (None, self.vmin):
            return 0.0
        return (v - self.vmin) / (self.vmax - self.vmin)
