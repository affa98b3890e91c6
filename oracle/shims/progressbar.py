"""Stub of progressbar2 (TEST INFRASTRUCTURE ONLY; see oracle/refload.py)."""


class ProgressBar:
    def __init__(self, *a, **k):
        pass

    def update(self, *a, **k):
        pass

    def start(self, *a, **k):
        return self

    def finish(self, *a, **k):
        pass


def progressbar(iterable, *a, **k):
    return iterable
