"""Stub of appdirs (TEST INFRASTRUCTURE ONLY)."""
import os
import tempfile


def user_data_dir(appname=None, appauthor=None, *a, **k):
    return os.path.join(tempfile.gettempdir(), "citam_shim_cache", appname or "citam")
