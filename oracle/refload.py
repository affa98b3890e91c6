"""Import the UNMODIFIED upstream CITAM reference from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module is used (a) by `oracle/make_golden.py`
to generate the committed fixtures under `tests/golden/`, and (b) by CPU tests
that cross-check the restatements in `oracle/port.py` / `oracle/grid_oracle.c`
against the real reference when `/root/reference` is present (it is not on the
GPU box, so those tests skip there).  The product package `citam_b200` never
imports it.

What it does:
  1. puts `oracle/shims` (svgpathtools straight-line subset, progressbar,
     matplotlib, svgwrite, appdirs stubs) on sys.path — none of those packages
     is installed and there is no network;
  2. pre-seeds `sys.modules['citam']` with an empty package whose __path__ is
     the reference tree, so `citam/__init__.py` (which imports the falcon/boto3
     API server) is not executed;
  3. patches networkx>=3.6 `node_link_graph` back to the `links` key the
     shipped hallway/route JSON uses.
"""
import functools
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CITAM_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "citam", "engine"))


def load():
    """Make `import citam.engine...` resolve to the reference; idempotent."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if getattr(load, "_done", False):
        return sys.modules["citam"]
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    pkg = types.ModuleType("citam")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "citam")]
    sys.modules["citam"] = pkg

    import networkx.readwrite.json_graph as jg
    import networkx.readwrite.json_graph.node_link as nl

    if not getattr(nl.node_link_graph, "_citam_patched", False):
        patched = functools.partial(nl.node_link_graph, edges="links")
        patched._citam_patched = True
        nl.node_link_graph = patched
        jg.node_link_graph = patched
    import networkx as nx

    nx.node_link_graph = jg.node_link_graph
    load._done = True
    return pkg
