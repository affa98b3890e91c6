"""Shared helpers: run the CPU oracles on a case (test infrastructure)."""
import os
import tempfile

import numpy as np

from citam_b200 import facility as fac
from citam_b200 import itinerary as itin
from oracle import grid, port


def grid_oracle_for(geometry, D, seg_off, segs, threads=0):
    return grid.GridOracle.from_geometry(geometry, D, seg_off, segs, fac.floor_tables, threads)


def run_port(case_or_itins, geometry, D, T, workdir=None):
    sim = port.PortSimulation(case_or_itins, geometry, D)
    sim.run(T, workdir)
    return sim


def tables_of_port(sim, n_floors):
    return port.summary_tables(sim.events.contact_data, [a.dur for a in sim.agents], n_floors)


def tables_of(obj):
    """(pairs list, coords dict, durations list) from anything with pairs()/coords()/agent_durations()."""
    p = [(int(r["a1"]), int(r["a2"]), int(r["first_step"]), int(r["n_events"]), int(r["duration"])) for r in obj.pairs()]
    c = {(int(r["floor"]), int(r["sx"]), int(r["sy"])): int(r["count"]) for r in obj.coords()}
    return p, c, [int(v) for v in obj.agent_durations()]
