"""Calculator plugin API and the close-contacts calculator on the GPU.

`Calculator` is the reference's plugin ABC (citam/engine/calculators/calculator.py:9-33).
`CloseContactsCalculator` keeps the reference's constructor, attributes and the
initialize / run / finalize protocol (citam/engine/calculators/close_contacts_calculator.py),
and computes on the device through the C ABI: `run(current_step, agents)` is the per-step plugin
call (host agent states in, one `citam_step_states`), `run_range(t0, t1)` is the block call
`Simulation` uses when the itineraries are device-resident.  There is no CPU fallback.
"""
from __future__ import annotations

import json
import os
from abc import ABC, abstractmethod
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np

from . import _lib as L
from .contacts import ContactEvents
from .engine import DeviceEngine


class Calculator(ABC):
    """calculator.py:9-33."""

    def __init__(self, facility, **kwargs) -> None:
        super().__init__()
        self.facility = facility

    @abstractmethod
    def initialize(self, agents, work_directory=None) -> None:
        ...

    @abstractmethod
    def run(self, current_step: int, agents: List, **kwargs) -> None:
        ...

    @abstractmethod
    def finalize(self, agents, work_directory=None) -> None:
        ...


class CloseContactsCalculator(Calculator):
    def __init__(
        self,
        facility,
        contact_distance: float = 6.0,
        *,
        device: int = 0,
        full_fidelity: bool = True,
        log_capacity: int = 1 << 24,
        pair_capacity: int = 0,
        coord_capacity: int = 0,
    ) -> None:
        """close_contacts_calculator.py:21-38.  `full_fidelity` keeps the per-contact-step log
        the per-step outputs need (contacts.txt, trajectory contact counts, raw_contact_data.ccd);
        switch it off for large runs that only need the end-of-run tables."""
        super().__init__(facility)
        self.contact_distance = contact_distance
        self.contact_events = ContactEvents()
        self.step_contact_locations: List[Dict] = []
        self.full_fidelity = bool(full_fidelity)
        self._opts = dict(device=device, pair_capacity=pair_capacity, coord_capacity=coord_capacity,
                          log_capacity=log_capacity if full_fidelity else 0)
        self._engine: Optional[DeviceEngine] = None
        self._agents = None
        self._files = []
        self._durations = None
        self.heatmap_max_points = 200_000  # one svg circle per coordinate: skipped beyond this

    # -- protocol -----------------------------------------------------------------------------
    @property
    def engine(self) -> DeviceEngine:
        if self._engine is None:
            raise RuntimeError("CloseContactsCalculator.initialize() has not been called")
        return self._engine

    def initialize(self, agents, work_directory=None) -> None:
        """close_contacts_calculator.py:40-62."""
        for agent in agents.values():
            if "contact_duration" in agent.cumulative_properties:
                raise ValueError("Property clash detected. Agents already have a 'contact_duration' property.")
            agent.cumulative_properties["contact_duration"] = 0
        self._agents = agents
        n = len(agents)
        if self._engine is not None:
            self._engine.close()
        self._engine = DeviceEngine(n, self.facility.number_of_floors, self.contact_distance, **self._opts)
        self._engine.load_geometry(self.facility)
        self.contact_events.bind(self._engine)
        self._durations = np.zeros(n, np.int64)
        self._index = {a.unique_id: i for i, a in enumerate(agents.values())}
        self._files = []
        if work_directory is not None:
            for f in range(self.facility.number_of_floors):
                d = os.path.join(work_directory, "floor_" + str(f))
                os.makedirs(d, exist_ok=True)
                self._files.append(open(os.path.join(d, "contacts.txt"), "w"))

    def run(self, current_step: int, agents: List, **kwargs) -> None:
        """close_contacts_calculator.py:64-137 for one step; `agents` are the active agents."""
        n_floors = self.facility.number_of_floors
        self.step_contact_locations = [{}] * n_floors  # one dict shared by all floors, as upstream (:86-88)
        if len(agents) <= 1:
            return
        located = [a for a in agents if a.current_location is not None]
        if len(located) == 0:
            return
        eng = self.engine
        states = np.zeros(eng.n_agents, L.AGENT_STATE)
        states["floor"] = -1
        states["loc"] = -1
        for a in agents:
            i = self._index[a.unique_id]
            states["x"][i], states["y"][i] = a.pos[0], a.pos[1]
            states["floor"][i] = a.current_floor
            states["loc"][i] = -1 if a.current_location is None else a.current_location
        eng.step_states(current_step, states)
        if eng.log_enabled:
            log = eng.contact_log()  # this step's contacts, in the reference's visiting order
            eng.clear_contact_log()
            self._apply_step_log(log, states, list(self._agents.values()), current_step)
            self._write_contacts_record(current_step)

    def run_range(self, t_begin: int, t_end: int) -> None:
        """Steps [t_begin, t_end) from the device-resident itineraries (one C-ABI call)."""
        self.engine.run(t_begin, t_end)

    def finalize(self, agents, work_directory=None) -> None:
        """close_contacts_calculator.py:139-150."""
        for f in self._files:
            f.close()
        self._files = []
        dur = self.engine.agent_durations()
        for a, d in zip(agents.values(), dur.tolist()):
            a.cumulative_properties["contact_duration"] = int(d)
        if work_directory is not None:
            self.save_to_files(agents, work_directory)

    # -- per-step bookkeeping (full fidelity) ---------------------------------------------------
    def _apply_step_log(self, log, states, agent_list, step=None) -> None:
        """add_contact_event (close_contacts_calculator.py:205-231) for one step's contacts."""
        locs = self.step_contact_locations[0] if self.step_contact_locations else {}
        x, y = states["x"].tolist(), states["y"].tolist()
        pairs = list(zip(log["a1"].tolist(), log["a2"].tolist()))
        for a1, a2 in pairs:
            pos = (x[a1] + (x[a2] - x[a1]) / 2.0, y[a1] + (y[a2] - y[a1]) / 2.0)
            locs[pos] = locs.get(pos, 0) + 1
            self._durations[a1] += 1
            self._durations[a2] += 1
            if agent_list is not None:
                agent_list[a1].cumulative_properties["contact_duration"] += 1
                agent_list[a2].cumulative_properties["contact_duration"] += 1
        if step is not None:
            loc = [None if v < 0 else v for v in states["loc"].tolist()]
            self.contact_events.feed_step(step, pairs, x, y, states["floor"].tolist(), loc)

    def _write_contacts_record(self, current_step: int) -> None:
        for fi, fh in enumerate(self._files):
            locs = self.step_contact_locations[fi]
            fh.write(str(len(locs)) + "\nstep :" + str(current_step) + "\n")
            for cl, nc in locs.items():
                fh.write(str(cl[0]) + "\t" + str(cl[1]) + "\t" + str(nc) + "\n")

    # -- outputs -----------------------------------------------------------------------------------
    def extract_contact_distribution_per_agent(self, agents):
        """close_contacts_calculator.py:233-248."""
        ids, n = [], []
        for unique_id, agent in agents.items():
            ids.append("ID_" + str(unique_id + 1))
            n.append(agent.cumulative_properties["contact_duration"])
        return ids, n

    def save_to_files(self, agents, work_directory: str) -> None:
        """close_contacts_calculator.py:250-318, including floor_k/heatmap.svg when floor_k/map.svg
        exists (citam_b200/svg.py)."""
        ids, n = self.extract_contact_distribution_per_agent(agents)
        with open(os.path.join(work_directory, "contact_dist_per_agent.csv"), "w") as f:
            f.write("agent_ID,Number_of_Contacts\n")
            for i, c in zip(ids, n):
                f.write(i + "," + str(c) + "\n")
        self.contact_events.save_pairwise_contacts(os.path.join(work_directory, "pair_contact.csv"))
        if self.engine.log_enabled:
            self.contact_events.save_raw_contact_data(os.path.join(work_directory, "raw_contact_data.ccd"))
        with open(os.path.join(work_directory, "statistics.json"), "w") as f:
            json.dump({"data": self.contact_events.extract_statistics()}, f)
        for fn in range(self.facility.number_of_floors):
            d = os.path.join(work_directory, "floor_" + str(fn))
            os.makedirs(d, exist_ok=True)
            per_coord = self.contact_events.get_contacts_per_coordinates(fn)
            with open(os.path.join(d, "contact_dist_per_coord.csv"), "a") as f:  # append mode, as upstream (:309)
                f.write("x,y,n_contacts\n")
                for (x, y), c in per_coord.items():
                    f.write(str(x) + "," + str(y) + "," + str(c) + "\n")
            if os.path.isfile(os.path.join(d, "map.svg")) and len(per_coord) <= self.heatmap_max_points:
                from . import svg

                svg.write_heatmap_svg(per_coord, d, self.contact_distance)
