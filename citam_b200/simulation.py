"""Simulation — the reference's step-loop driver (citam/engine/core/simulation.py) on the GPU.

Same constructor, attributes and methods as upstream (`run_serial`, `create_agents`, `step`,
`move_agents`, `run_simulation_and_save_results`, `save_manifest`, `save_schedules`) and the same
output files.  Two ways through the loop:

* `step()` — one step at a time, exactly the upstream sequence: every agent takes its next
  itinerary entry on the host, the calculators' `run(current_step, active_agents)` plugin call
  goes to the device, the trajectory record is written.
* `run_simulation_and_save_results()` — when the only calculator is the device
  `CloseContactsCalculator`, the itineraries are loaded into HBM once and the T steps run in
  blocks on the device (`citam_run`); with `full_fidelity` the per-step files
  (`trajectory.txt`, `floor_k/contacts.txt`, `raw_contact_data.ccd`) are rebuilt block by block
  from the device's agent states and contact log, byte-identical to upstream.
"""
from __future__ import annotations

import hashlib
import json
import logging
import os
import uuid
from collections import OrderedDict
from typing import List, Optional, TextIO

import numpy as np

from . import itinerary as itin
from .agent import Agent
from .calculator import Calculator, CloseContactsCalculator
from .scheduler import Scheduler, SegmentScheduler

LOG = logging.getLogger(__name__)


class Simulation:
    def __init__(self, facility, total_timesteps: int, n_agents: int, calculators=None, scheduler=None, dry_run=False):
        """simulation.py:52-148 (same defaults and the same ValueErrors)."""
        self.facility = facility
        self.total_timesteps = total_timesteps
        self.n_agents = n_agents
        if calculators is None:
            LOG.info("No calculator provided. Defaulting to close contacts calculator.")
            self.calculators = [CloseContactsCalculator(facility)]
        elif isinstance(calculators, Calculator):
            self.calculators = [calculators]
        elif isinstance(calculators, list):
            for calc in calculators:
                if not isinstance(calc, Calculator):
                    raise ValueError("Invalid calculator. Must be an instance of Calculator.")
            self.calculators = calculators
        else:
            raise ValueError("Invalid calculator. Must be an instance of Calculator")
        if scheduler is None:
            # upstream defaults to its OfficeScheduler (simulation.py:133-141); itinerary generation is
            # not rebuilt here, so the default exists only where the upstream package is installed
            try:
                from citam.engine.schedulers.office_scheduler import OfficeScheduler
            except ImportError:
                raise ValueError(
                    "No scheduler provided. Itinerary generation (the reference's OfficeScheduler) is outside "
                    "this package: pass the reference's scheduler or a PrecomputedScheduler / SegmentScheduler."
                )
            LOG.info("No scheduler provided. Defaulting to office scheduler.")
            scheduler = OfficeScheduler(facility, total_timesteps)
        if not isinstance(scheduler, Scheduler) and not (hasattr(scheduler, "run") and hasattr(scheduler, "save_to_files")):
            raise ValueError("Invalid scheduler. Must be an instance of Scheduler")
        self.scheduler = scheduler
        self.current_step = 0
        self.agents = OrderedDict()
        self.dry_run = dry_run
        self.simulation_hash = None
        self.run_id = str(uuid.uuid4())
        self.block_steps = 4096  # device steps per citam_run call in the block path
        self._lut = None

    # -- setup -----------------------------------------------------------------------------------
    def create_sim_hash(self) -> None:
        """simulation.py:163-213: a digest of the inputs (upstream's is not reproducible either:
        it hashes object reprs)."""
        m = hashlib.blake2b(digest_size=10)
        m.update(repr([self.total_timesteps, self.n_agents, getattr(self.facility, "entrances", None),
                       getattr(self.facility, "facility_name", None), self.dry_run]).encode("utf-8"))
        self.simulation_hash = m.hexdigest()

    def create_agents(self) -> None:
        """simulation.py:277-289: agent ids are 0..n-1 in schedule order."""
        schedules = self.scheduler.run(self.n_agents)
        for current_agent, schedule in enumerate(schedules):
            agent = Agent(current_agent, schedule)
            self.agents[agent.unique_id] = agent

    def run_serial(self, workdir, sim_name: str, run_name: str) -> None:
        """simulation.py:215-246 ."""
        self.create_sim_hash()
        self.save_manifest(workdir, sim_name, run_name)
        self.save_maps(workdir)
        self.create_agents()
        self.save_schedules(workdir)
        for calculator in self.calculators:
            calculator.initialize(self.agents, workdir)
        LOG.info("Running simulation with %d agents", self.n_agents)
        self.run_simulation_and_save_results(workdir)

    # -- the loop ----------------------------------------------------------------------------------
    def _device_calculator(self) -> Optional[CloseContactsCalculator]:
        if self.dry_run or len(self.calculators) != 1 or not isinstance(self.calculators[0], CloseContactsCalculator):
            return None
        for a in self.agents.values():
            if not hasattr(a.schedule, "itinerary"):
                return None
        return self.calculators[0]

    def run_simulation_and_save_results(self, workdir) -> None:
        """simulation.py:248-275."""
        calc = self._device_calculator()
        if calc is None:
            with open(os.path.join(workdir, "trajectory.txt"), "w") as t_outfile:
                for i in range(self.total_timesteps):
                    self.step(traj_outfile=t_outfile)
        else:
            self._run_blocks(calc, workdir)
        LOG.info("Done with all %d steps.", self.current_step)
        for c in self.calculators:
            c.finalize(self.agents, work_directory=workdir)

    def _run_blocks(self, calc: CloseContactsCalculator, workdir) -> None:
        eng = calc.engine
        if isinstance(self.scheduler, SegmentScheduler):
            eng.load_segments(self.scheduler.seg_off, self.scheduler.segs)
        else:
            eng.load_itineraries([a.schedule.itinerary for a in self.agents.values()])
        T = self.total_timesteps
        full = calc.full_fidelity and workdir is not None
        traj = open(os.path.join(workdir, "trajectory.txt"), "w") if full else None
        block = self.block_steps if not full else min(self.block_steps, max(1, (1 << 22) // max(1, self.n_agents)))
        try:
            for t0 in range(0, T, block):
                t1 = min(T, t0 + block)
                calc.run_range(t0, t1)
                if full:
                    self._write_block(calc, t0, t1, traj)
                self.current_step = t1
        finally:
            if traj is not None:
                traj.close()
        # leave the agents where the loop ends, like upstream's objects
        st = eng.states_at(T - 1) if T > 0 else None
        if st is not None:
            for a, s in zip(self.agents.values(), st):
                if s["floor"] >= 0:
                    a.pos = (int(s["x"]), int(s["y"]))
                    a.current_floor = int(s["floor"])
                    a.current_location = None if s["loc"] < 0 else int(s["loc"])

    def _write_block(self, calc, t0: int, t1: int, traj: TextIO) -> None:
        """Per-step records of steps [t0, t1): trajectory.txt (simulation.py:312-334) and
        floor_k/contacts.txt (close_contacts_calculator.py:126-137), from the device's agent states
        and contact log."""
        eng = calc.engine
        states = eng.states_range(t0, t1)
        log = eng.contact_log()  # sorted by (step, a1, a2): the reference's visiting order
        eng.clear_contact_log()
        steps = log["step"]
        lo = np.searchsorted(steps, np.arange(t0, t1), side="left")
        hi = np.searchsorted(steps, np.arange(t0, t1), side="right")
        n_floors = self.facility.number_of_floors
        dur = calc._durations
        for k, t in enumerate(range(t0, t1)):
            st = states[k]
            active = np.flatnonzero(st["floor"] >= 0)
            calc.step_contact_locations = [{}] * n_floors
            n_located = int((st["loc"][active] >= 0).sum())
            if len(active) > 1 and n_located > 0:
                calc._apply_step_log(log[lo[k] : hi[k]], st, None, t)
                calc._write_contacts_record(t)
            traj.write(str(len(active)) + "\nstep: " + str(t) + "\n")
            x, y, fl = st["x"], st["y"], st["floor"]
            for a in active.tolist():
                traj.write("%d\t%d\t%d\t%d\t%d\n" % (a, x[a], y[a], fl[a], dur[a]))

    def step(self, traj_outfile: TextIO = None) -> None:
        """simulation.py:291-336: one step through the plugin interface."""
        active_agents, _ = self.move_agents()
        if not self.dry_run:
            for calculator in self.calculators:
                calculator.run(self.current_step, active_agents)
        if traj_outfile is not None:
            traj_outfile.write(str(len(active_agents)) + "\nstep: " + str(self.current_step) + "\n")
            for agent in active_agents:
                contact_count = agent.cumulative_properties.get("contact_duration", 0)
                traj_outfile.write(
                    str(agent.unique_id) + "\t" + str(agent.pos[0]) + "\t" + str(agent.pos[1]) + "\t"
                    + str(agent.current_floor) + "\t" + str(contact_count) + "\n"
                )
        self.current_step += 1

    def _locate(self, floor, x, y):
        """Space index from the device-built location table (fetched once per floor)."""
        if self._lut is None:
            self._lut = {}
        if floor not in self._lut:
            eng = self._engine_for_lookup()
            minx, miny, W, H, _ = eng.floor_meta[floor]
            self._lut[floor] = (minx, miny, eng.floor_lut(floor))
        minx, miny, lut = self._lut[floor]
        fx, fy = x - minx, y - miny
        if fx < 0 or fy < 0 or fy >= lut.shape[0] or fx >= lut.shape[1]:
            return None
        v = int(lut[fy, fx])
        return None if v < 0 else v

    def _engine_for_lookup(self):
        for c in self.calculators:
            if isinstance(c, CloseContactsCalculator):
                return c.engine
        raise RuntimeError("locating agents needs an initialised CloseContactsCalculator (it owns the location table)")

    def move_agents(self):
        """simulation.py:338-361."""
        active_agents, moving_agents = [], []
        for unique_id, agent in self.agents.items():
            if agent.step(self._locate):
                moving_agents.append(agent)
            if agent.pos is not None:
                active_agents.append(agent)
        return active_agents, moving_agents

    # -- files ---------------------------------------------------------------------------------------
    def save_manifest(self, work_directory, sim_name: str, run_name: str) -> None:
        """simulation.py:363-420: the keys the dashboard reads."""
        floors = [{"name": str(f), "directory": "floor_" + str(f) + "/"} for f in range(self.facility.number_of_floors)]
        tp = getattr(self.facility, "traffic_policy", None)
        n_one_way = 0 if tp is None else len([p for p in tp if p["direction"] != 0])
        fp0 = self.facility.floorplans[0]
        fp_width = fp0.maxx - fp0.minx
        manifest = {
            "TimestepInSec": 1,
            "NumberOfFloors": self.facility.number_of_floors,
            "NumberOfOneWayAisles": n_one_way,
            "NumberOfAgents": self.n_agents,
            "SimulationName": sim_name,
            "RunName": run_name,
            "SimulationHash": self.simulation_hash,
            "RunID": self.run_id,
            "FacilityName": self.facility.facility_name,
            "MaxRoomOccupancy": 1.0,
            "NumberOfShifts": 1,
            "NumberOfEntrances": 1,
            "NumberOfExits": 1,
            "EntranceScreening": False,
            "TrajectoryFile": "trajectory.txt",
            "Floors": floors,
            "ScaleMultiplier": max(1, round(fp_width / 1500.0)),
            "Timestep": 1,
            "TotalTimesteps": self.total_timesteps,
        }
        with open(os.path.join(work_directory, "manifest.json"), "w") as f:
            json.dump(manifest, f, indent=4)

    def save_maps(self, work_directory) -> None:
        """simulation.py:422-458: floor_k/map.svg -- the floorplan's walls inside `<g id="root">` with
        the upstream viewBox rule (citam_b200/svg.py writes what `export_world_to_svg` would)."""
        from . import svg

        geom = None
        for f in range(self.facility.number_of_floors):
            d = os.path.join(work_directory, "floor_" + str(f))
            os.makedirs(d, exist_ok=True)
            fp = self.facility.floorplans[f]
            if getattr(fp, "walls", None) is not None:
                walls = svg.wall_paths_of_floorplan(fp)
            else:
                if geom is None:
                    from .facility import geometry_from_facility

                    geom = geometry_from_facility(self.facility)
                walls = svg.floor_walls(geom["floors"][f])
            svg.write_map_svg(walls, os.path.join(d, "map.svg"), [fp.minx - 50, fp.miny - 50, fp.maxx + 50, fp.maxy + 50])

    def save_schedules(self, work_directory) -> None:
        """simulation.py:460-486."""
        self.scheduler.save_to_files(work_directory)
        if self.n_agents > 100_000:
            return  # one text block per agent is a dashboard nicety, not for 10^6 agents
        with open(os.path.join(work_directory, "schedules.txt"), "w") as outfile:
            for unique_id, agent in self.agents.items():
                outfile.write("Agent ID: " + str(unique_id) + str(agent.schedule) + "\n\n")
