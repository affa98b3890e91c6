"""Schedule / Scheduler interfaces of the hot path's input side.

The reference's `Scheduler` ABC (citam/engine/schedulers/scheduler.py:5-23: `run(n_agents)` ->
list of schedules, `save_to_files(workdir)`) and the de-facto `Schedule` contract the step loop
uses (citam/engine/schedulers/office_schedule.py:632-638 `get_next_position`, `.itinerary`).
Itinerary *generation* (office scheduler, A* routes) is upstream of the path and not rebuilt here:
any object with an `.itinerary` list in the reference's format — including the reference's own
`OfficeSchedule` — can be handed to `Simulation`.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Sequence


class Schedule:
    """An agent's itinerary: `[((x, y) | None, floor | None), ...]`, one entry per time step."""

    def __init__(self, itinerary: Sequence, navigation=None):
        self.itinerary = itinerary
        self.navigation = navigation
        self.current_standing = 0

    def get_next_position(self):
        """office_schedule.py:632-638: entry at the cursor; the cursor stops on the last entry."""
        position = self.itinerary[self.current_standing]
        if self.current_standing < len(self.itinerary) - 1:
            self.current_standing += 1
        return position

    def __str__(self) -> str:
        return " precomputed itinerary of %d steps" % len(self.itinerary)


class Scheduler(ABC):
    """scheduler.py:5-23."""

    def __init__(self, facility, timestep=1, total_timesteps=0):
        self.facility = facility
        self.timestep = timestep
        self.total_timesteps = total_timesteps

    @abstractmethod
    def run(self, n_agents: int) -> List[Schedule]:
        ...

    @abstractmethod
    def save_to_files(self, work_directory: str) -> None:
        ...


class PrecomputedScheduler(Scheduler):
    """Hands out given itineraries (reference-format lists) in order."""

    def __init__(self, facility, total_timesteps: int, itineraries: Sequence[Sequence]):
        super().__init__(facility, 1, total_timesteps)
        self.itineraries = itineraries

    def run(self, n_agents: int) -> List[Schedule]:
        nav = getattr(self.facility, "navigation", None)
        return [Schedule(it, nav) for it in self.itineraries[:n_agents]]

    def save_to_files(self, work_directory: str) -> None:
        pass


class SegmentScheduler(Scheduler):
    """Itineraries already in the device format (seg_off, segments): no per-step lists exist.

    `run` returns lightweight schedules that only carry their agent index; `Simulation` loads the
    segment store straight into the engine (the path for 10^5-10^6 agents)."""

    def __init__(self, facility, total_timesteps: int, seg_off, segs):
        super().__init__(facility, 1, total_timesteps)
        self.seg_off = seg_off
        self.segs = segs

    def run(self, n_agents: int) -> List[Schedule]:
        if n_agents != len(self.seg_off) - 1:
            raise ValueError("segment store holds %d agents, %d requested" % (len(self.seg_off) - 1, n_agents))
        return [_SegmentSchedule(a) for a in range(n_agents)]

    def save_to_files(self, work_directory: str) -> None:
        pass


class _SegmentSchedule(Schedule):
    def __init__(self, agent_index: int):
        super().__init__((), None)
        self.agent_index = agent_index

    def get_next_position(self):
        raise RuntimeError("segment-backed schedules are stepped on the device, not on the host")

    def __str__(self) -> str:
        return " device-resident itinerary"
