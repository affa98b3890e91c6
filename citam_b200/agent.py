"""Agent — same attributes as the reference's (citam/engine/core/agent.py:25-65)."""
from __future__ import annotations


class Agent:
    def __init__(self, unique_id, schedule):
        self.unique_id = unique_id
        self.schedule = schedule
        self.pos = None
        self.current_location = None
        self.current_floor = None
        self.cumulative_properties = {}

    def step(self, locate=None) -> bool:
        """agent.py:67-109: take the next itinerary entry; a `None` entry changes nothing.
        Returns upstream's `has_moved`: the entry differs from the current position or floor (so a
        `None` entry counts as a move once the agent is placed, a repeated position does not).
        `locate(floor, x, y)` gives the space index (the device-built location table)."""
        xy, floor = self.schedule.get_next_position()
        has_moved = xy != self.pos or floor != self.current_floor
        if xy is not None:
            self.pos = xy
            self.current_floor = floor
            if locate is not None:
                self.current_location = locate(floor, xy[0], xy[1])
        return has_moved
