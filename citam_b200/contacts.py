"""ContactEvent / ContactEvents over the device tables.

Mirror of citam/engine/calculators/contacts.py: the same container API
(`contact_data`, `count`, `save_pairwise_contacts`, `extract_statistics`,
`get_floor_contact_coords`, `get_contacts_per_coordinates`, `save_raw_contact_data`), but the
numbers live on the GPU: the pair table, the per-coordinate histogram and (full-fidelity mode)
the contact log.  `contact_data` — the reference's dict of per-pair event lists with per-step
positions — is materialised on demand from the contact log.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .engine import statistics_from_sums


class ContactEvent:
    """contacts.py:40-70."""

    def __init__(self, floor_number, location, position, current_step):
        self.locations = [location]
        self.floor_numbers = [floor_number]
        self.positions = [position]
        self.start_step = current_step
        self.duration = 1


class ContactEvents:
    def __init__(self, engine=None):
        self._engine = engine
        self._data: Dict[str, List[ContactEvent]] = {}

    def bind(self, engine) -> None:
        self._engine = engine
        self._data = {}

    # -- tables (cheap) ---------------------------------------------------------------------
    def pair_table(self) -> np.ndarray:
        """Rows (a1, a2, first_step, n_events, duration) in first-contact order."""
        return self._engine.pairs()

    def count(self) -> int:
        """contacts.py:144-153: total number of contact events."""
        return int(self.pair_table()["n_events"].sum())

    def save_pairwise_contacts(self, filename: str) -> None:
        """contacts.py:155-183."""
        p = self.pair_table()
        with open(filename, "w") as f:
            f.write("Agent1,Agent2,N_Contacts,TotalContactDuration\n")
            for a1, a2, n, d in zip(p["a1"].tolist(), p["a2"].tolist(), p["n_events"].tolist(), p["duration"].tolist()):
                f.write("%d,%d,%d,%d\n" % (a1, a2, n, d))

    def extract_statistics(self) -> List[Dict]:
        """contacts.py:185-294 (integer sums on the device, Python's division and round())."""
        return statistics_from_sums(self._engine.statistics_sums())

    def get_contacts_per_coordinates(self, floor_number: int) -> Dict[Tuple[float, float], int]:
        """contacts.py:321-346.  With the contact log the dict has the reference's own insertion
        order; without it the same mapping ordered by coordinate."""
        if self._engine.log_enabled:
            out: Dict[Tuple[float, float], int] = {}
            for key in self.contact_data:
                fpos = self.get_floor_contact_coords(key, floor_number)
                for pos in list(set(fpos)):
                    out[pos] = out.get(pos, 0) + fpos.count(pos)
            return out
        co = self._engine.coords()
        co = co[co["floor"] == floor_number]
        return {(sx / 2.0, sy / 2.0): int(c) for sx, sy, c in zip(co["sx"].tolist(), co["sy"].tolist(), co["count"].tolist())}

    # -- full fidelity (contact log) --------------------------------------------------------
    def feed_step(self, t: int, pairs, x, y, fl, loc) -> None:
        """Apply one step's contacts, given in the reference's visiting order ((a1, a2) ascending),
        with that step's agent states: `ContactEvents.add_contact` (contacts.py:92-142)."""
        data = self._data
        for a1, a2 in pairs:
            pos = (x[a1] + (x[a2] - x[a1]) / 2.0, y[a1] + (y[a2] - y[a1]) / 2.0)
            k = "%d-%d" % (a1, a2)
            evs = data.get(k)
            if evs is not None and evs[-1].start_step + evs[-1].duration == t:
                last = evs[-1]
                last.duration += 1
                last.positions.append(pos)
                last.locations.append(loc[a1])
                last.floor_numbers.append(fl[a1])
            elif evs is not None:
                evs.append(ContactEvent(fl[a1], loc[a1], pos, t))
            else:
                data[k] = [ContactEvent(fl[a1], loc[a1], pos, t)]

    @property
    def contact_data(self) -> Dict[str, List[ContactEvent]]:
        """The reference's dict of per-pair event lists, built from the device's contact log as the
        calculator / simulation exports it step block by step block."""
        if self._engine is not None and not self._engine.log_enabled:
            raise RuntimeError(
                "contact_data (per-step positions of every event) needs the contact log: create the "
                "calculator with full_fidelity=True; the aggregate tables are available without it"
            )
        return self._data

    def get_floor_contact_coords(self, key: str, floor_number: int):
        """contacts.py:296-319."""
        out = []
        for ce in self.contact_data[key]:
            for i, pos in enumerate(ce.positions):
                if ce.floor_numbers[i] == floor_number:
                    out.append(pos)
        return out

    def save_raw_contact_data(self, filename: str) -> None:
        """contacts.py:348-368."""
        data = {k: [v.__dict__ for v in evs] for k, evs in self.contact_data.items()}
        with open(filename, "w") as f:
            f.write(str(data))
