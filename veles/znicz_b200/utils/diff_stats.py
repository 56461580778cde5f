"""DiffStats: per-step sum-abs-delta statistics of selected arrays, pickled at stop.
Parity: /root/reference/diff_stats.py:48-129."""
from __future__ import annotations

import pickle
from collections import defaultdict

import numpy

from ..core.units import Unit


class DiffStats(Unit):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "PLOTTER")
        self._arrays = {u: set(v) for u, v in kwargs.get("arrays", {}).items()}
        super().__init__(workflow, **kwargs)
        self._stats = {u: defaultdict(list) for u in self._arrays}
        self._file_name = kwargs.get("file_name")

    def init_unpickled(self):
        super().init_unpickled()
        self._previous_ = {}

    stats = property(lambda self: self._stats)

    @property
    def file_name(self):
        return self._file_name

    @file_name.setter
    def file_name(self, value):
        if not isinstance(value, str):
            raise TypeError("file_name must be a string")
        self._file_name = value

    @property
    def size(self):
        return sum(sum(len(a) for a in v.values()) for v in self._stats.values())

    def initialize(self, **kwargs):
        pass

    def register(self, unit, attr):
        self._arrays.setdefault(unit, set()).add(attr)
        self._stats.setdefault(unit, defaultdict(list))

    def unregister(self, unit, attr):
        self._arrays[unit].remove(attr)

    def run(self):
        for unit, anames in self._arrays.items():
            prev = self._previous_.setdefault(id(unit), {})
            for aname in anames:
                vector = getattr(unit, aname, None)
                if vector is None or not vector:
                    continue
                vector.map_read()
                array = vector.mem
                if aname not in prev:
                    prev[aname] = array.copy()
                    continue
                delta = float(numpy.sum(numpy.abs(array - prev[aname])))
                self._stats[unit][aname].append(
                    {"delta": delta, "abs": float(numpy.sum(numpy.abs(array)))})
                prev[aname][...] = array

    def stop(self):
        super().stop()
        if not self._file_name or not self.size:
            return
        with open(self._file_name, "wb") as fout:
            pickle.dump({u.name: dict(vals) for u, vals in self._stats.items()}, fout,
                        protocol=pickle.HIGHEST_PROTOCOL)

    def generate_data_for_master(self):
        out = {}
        for u, vals in self._stats.items():
            out[u.name] = {attr: st[-1] for attr, st in vals.items() if st}
        return out
