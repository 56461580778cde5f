"""InputJoiner: concatenates several [batch, n_i] inputs along the feature axis
(``veles.input_joiner.InputJoiner``; used by the LSTM cell, /root/reference/lstm.py:75-108).
Exposes ``offset_i`` / ``length_i`` for every input so backward units can slice."""
from __future__ import annotations

import numpy

from .accelerated_units import AcceleratedUnit
from .memory import Array


class InputJoiner(AcceleratedUnit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.output = Array()
        self._sources = []          # (unit, attr)
        self.offsets = []
        self.lengths = []

    def link_inputs(self, other, *attrs):
        for a in attrs:
            i = len(self._sources)
            self._sources.append((other, a))
            self.link_attrs(other, ("input_%d" % i, a))
        return self

    @property
    def inputs(self):
        return [getattr(self, "input_%d" % i) for i in range(len(self._sources))]

    def __getattr__(self, name):
        # offset_i / length_i resolve lazily (available after initialize)
        if name.startswith("offset_") or name.startswith("length_"):
            kind, idx = name.split("_")
            lst = self.__dict__.get("offsets" if kind == "offset" else "lengths", [])
            i = int(idx)
            if i < len(lst):
                return lst[i]
            return None
        raise AttributeError(name)

    def initialize(self, device=None, **kwargs):
        ins = self.inputs
        if any(a is None or not a for a in ins):
            return True
        super().initialize(device=device, **kwargs)
        batch = ins[0].shape[0]
        self.lengths = [a.size // a.shape[0] for a in ins]
        self.offsets = [int(sum(self.lengths[:i])) for i in range(len(ins))]
        total = int(sum(self.lengths))
        if not self.output or self.output.shape != (batch, total):
            self.output.reset(numpy.zeros((batch, total), dtype=ins[0].dtype))
            self.output.dev_dtype = ins[0].dev_dtype
        self.init_vectors(self.output, *ins)
        return None

    def numpy_run(self):
        self.output.map_invalidate()
        for a, off, n in zip(self.inputs, self.offsets, self.lengths):
            a.map_read()
            self.output.mem[:, off:off + n] = a.matrix

    def cuda_run(self):
        ext = self.ext_
        out = self.output.dev
        for a, off, n in zip(self.inputs, self.offsets, self.lengths):
            src = a.dev
            ext.axpby_2d(src.view(src.shape[0], -1), 0, out, off, n, 1.0, 0.0)
        self.output.dev_written()
