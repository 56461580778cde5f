"""Activation-range histograms: FixAccumulator (fixed range per activation type) and
RangeAccumulator (adaptive range). Parity: /root/reference/accumulator.py:51-231
(vectorised with numpy.histogram instead of per-element python loops)."""
from __future__ import annotations

import numpy

from ..core.memory import Array
from ..core.mutable import Bool
from ..core.units import Unit


class FixAccumulator(Unit):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **{k: v for k, v in kwargs.items()
                                      if k in ("name", "view_group")})
        self.bars = kwargs.get("bars", 200)
        self.type = kwargs.get("type", "relu")
        self.input = None
        self.output = Array()
        self.reset_flag = Bool(True)
        self.n_bars = [0]
        self.max = 100
        self.min = 0

    def initialize(self, **kwargs):
        self.output.reset(numpy.zeros(self.bars + 2, dtype=numpy.int64))

    def run(self):
        if self.type == "relu":
            self.max, self.min = 10000, 0
        elif self.type == "tanh":
            self.max, self.min = 1.7159, -1.7159
        else:
            raise ValueError("Unsupported type %s" % self.type)
        self.input.map_read()
        self.output.map_write()
        if bool(self.reset_flag):
            self.output.mem[:] = 0
        self.n_bars[0] = self.bars + 2
        y = self.input.mem.ravel()
        d = (self.bars - 1) / (self.max - self.min)
        self.output.mem[0] += int((y < self.min).sum())
        self.output.mem[self.bars + 1] += int((y > self.max).sum())
        inside = y[(y > self.min) & (y <= self.max)]
        idx = numpy.floor((inside - self.min) * d).astype(numpy.int64)
        numpy.add.at(self.output.mem, idx, 1)


class RangeAccumulator(Unit):
    """Histogram whose range grows with the data; ``reset_flag`` publishes x_out/y_out
    (optionally squashing empty bars) and starts over."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **{k: v for k, v in kwargs.items()
                                      if k in ("name", "view_group")})
        self.bars = kwargs.get("bars", 20)
        self.squash = kwargs.get("squash", True)
        self.input = None
        self.reset_flag = Bool(False)
        self.x, self.y, self.x_out, self.y_out = [], [], [], []
        self.gl_min, self.gl_max = None, None

    def initialize(self, **kwargs):
        pass

    @staticmethod
    def squash_bars(x, y):
        xo, yo = [], []
        for xv, yv in zip(x, y):
            if yv or not xo or yo[-1]:
                xo.append(xv)
                yo.append(yv)
        return xo, yo

    def run(self):
        if bool(self.reset_flag) and self.x:
            self.x_out, self.y_out = (self.squash_bars(self.x, self.y) if self.squash
                                      else (list(self.x), list(self.y)))
            self.x, self.y = [], []
            self.gl_min = self.gl_max = None
        self.input.map_read()
        data = self.input.mem.ravel()
        lo, hi = float(data.min()), float(data.max())
        if self.gl_min is None:
            self.gl_min, self.gl_max = lo, hi
            if hi == lo:
                hi = lo + 1.0
            edges = numpy.linspace(lo, hi, self.bars + 1)
            self.x = list((edges[:-1] + edges[1:]) * 0.5)
            self.y = list(numpy.histogram(data, bins=edges)[0])
            self._edges = edges
            return
        if lo < self.gl_min or hi > self.gl_max:
            # re-bin the accumulated histogram onto the widened range
            nlo, nhi = min(lo, self.gl_min), max(hi, self.gl_max)
            edges = numpy.linspace(nlo, nhi, self.bars + 1)
            old = numpy.histogram(numpy.array(self.x), bins=edges,
                                  weights=numpy.array(self.y, dtype=numpy.float64))[0]
            self.y = list(old.astype(numpy.int64))
            self.x = list((edges[:-1] + edges[1:]) * 0.5)
            self._edges = edges
            self.gl_min, self.gl_max = nlo, nhi
        h = numpy.histogram(data, bins=self._edges)[0]
        self.y = list(numpy.array(self.y) + h)
