"""Pointwise sum of two arrays and its gradient (LSTM glue).

Parity: /root/reference/summator.py (Summator :47, GDSummator :112).
"""
from __future__ import annotations

import numpy

from ..core.accelerated_units import AcceleratedUnit
from ..core.memory import Array
from .multiplier import _BinaryBase


class Summator(_BinaryBase):
    """output = x + y."""

    def numpy_run(self):
        self.x.map_read()
        self.y.map_read()
        self.output.map_invalidate()
        numpy.add(self.x.mem, self.y.mem, self.output.mem)

    def cuda_run(self):
        from ..kernels import api
        api.binary_forward(self, "add")


class GDSummator(AcceleratedUnit):
    """err_x = err_y = err_output."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.err_x = Array()
        self.err_y = Array()
        self.demand("err_output")

    def initialize(self, device=None, **kwargs):
        if not self.err_output:
            return True
        super().initialize(device=device, **kwargs)
        for arr in (self.err_x, self.err_y):
            if not arr or arr.shape != self.err_output.shape:
                arr.reset(numpy.zeros_like(self.err_output.mem))
                arr.dev_dtype = self.err_output.dev_dtype
        self.init_vectors(self.err_x, self.err_y, self.err_output)
        return None

    def numpy_run(self):
        self.err_output.map_read()
        self.err_x.map_invalidate()
        self.err_y.map_invalidate()
        self.err_x.mem[...] = self.err_output.mem
        self.err_y.mem[...] = self.err_output.mem

    def cuda_run(self):
        eo = self.err_output.dev
        self.err_x.unmap()
        self.err_y.unmap()
        self.err_x.devmem.copy_(eo)
        self.err_y.devmem.copy_(eo)
        self.err_x.dev_written()
        self.err_y.dev_written()
