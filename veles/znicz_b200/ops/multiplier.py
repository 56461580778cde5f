"""Pointwise product of two arrays and its gradient (LSTM glue).

Parity: /root/reference/multiplier.py (Multiplier :47, GDMultiplier :112).
"""
from __future__ import annotations

import numpy

from ..core.accelerated_units import AcceleratedUnit
from ..core.memory import Array


class _BinaryBase(AcceleratedUnit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.output = Array()
        self.demand("x", "y")

    def initialize(self, device=None, **kwargs):
        src = self.x if self.x else self.y
        if src and (not self.output or self.output.shape != src.shape):
            self.output.reset(numpy.zeros_like(src.mem))
            self.output.dev_dtype = src.dev_dtype
        if not self.x or not self.y:
            return True
        super().initialize(device=device, **kwargs)
        if not (self.output.shape == self.x.shape == self.y.shape):
            raise ValueError("%s: shapes differ: %s %s" % (self, self.x.shape,
                                                          self.y.shape))
        self.init_vectors(self.x, self.y, self.output)
        return None


class Multiplier(_BinaryBase):
    """output = x * y."""

    def numpy_run(self):
        self.x.map_read()
        self.y.map_read()
        self.output.map_invalidate()
        numpy.multiply(self.x.mem, self.y.mem, self.output.mem)

    def cuda_run(self):
        from ..kernels import api
        api.binary_forward(self, "mul")


class GDMultiplier(AcceleratedUnit):
    """err_x = err_output * y; err_y = err_output * x."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.err_x = Array()
        self.err_y = Array()
        self.demand("x", "y", "err_output")

    def initialize(self, device=None, **kwargs):
        if not self.x or not self.y or not self.err_output:
            return True
        super().initialize(device=device, **kwargs)
        for arr, ref in ((self.err_x, self.x), (self.err_y, self.y)):
            if not arr or arr.shape != ref.shape:
                arr.reset(numpy.zeros_like(ref.mem))
                arr.dev_dtype = ref.dev_dtype
        self.init_vectors(self.err_x, self.err_y, self.x, self.y, self.err_output)
        return None

    def numpy_run(self):
        self.x.map_read()
        self.y.map_read()
        self.err_output.map_read()
        self.err_x.map_invalidate()
        self.err_y.map_invalidate()
        eo = self.err_output.mem.reshape(self.x.shape)
        numpy.multiply(eo, self.y.mem, self.err_x.mem)
        numpy.multiply(eo, self.x.mem, self.err_y.mem)

    def cuda_run(self):
        from ..kernels import api
        api.multiplier_backward(self)
