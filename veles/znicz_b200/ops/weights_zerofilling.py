"""ZeroFiller: block mask on the next layer's weights each step (AlexNet "grouping").

Parity: /root/reference/weights_zerofilling.py:46-137. The mask keeps
``kernel % g != chan % g`` exactly as the reference code computes it (:95-98);
``grouping`` must be >= 2 (the reference's default of 1 is rejected by its own
setter — SURVEY §9 — so the default here is 2). The device kernel is bounds-checked
(the reference launches grid = W.size, block = 1 without a check).
"""
from __future__ import annotations

import numpy

from ..core.distributable import TriviallyDistributable
from ..core.memory import Array
from .nn_units import ForwardBase


class ZeroFiller(ForwardBase, TriviallyDistributable):
    MAPPING = {"zero_filter"}

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.mask = Array()
        self.grouping = kwargs.get("grouping", 2)
        self.demand("weights")

    @property
    def effective_shape(self):
        return (self.weights.shape[0], self.weights.size // self.weights.shape[0])

    @property
    def grouping(self):
        return self._grouping

    @grouping.setter
    def grouping(self, value):
        if not isinstance(value, int):
            raise TypeError("grouping value must be an integer (got %s)" % type(value))
        if value < 2:
            raise ValueError("grouping value %d is invalid" % value)
        self._grouping = value

    def initialize(self, device=None, **kwargs):
        super().initialize(device=device, **kwargs)
        if not self.weights:
            return True
        if not self.mask:
            if self.effective_shape[1] % self.grouping != 0:
                raise ValueError("Non-multiple of grouping weights shape detected: "
                                 "%s, grouping=%d" % (self.weights.shape, self.grouping))
            k = numpy.arange(self.effective_shape[0])[:, None] % self.grouping
            c = numpy.arange(self.effective_shape[1])[None, :] % self.grouping
            self.mask.reset((k != c).astype(self.weights.dtype))
        elif self.mask.shape != self.effective_shape:
            raise ValueError("mask shape mismatch")
        self.init_vectors(self.mask, self.weights)
        return None

    def numpy_run(self):
        self.mask.map_read()
        self.weights.map_write()
        self.weights.mem.reshape(self.effective_shape)[...] *= self.mask.mem

    def cuda_run(self):
        from ..kernels import api
        api.zero_filler(self)
