"""Depooling: ``output[offset[i]] = input[i]`` using a pooling unit's ``input_offset``.

Parity: /root/reference/depooling.py:48-144 (memset + ``feed_layer`` kernel,
/root/reference/cuda/depooling.cu:4). The reference has no numpy path (:137-138).
"""
from __future__ import annotations

import numpy

from ..core.distributable import TriviallyDistributable
from ..core.memory import Array
from . import nn_units


class Depooling(nn_units.Forward, TriviallyDistributable):
    MAPPING = {"depooling"}

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.output_offset = None
        self.output_shape_source = None
        self.demand("input", "output_offset", "output_shape_source")

    def initialize(self, device=None, **kwargs):
        if not self.input or not self.output_offset or not self.output_shape_source:
            return True
        super().initialize(device=device, **kwargs)
        if self.output_offset.size != self.input.size:
            raise ValueError("output_offset.size must equal input.size")
        shape = tuple(self.output_shape_source.shape)
        self.make_output(shape, self.input.dtype)
        self.output.dev_dtype = self.input.dev_dtype
        self.init_vectors(self.input, self.output, self.output_offset)
        return None

    def numpy_run(self):
        self.input.map_read()
        self.output_offset.map_read()
        self.output.map_invalidate()
        flat = self.output.mem.reshape(-1)
        flat[:] = 0
        flat[self.output_offset.mem.ravel()] = self.input.mem.ravel()

    def cuda_run(self):
        from ..kernels import api
        api.depooling_forward(self)
