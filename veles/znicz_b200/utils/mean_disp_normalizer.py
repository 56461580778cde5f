"""MeanDispNormalizer unit: output = (input − mean) · rdisp on the device
(``veles.mean_disp_normalizer``; linked by StandardWorkflow.link_meandispnorm,
/root/reference/standard_workflow.py:603-624)."""
from __future__ import annotations

import numpy

from ..core.accelerated_units import AcceleratedUnit
from ..core.memory import Array


class MeanDispNormalizer(AcceleratedUnit):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.output = Array()
        self.demand("input", "mean", "rdisp")

    def initialize(self, device=None, **kwargs):
        if not self.input:
            return True
        super().initialize(device=device, **kwargs)
        if not self.output or self.output.shape != self.input.shape:
            self.output.reset(numpy.zeros(self.input.shape, numpy.float32))
        self.init_vectors(self.input, self.output)
        return None

    def _coeffs(self):
        mean = self.mean.mem if isinstance(self.mean, Array) else numpy.asarray(self.mean)
        rdisp = self.rdisp.mem if isinstance(self.rdisp, Array) else numpy.asarray(self.rdisp)
        return mean.astype(numpy.float32), rdisp.astype(numpy.float32)

    def numpy_run(self):
        self.input.map_read()
        self.output.map_invalidate()
        mean, rdisp = self._coeffs()
        self.output.mem[...] = (self.input.mem.astype(numpy.float32) - mean) * rdisp

    def cuda_run(self):
        import torch
        mean, rdisp = self._coeffs()
        key = "coef_"
        co = self.__dict__.get(key)
        if co is None:
            dev = self.device.torch_device
            co = (torch.from_numpy(numpy.ascontiguousarray(numpy.broadcast_to(
                      mean, self.input.shape[1:]))).to(dev),
                  torch.from_numpy(numpy.ascontiguousarray(numpy.broadcast_to(
                      rdisp, self.input.shape[1:]))).to(dev))
            self.__dict__[key] = co
        x = self.input.dev.float()
        self.output.dev_out.copy_((x - co[0]) * co[1])
