"""RProp: sign-based per-weight learning rates (host maths, like the reference).

Parity: /root/reference/rprop_gd.py:44-129. The reference discards the result of
``lrs * decrease_ratios`` (:87,113 — no decrease is ever applied); here the
decrease really happens. CPU-only by design (the reference has no device path either).
"""
from __future__ import annotations

import numpy

from ..core.memory import Array
from .gd import GradientDescent


class GDRProp(GradientDescent):
    MAPPING = {"rprop_gd"}

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.initial_learning_rate = kwargs.get("initial_learning_rate", 0.01)
        self.min_learning_rate = kwargs.get("min_learning_rate", 1e-6)
        self.max_learning_rate = kwargs.get("max_learning_rate", 1.0)
        self.increase = kwargs.get("increase", 1.05)
        self.decrease = kwargs.get("decrease", 0.80)
        self.weight_lrs = Array()
        self.bias_lrs = Array()
        self.force_numpy = True

    def initialize(self, device=None, **kwargs):
        r = super().initialize(device=device, **kwargs)
        if r:
            return r
        self.weight_lrs.reset(numpy.full(self.weights.shape, self.initial_learning_rate,
                                         dtype=self.weights.dtype))
        if self.include_bias and self.bias:
            self.bias_lrs.reset(numpy.full(self.bias.shape, self.initial_learning_rate,
                                           dtype=self.bias.dtype))
        return None

    def _step(self, vec, prev_grad, lrs, gradient):
        sign = numpy.sign(gradient)
        delta = numpy.sign(prev_grad.mem * gradient)
        lrs.mem *= numpy.where(delta > 0, self.increase, 1.0)
        lrs.mem *= numpy.where(delta < 0, self.decrease, 1.0)
        numpy.clip(lrs.mem, self.min_learning_rate, self.max_learning_rate, lrs.mem)
        if self.apply_gradient:
            vec.mem -= sign * lrs.mem
        prev_grad.mem[...] = gradient

    def numpy_weights_update(self):
        if not self.need_gradient_weights:
            return
        self.input.map_read()
        self.err_output.map_read()
        self.weights.map_write()
        self.gradient_weights.map_write()
        gradient = numpy.dot(self.err_output.matrix.transpose(), self.input.matrix)
        if self.weights_transposed:
            gradient = gradient.transpose()
        self._step(self.weights, self.gradient_weights, self.weight_lrs, gradient)

    def numpy_bias_update(self):
        if not self.need_gradient_weights or not self.include_bias:
            return
        self.err_output.map_read()
        self.bias.map_write()
        self.gradient_bias.map_write()
        self._step(self.bias, self.gradient_bias, self.bias_lrs,
                   self.err_output.matrix.sum(axis=0))
