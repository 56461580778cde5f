"""ResizableAll2All: grow/shrink the neuron count after initialisation, preserving the
existing weights (/root/reference/resizable_all2all.py:41-80)."""
from __future__ import annotations

import numpy

from .all2all import All2All


class ResizableAll2All(All2All):
    MAPPING = {"all2all_resizable"}

    @property
    def output_sample_shape(self):
        return self._output_sample_shape

    @output_sample_shape.setter
    def output_sample_shape(self, value):
        ready = self.is_initialized or bool(self.__dict__.get("weights"))
        old = self.neurons_number if ready else 0
        self._set_output_sample_shape(value)
        if not ready:
            return
        if self.neurons_number <= 0:
            raise ValueError("Neurons number must be greater than 0 (got %s)" % (value,))
        self._adjust_neurons_number(self.neurons_number - old)

    def _adjust_neurons_number(self, delta):
        if delta == 0:
            return
        self.weights.map_read()
        if not self.weights_transposed:
            old_nn = self.weights.shape[0]
            new_w = numpy.zeros((old_nn + delta, self.weights.shape[1]), self.weights.dtype)
            if delta > 0:
                new_w[:old_nn] = self.weights.mem
                self.fill_array(self.weights_filling, new_w[old_nn:], self.weights_stddev)
            else:
                new_w[:] = self.weights.mem[:new_w.shape[0]]
        else:
            old_nn = self.weights.shape[1]
            new_w = numpy.zeros((self.weights.shape[0], old_nn + delta), self.weights.dtype)
            if delta > 0:
                new_w[:, :old_nn] = self.weights.mem
                self.fill_array(self.weights_filling, new_w[:, old_nn:], self.weights_stddev)
            else:
                new_w[:] = self.weights.mem[:, :new_w.shape[1]]
        self.weights.reset(new_w)
        if self.include_bias and self.bias:
            self.bias.map_read()
            new_b = numpy.zeros(old_nn + delta, self.bias.dtype)
            n = min(old_nn, old_nn + delta)
            new_b[:n] = self.bias.mem[:n]
            if delta > 0:
                self.fill_array(self.bias_filling, new_b[old_nn:], self.bias_stddev)
            self.bias.reset(new_b)
        self.output.reset()
        self.make_output(self.output_shape, self.input.dtype)
        self.weights_shape = (self.neurons_number, self.input.sample_size)
        self.init_vectors(self.weights, self.bias, self.output)
        self.weights_lp_ = None
        self.weights_lp_t_ = None
        if self.on_cuda:
            self.refresh_shadows()
