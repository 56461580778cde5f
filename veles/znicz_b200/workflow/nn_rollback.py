"""NNRollback: divergence recovery by learning-rate back-off + weight restore.

Parity: /root/reference/nn_rollback.py:44-190 — on train-improved: lr×1.04 and stash
weights; otherwise after ``minus_steps``: lr×0.65 and restore; NaN check. The
reference's restore is a no-op (``setattr(gd, "weights.mem[:]", …)`` creates a junk
attribute); here weights really are rolled back (host copy → device).
"""
from __future__ import annotations

import numpy

from ..core.units import Unit


class NNRollback(Unit):
    weights_names = ("weights", "bias", "gradient_weights", "gradient_bias")

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.lr_plus = kwargs.get("lr_plus", 1.04)
        self.lr_minus = kwargs.get("lr_minus", 0.65)
        self.plus_steps = kwargs.get("plus_steps", 1)
        self.minus_steps = kwargs.get("minus_steps", 3)
        self._plus_steps = self.plus_steps
        self._minus_steps = self.minus_steps
        self.improved = None
        self.demand("improved")
        self._gds = []
        self.history_limit = 2
        self._first_run = True

    def init_unpickled(self):
        super().init_unpickled()
        self.slaves_ = {}

    def initialize(self, **kwargs):
        self.info("lr_plus=%.2f lr_minus=%.2f", self.lr_plus, self.lr_minus)

    # -- elastic hooks (IDistributable) -------------------------------------------------
    def generate_data_for_slave(self, slave=None):
        self.slaves_[getattr(slave, "id", slave)] = 1

    def generate_data_for_master(self):
        return True

    def apply_data_from_master(self, data):
        pass

    def apply_data_from_slave(self, data, slave=None):
        self._slave_ended(slave)

    def _slave_ended(self, slave):
        self.slaves_.pop(getattr(slave, "id", slave), None)
        if not self.slaves_ and not bool(self.gate_skip) and not bool(self.gate_block):
            self.run()

    def drop_slave(self, slave=None):
        self._slave_ended(slave)

    # -- bookkeeping ------------------------------------------------------------------------
    def add_gd(self, gd, lr_plus=None, lr_minus=None):
        for kv in self._gds:
            if kv["gd"] is gd:
                kv["lr_plus"], kv["lr_minus"] = lr_plus, lr_minus
                return
        self._gds.append({"gd": gd, "lr_plus": lr_plus, "lr_minus": lr_minus,
                          "history": {}})

    def reset(self):
        del self._gds[:]

    def _stash(self, kv):
        gd = kv["gd"]
        for name in self.weights_names:
            arr = getattr(gd, name, None)
            if arr is None or not arr:
                continue
            arr.map_read()
            hist = kv["history"].setdefault(name, [])
            hist.append(arr.mem.copy())
            while len(hist) > self.history_limit:
                hist.pop(0)

    def _restore(self, kv, rollback_to=0):
        gd = kv["gd"]
        for name in self.weights_names:
            arr = getattr(gd, name, None)
            hist = kv["history"].get(name)
            if arr is None or not arr:
                continue
            if not hist:
                self.warning("No rollback for %s of %s", name, gd)
                continue
            arr.map_invalidate()
            arr.mem[...] = hist[rollback_to]
            del hist[rollback_to + 1:]
            arr.unmap()
        fu = getattr(gd, "forward_unit", None)
        if fu is not None and getattr(fu, "on_cuda", False):
            fu.refresh_shadows()

    def _has_nans(self, kv):
        gd = kv["gd"]
        for name in self.weights_names:
            arr = getattr(gd, name, None)
            if arr is None or not arr:
                continue
            arr.map_read()
            if not numpy.isfinite(arr.mem).all():
                return True
        return False

    def run(self):
        if bool(self.improved):
            self._plus_steps += 1
            if self._plus_steps < self.plus_steps:
                self._first_run = False
                return
            self._plus_steps = 0
            self._minus_steps = 0
            for kv in self._gds:
                k = kv["lr_plus"] if kv["lr_plus"] is not None else self.lr_plus
                gd = kv["gd"]
                gd.learning_rate *= k
                gd.learning_rate_bias *= k
                self.info("Increased lr of %r by %.2f, new_lr %.2e", gd, k,
                          gd.learning_rate)
                self._stash(kv)
        elif not self._first_run:
            if any(self._has_nans(kv) for kv in self._gds):
                self.warning("NaNs encountered, will rollback")
                self._minus_steps = self.minus_steps
            self._minus_steps += 1
            if self._minus_steps < self.minus_steps:
                return
            self._minus_steps = 0
            self._plus_steps = 0
            for kv in self._gds:
                k = kv["lr_minus"] if kv["lr_minus"] is not None else self.lr_minus
                gd = kv["gd"]
                gd.learning_rate *= k
                gd.learning_rate_bias *= k
                self.info("Decreased lr of %r by %.2f, new_lr %.2e", gd, k,
                          gd.learning_rate)
                self._restore(kv, 0)
        self._first_run = False
