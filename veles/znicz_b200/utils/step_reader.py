"""Per-step device→host read of the evaluator's result (end-to-end serving of metrics).

The reference reads ``n_err`` only at epoch ends (``DecisionGD.on_last_minibatch``,
/root/reference/decision.py:443-476). Production monitoring and the end-to-end benchmark
want the step result on the host every step: this hook copies the evaluator's ``n_err``
(2×int32) into pinned memory asynchronously right after the forward segment was enqueued
and exposes the most recently *completed* value without stalling the pipeline.
"""
from __future__ import annotations

import torch


class StepResultReader(object):
    def __init__(self, evaluator, depth=4):
        self.evaluator = evaluator
        self.depth = depth
        self.slots = [torch.zeros(2, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.views = [t.numpy() for t in self.slots]      # host reads without a torch dispatch
        self.events = [torch.cuda.Event() for _ in range(depth)]
        self.pending = [False] * depth
        self.i = 0
        self.last = None
        self.bytes_per_step = 8

    def __call__(self, workflow):
        ev = self.evaluator
        if not getattr(ev, "on_cuda", False) or not ev.n_err:
            return
        s = self.i % self.depth
        if self.pending[s]:
            self.events[s].synchronize()
            self.last = self.views[s].copy()
        ext = getattr(getattr(ev, "device", None), "ext", None)
        if ext is not None and hasattr(ext, "push_to_host"):
            ext.push_to_host(ev.n_err.devmem, self.slots[s])   # a kernel stores into pinned memory
        else:
            self.slots[s].copy_(ev.n_err.devmem, non_blocking=True)
        self.events[s].record()
        self.pending[s] = True
        self.i += 1
