"""LabelsPrinter: logs the top-k labels of the first sample.
Parity: /root/reference/labels_printer.py:45-68."""
from __future__ import annotations

import numpy

from ..core.units import Unit


class LabelsPrinter(Unit):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.top_number = kwargs.get("top_number", 5)
        self.top = []
        self.demand("input")

    def initialize(self, **kwargs):
        pass

    def run(self):
        self.input.map_read()
        mem = self.input.mem[0].ravel()
        rlm = getattr(self, "reversed_labels_mapping", None)
        labels = sorted(((float(v), rlm[i] if rlm else i) for i, v in enumerate(mem)),
                        key=lambda t: -t[0])
        self.top = labels[:self.top_number]
        rows = "\n".join("  %-20s %.5f" % (str(l), v) for v, l in self.top)
        self.info("Results:\n%s", rows)
        mean = float(numpy.mean(mem))
        if mean:
            self.info("Max to mean ratio: %.1f", float(numpy.max(mem)) / mean)
