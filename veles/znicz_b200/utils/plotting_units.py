"""Plotting units (``veles.plotting_units`` equivalents).

The reference links live matplotlib plotters into the loop
(/root/reference/standard_workflow.py:672-1101). Here every plotter is a *recording*
unit: it always accumulates the values it would draw (``.values`` — handy for tests
and for publishing) and renders a PNG with the Agg backend only when
``root.common.disable.plotting`` is False.
"""
from __future__ import annotations

import os

import numpy

from ..core.config import root
from ..core.memory import Array
from ..core.units import Unit


def _plotting_enabled():
    return not root.common.disable.get("plotting", True)


def _figure_dir():
    d = os.path.join(str(root.common.dirs.get("cache", ".")), "plots")
    os.makedirs(d, exist_ok=True)
    return d


class Plotter(Unit):
    hide_from_registry = True

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("view_group", "PLOTTER")
        super().__init__(workflow, **kwargs)
        self.redraw_plot = kwargs.get("redraw_plot", True)
        self.last_file = None

    def initialize(self, **kwargs):
        pass

    def run(self):
        self.record()
        if _plotting_enabled():
            try:
                self.redraw()
            except Exception as e:  # plotting must never kill training
                self.warning("plotting failed: %s", e)

    def record(self):
        pass

    def redraw(self):
        pass

    def _savefig(self, fig):
        path = os.path.join(_figure_dir(), "%s.png" % self.name.replace(" ", "_"))
        fig.savefig(path)
        self.last_file = path
        import matplotlib.pyplot as plt
        plt.close(fig)

    @staticmethod
    def _pyplot():
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
        return plt

    def _resolve_input(self):
        v = getattr(self, "input", None)
        field = getattr(self, "input_field", None)
        if field is not None and v is not None:
            v = v[field] if not isinstance(field, str) else getattr(v, field)
        off = getattr(self, "input_offset", None)
        if off is not None and v is not None and hasattr(v, "__getitem__"):
            v = v[off]
        if isinstance(v, Array):
            v.map_read()
            v = v.mem
        return v


class AccumulatingPlotter(Plotter):
    """Appends a scalar each run and draws the curve."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.plot_style = kwargs.get("plot_style", "k-")
        self.clear_plot = kwargs.get("clear_plot", False)
        self.values = []
        self.input = None
        self.input_field = None
        self.input_offset = None
        self.fit_poly_power = kwargs.get("fit_poly_power", 0)

    def record(self):
        v = self._resolve_input()
        if v is None:
            return
        try:
            self.values.append(float(v))
        except (TypeError, ValueError):
            pass

    def redraw(self):
        plt = self._pyplot()
        fig = plt.figure(self.name)
        ax = fig.add_subplot(111)
        ax.plot(self.values, self.plot_style)
        ax.set_title(self.name)
        self._savefig(fig)


class MatrixPlotter(Plotter):
    """Keeps the last matrix (e.g. a confusion matrix) and draws it as a table."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.input = None
        self.input_field = None
        self.matrix = None

    def record(self):
        v = self._resolve_input()
        if v is not None:
            self.matrix = numpy.array(v)

    def redraw(self):
        if self.matrix is None:
            return
        plt = self._pyplot()
        fig = plt.figure(self.name)
        ax = fig.add_subplot(111)
        ax.imshow(self.matrix, interpolation="nearest")
        self._savefig(fig)


class MultiHistogram(Plotter):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.input = None
        self.n_bars = kwargs.get("n_bars", 25)
        self.hist_number = kwargs.get("hist_number", 16)
        self.histograms = None

    def record(self):
        v = self._resolve_input()
        if v is None:
            return
        m = v.reshape(v.shape[0], -1)[:self.hist_number]
        self.histograms = [numpy.histogram(row, bins=self.n_bars)[0] for row in m]

    def redraw(self):
        if not self.histograms:
            return
        plt = self._pyplot()
        n = len(self.histograms)
        cols = int(numpy.ceil(numpy.sqrt(n)))
        rows = int(numpy.ceil(n / cols))
        fig = plt.figure(self.name)
        for i, h in enumerate(self.histograms):
            ax = fig.add_subplot(rows, cols, i + 1)
            ax.bar(range(len(h)), h)
            ax.set_xticks([])
            ax.set_yticks([])
        self._savefig(fig)


class Histogram(MultiHistogram):
    pass


class TableMaxMin(Plotter):
    """Table with max/min of registered arrays."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.sources = []
        self.rows = []

    def add(self, unit, attr):
        self.sources.append((unit, attr))

    def record(self):
        self.rows = []
        for unit, attr in self.sources:
            a = getattr(unit, attr, None)
            if isinstance(a, Array) and a:
                a.map_read()
                self.rows.append(("%s.%s" % (unit.name, attr), float(a.mem.max()),
                                  float(a.mem.min())))

    def redraw(self):
        for name, mx, mn in self.rows:
            self.info("%-40s max %.6f min %.6f", name, mx, mn)


class ImagePlotter(Plotter):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.inputs = []
        self.input_fields = []
        self.images = []

    def record(self):
        self.images = []
        for arr, field in zip(self.inputs, self.input_fields):
            if isinstance(arr, Array) and arr:
                arr.map_read()
                self.images.append(numpy.array(arr.mem[field]))

    def redraw(self):
        if not self.images:
            return
        plt = self._pyplot()
        fig = plt.figure(self.name)
        for i, img in enumerate(self.images):
            ax = fig.add_subplot(1, len(self.images), i + 1)
            if img.ndim == 1:
                side = int(numpy.sqrt(img.size))
                img = img[:side * side].reshape(side, side)
            elif img.ndim == 3 and img.shape[2] == 1:
                img = img[:, :, 0]
            ax.imshow(img, interpolation="nearest")
        self._savefig(fig)


class ImmediatePlotter(Plotter):
    """Plots several 1-D arrays on one axis."""

    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.inputs = []
        self.input_fields = []
        self.input_styles = []
        self.curves = []

    def record(self):
        self.curves = []
        for arr, field in zip(self.inputs, self.input_fields):
            if isinstance(arr, Array) and arr:
                arr.map_read()
                self.curves.append(numpy.array(arr.mem[field]).ravel())

    def redraw(self):
        if not self.curves:
            return
        plt = self._pyplot()
        fig = plt.figure(self.name)
        ax = fig.add_subplot(111)
        for c, st in zip(self.curves, self.input_styles or ["k-"] * len(self.curves)):
            ax.plot(c, st)
        self._savefig(fig)


class SlaveStats(Plotter):
    """Per-worker throughput table (``veles.plotting_units.SlaveStats`` in
    /root/reference/tests/research/MnistSimple/mnist.py:180). In the one-process-per-GPU
    design a "slave" is a data-parallel rank: each run records
    (rank, world_size, minibatches seen, wall seconds since the previous record)."""

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Slave Stats")
        super().__init__(workflow, **kwargs)
        self.records = []
        self._count = 0
        self._last = None

    def record(self):
        import time
        now = time.time()
        self._count += 1
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        self.records.append((rank, world, self._count,
                             0.0 if self._last is None else now - self._last))
        self._last = now
        if len(self.records) > 4096:
            del self.records[:2048]

    def redraw(self):
        if self.records:
            r = self.records[-1]
            self.info("rank %d/%d: %d minibatches, last step %.4f s", *r)
