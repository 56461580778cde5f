"""NN-specific plotters: Weights2D, MSEHistogram and the Kohonen map plotters.
Parity: /root/reference/nn_plotting_units.py:52 (Weights2D), :220 (MSEHistogram),
:411 (KohonenHits), :497 (KohonenInputMaps), :590 (KohonenNeighborMap),
:767 (KohonenValidationResults).

All of them are recording units (``Plotter`` in plotting_units.py): ``record()``
computes the data to draw (``pics``, ``val_mse``, ``hits``, ``maps``, ``link_values``,
``cells``); ``redraw()`` rasterises it with PIL (matplotlib is optional and absent from
the target image) into ``<cache>/plots/<name>.png``.
"""
from __future__ import annotations

import os

import numpy

from ..core.memory import Array, roundup
from ..core.mutable import Bool
from .image_saver import normalize_image
from .plotting_units import Plotter, _figure_dir


def _host(v):
    if isinstance(v, Array):
        v.map_read()
        return v.mem
    return v


def _heat(v):
    """Scalar in [0,1] → yellow-orange-red RGB (the reference's default YlOrRd scheme)."""
    v = float(min(max(v, 0.0), 1.0))
    return (255, int(255 - 175 * v), int(204 * (1 - v) ** 2))


class _PILPlotter(Plotter):
    hide_from_registry = True
    CELL = 48

    def _save_image(self, img):
        path = os.path.join(_figure_dir(), "%s.png" % self.name.replace(" ", "_"))
        img.save(path)
        self.last_file = path
        return path

    def _hex_canvas(self, width, height):
        from PIL import Image, ImageDraw
        c = self.CELL
        img = Image.new("RGB", (int((width + 1) * c), int((height * 0.87 + 0.6) * c)),
                        "white")
        return img, ImageDraw.Draw(img)

    def _hex_center(self, x, y):
        c = self.CELL
        return ((x + (0.5 if y & 1 else 0.0) + 0.6) * c, (y * 0.87 + 0.6) * c)

    def _hexagon(self, draw, x, y, scale, fill, outline="black"):
        cx, cy = self._hex_center(x, y)
        r = self.CELL * 0.577 * scale
        pts = [(cx + r * numpy.sin(a), cy + r * numpy.cos(a))
               for a in numpy.arange(6) * numpy.pi / 3]
        draw.polygon(pts, fill=fill, outline=outline)
        return cx, cy


class Weights2D(_PILPlotter):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Weights")
        super().__init__(workflow, **kwargs)
        self.color_space = kwargs.get("color_space", "RGB")
        self.get_shape_from = None
        self.limit = kwargs.get("limit", 64)
        self.transposed = kwargs.get("transposed", False)
        self.yuv = Bool(kwargs.get("yuv", False))
        self.split_channels = kwargs.get("split_channels", False)
        self.column_align = kwargs.get("column_align", 4)
        self.pics = []
        self.demand("input")

    def get_number_of_channels(self, inp):
        src = self.input if self.get_shape_from is None else self.get_shape_from
        n_channels = 1
        shape = src.shape if isinstance(src, Array) else tuple(
            ((s.shape[-1] if s else None) if isinstance(s, Array) else s) for s in src)
        if isinstance(src, Array):
            if len(shape) < 2:
                return None, None, None
            if len(shape) == 2:
                n = inp.shape[1] if inp is not None else shape[1]
                sx = int(numpy.round(numpy.sqrt(n)))
                sy = n // sx if sx else 0
                if sx * sy != n:
                    return None, None, None
            else:
                sy, sx = shape[1], shape[2]
                if len(shape) == 4:
                    n_channels = shape[3]
        elif len(shape) == 2:
            sx, sy = shape
        else:
            sx, sy, n_channels = shape[-2], shape[-3], shape[-1]
        if n_channels is None or sx is None or sy is None:
            return None, None, None
        return int(n_channels), int(sx), int(sy)

    def prepare_pics(self, inp, transposed):
        if not isinstance(inp, numpy.ndarray) or inp.ndim < 2:
            raise ValueError("input should be a numpy array (2D at least)")
        inp = inp.reshape(inp.shape[0], -1)
        if transposed:
            inp = inp.transpose()
        n_channels, sx, sy = self.get_number_of_channels(inp)
        if n_channels is None:
            return None
        sz = sx * sy * n_channels
        pics = []
        for row in inp:
            if len(pics) >= self.limit:
                break
            mem = row.ravel()[:sz]
            if mem.size < sz:
                return None
            if n_channels <= 1:
                pics.append(normalize_image(mem.reshape(sy, sx)))
                continue
            w = mem.reshape(sy, sx, n_channels)
            if self.split_channels:
                for ch in range(n_channels):
                    pics.append(normalize_image(w[:, :, ch]))
            elif n_channels == 2:
                pics.append(normalize_image(w[:, :, 0]))
            else:
                pics.append(normalize_image(w[:, :, :3], self.color_space))
        return pics[:self.limit]

    def record(self):
        mem = _host(self.input)
        if mem is None:
            return
        self.pics = self.prepare_pics(mem[:self.limit] if not self.transposed else mem,
                                      self.transposed) or []

    def redraw(self):
        from PIL import Image
        pics = self.pics
        if not pics:
            return None
        n_cols = roundup(int(numpy.round(numpy.sqrt(len(pics)))), self.column_align)
        n_rows = int(numpy.ceil(len(pics) / n_cols))
        h, w = pics[0].shape[:2]
        scale = max(1, 32 // max(h, w))
        sheet = Image.new("RGB", (n_cols * (w * scale + 2), n_rows * (h * scale + 2)),
                          "white")
        for i, p in enumerate(pics):
            im = Image.fromarray(p).convert("RGB").resize((w * scale, h * scale),
                                                          Image.NEAREST)
            sheet.paste(im, ((i % n_cols) * (w * scale + 2) + 1,
                             (i // n_cols) * (h * scale + 2) + 1))
        return self._save_image(sheet)


class MSEHistogram(_PILPlotter):
    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Histogram")
        super().__init__(workflow, **kwargs)
        self.n_bars = kwargs.get("n_bars", 35)
        self.val_mse = numpy.zeros(self.n_bars, numpy.float32)
        self.mse_min = self.mse_max = None
        self.val_max = self.val_min = 0
        self.demand("mse")

    def fill(self):
        mem = numpy.asarray(_host(self.mse)).ravel()
        self.mse_max, self.mse_min = float(mem.max()), float(mem.min())
        d = self.mse_max - self.mse_min
        if not d:
            return
        idx = numpy.floor((mem - self.mse_min) * ((self.n_bars - 1) / d)).astype(int)
        self.val_mse = numpy.bincount(idx, minlength=self.n_bars).astype(numpy.float32)
        self.val_max, self.val_min = self.val_mse.max(), self.val_mse.min()

    record = fill

    def redraw(self):
        from PIL import Image, ImageDraw
        W, H = 20 * self.n_bars, 240
        img = Image.new("RGB", (W, H), "#ffe6ca")
        d = ImageDraw.Draw(img)
        top = max(float(self.val_max), 1.0)
        for i, v in enumerate(self.val_mse):
            d.rectangle([i * 20 + 2, H - 10 - int((H - 30) * v / top), i * 20 + 18, H - 10],
                        fill="#ffa0ef", outline="red")
        d.text((4, 2), "%s  min %.4g  max %.4g" % (self.name, self.mse_min or 0,
                                                  self.mse_max or 0), fill="black")
        return self._save_image(img)


class KohonenGridBase(_PILPlotter):
    def __init__(self, workflow, **kwargs):
        super().__init__(workflow, **kwargs)
        self.demand("shape")

    width = property(lambda self: int(self.shape[0]))
    height = property(lambda self: int(self.shape[1]))


class KohonenHits(KohonenGridBase):
    """Winner counts per neuron drawn as hexagons scaled by hit share."""
    SIZE_TEXT_THRESHOLD = 0.33

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Kohonen Hits")
        super().__init__(workflow, **kwargs)
        self.color_bins = kwargs.get("color_bins", "#666699")
        self.color_text = kwargs.get("color_text", "white")
        self.hits = None
        self.demand("input")

    def record(self):
        self.hits = numpy.array(_host(self.input)).reshape(self.height, self.width)

    def redraw(self):
        img, d = self._hex_canvas(self.width, self.height)
        mx = max(float(self.hits.max()), 1.0)
        for y in range(self.height):
            for x in range(self.width):
                self._hexagon(d, x, y, 1.0, "white")
                n = float(self.hits[y, x])
                if n:
                    s = numpy.sqrt(n / mx)
                    cx, cy = self._hexagon(d, x, y, s, self.color_bins, None)
                    if s > self.SIZE_TEXT_THRESHOLD:
                        d.text((cx - 6, cy - 5), "%d" % n, fill=self.color_text)
        return self._save_image(img)


class KohonenInputMaps(_PILPlotter):
    """One heat map of the neuron grid per input feature."""

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Kohonen Maps")
        super().__init__(workflow, **kwargs)
        self.color_scheme = kwargs.get("color_scheme", "YlOrRd")
        self.color_grid = kwargs.get("color_grid", "none")
        self.maps = None
        self.demand("input", "shape")

    width = property(lambda self: int(self.shape[0]))
    height = property(lambda self: int(self.shape[1]))

    def record(self):
        w = numpy.array(_host(self.input), dtype=numpy.float32)
        w = w.reshape(self.width * self.height, -1)
        lo, hi = w.min(axis=0), w.max(axis=0)
        span = numpy.where(hi > lo, hi - lo, 1.0)
        self.maps = ((w - lo) / span).T.reshape(-1, self.height, self.width)

    def redraw(self):
        from PIL import Image
        tiles = []
        save_cell, self.CELL = self.CELL, 16
        try:
            for m in self.maps:
                img, d = self._hex_canvas(self.width, self.height)
                for y in range(self.height):
                    for x in range(self.width):
                        self._hexagon(d, x, y, 1.0, _heat(m[y, x]), None)
                tiles.append(img)
        finally:
            self.CELL = save_cell
        n_cols = int(numpy.ceil(numpy.sqrt(len(tiles))))
        n_rows = int(numpy.ceil(len(tiles) / n_cols))
        tw, th = tiles[0].size
        sheet = Image.new("RGB", (n_cols * tw, n_rows * th), "white")
        for i, t in enumerate(tiles):
            sheet.paste(t, ((i % n_cols) * tw, (i // n_cols) * th))
        return self._save_image(sheet)


class KohonenNeighborMap(_PILPlotter):
    """Weight-space distance between neighbouring neurons of the hexagonal grid."""
    NEURON_SIZE = 0.4

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Kohonen Neighbor Weight Distances")
        super().__init__(workflow, **kwargs)
        self.color_neurons = kwargs.get("color_neurons", "#666699")
        self.color_scheme = kwargs.get("color_scheme", "YlOrRd")
        self.links = []          # ((x1, y1), (x2, y2))
        self.link_values = None
        self.demand("input", "shape")

    width = property(lambda self: int(self.shape[0]))
    height = property(lambda self: int(self.shape[1]))

    def neighbor_pairs(self):
        W, H = self.width, self.height
        pairs = [((x, y), (x + 1, y)) for y in range(H) for x in range(W - 1)]
        for y in range(H - 1):
            for x in range(W):
                pairs.append(((x, y), (x, y + 1)))
                x2 = x + 1 if y & 1 else x - 1       # the other diagonal neighbour
                if 0 <= x2 < W:
                    pairs.append(((x, y), (x2, y + 1)))
        return pairs

    def record(self):
        w = numpy.array(_host(self.input), dtype=numpy.float32)
        w = w.reshape(self.width * self.height, -1)
        self.links = self.neighbor_pairs()
        a = numpy.array([y * self.width + x for (x, y), _ in self.links])
        b = numpy.array([y * self.width + x for _, (x, y) in self.links])
        self.link_values = numpy.linalg.norm(w[a] - w[b], axis=1)

    def redraw(self):
        img, d = self._hex_canvas(self.width, self.height)
        lv = self.link_values
        lo, hi = float(lv.min()), float(lv.max())
        span = (hi - lo) or 1.0
        for (n1, n2), v in zip(self.links, lv):
            d.line([self._hex_center(*n1), self._hex_center(*n2)],
                   fill=_heat((v - lo) / span), width=max(2, self.CELL // 5))
        for y in range(self.height):
            for x in range(self.width):
                self._hexagon(d, x, y, self.NEURON_SIZE, self.color_neurons, None)
        return self._save_image(img)


class KohonenValidationResults(KohonenGridBase):
    """Neuron → category map with per-neuron fitness."""

    def __init__(self, workflow, **kwargs):
        kwargs.setdefault("name", "Kohonen Validation Results")
        super().__init__(workflow, **kwargs)
        self.color_text = kwargs.get("color_text", "white")
        self.cmap = kwargs.get("color_map", "Set1")
        self.cells = None        # [H][W] → (label or None, hits, fitness)
        self.demand("input", "result", "fitness", "fitness_by_label",
                    "fitness_by_neuron")

    PALETTE = ("#e41a1c", "#377eb8", "#4daf4a", "#984ea3", "#ff7f00", "#ffff33",
               "#a65628", "#f781bf", "#999999")

    def record(self):
        hits = numpy.array(_host(self.input)).ravel()
        owner = {}
        res = self.result
        pairs = res.items() if isinstance(res, dict) else enumerate(res)
        labels = {}
        for label, neurons in pairs:
            idx = labels.setdefault(label, len(labels))
            for n in neurons:
                owner[int(n)] = idx
        fbn = self.fitness_by_neuron
        self.cells = [[(owner.get(y * self.width + x), int(hits[y * self.width + x]),
                        float(fbn[y * self.width + x]) if len(fbn) > y * self.width + x
                        else 0.0)
                       for x in range(self.width)] for y in range(self.height)]

    def redraw(self):
        img, d = self._hex_canvas(self.width, self.height)
        mx = max(max(c[1] for row in self.cells for c in row), 1)
        for y, row in enumerate(self.cells):
            for x, (label, n, fit) in enumerate(row):
                self._hexagon(d, x, y, 1.0, "white")
                if label is None or not n:
                    continue
                cx, cy = self._hexagon(d, x, y, max(numpy.sqrt(n / mx), 0.3),
                                       self.PALETTE[label % len(self.PALETTE)], None)
                d.text((cx - 8, cy - 5), "%d" % int(fit * 100), fill=self.color_text)
        d.text((2, 2), "fitness %.2f" % float(self.fitness), fill="black")
        return self._save_image(img)
