"""Kernel diversity analysis: groups of near-duplicate convolutional kernels.
Parity: /root/reference/diversity.py:57-197 (get_similar_kernels, SimilarWeights2D).

Each pair of kernels is compared by (a) distance of the cross-correlation peak from
the centre, (b) kurtosis of the correlation surface (peak sharpness) and (c) norm of
the normalised difference; pairs passing all three filters form a graph whose cliques
are the "similar" sets. The pairwise work is batched per kernel with FFT-free
``scipy.signal.correlate2d`` on the symmetric boundary, like the reference."""
from __future__ import annotations

from collections import namedtuple

import numpy

from .nn_plotting_units import Weights2D

SimilarityCalculationParameters = namedtuple(
    "SimilarityCalculationParameters",
    ["form_threshold", "peak_threshold", "magnitude_threshold"])


def _unit(v):
    n = numpy.linalg.norm(v)
    return v / n if n else v


def get_similar_kernels(weights, channels=3,
                        params=SimilarityCalculationParameters(1.1, .5, .65)):
    import scipy.signal
    import scipy.stats
    weights = numpy.asarray(weights, dtype=numpy.float64)
    N = weights.shape[0]
    S = int(numpy.sqrt(weights.shape[1] / channels))
    side = 2 * S - 1
    centre = side // 2
    maxdist = numpy.sqrt(2) * centre or 1.0
    planes = [weights[:, c::channels][:, :S * S].reshape(N, S, S) for c in range(channels)]
    unit_planes = [numpy.stack([_unit(p[i].ravel()) for i in range(N)]) for p in planes]
    form = numpy.zeros((N, N))
    magn = numpy.zeros((N, N))
    kurt = numpy.full((N, N), numpy.nan)
    for x in range(N):
        for y in range(N):
            if x == y:
                continue
            corr = numpy.zeros((side, side))
            for p in planes:
                corr += scipy.signal.correlate2d(p[x], p[y], boundary="symm")
            px, py = numpy.unravel_index(numpy.argmax(corr), corr.shape)
            form[x, y] = 1 - numpy.hypot(px - centre, py - centre) / maxdist
            kurt[x, y] = scipy.stats.kurtosis(corr.ravel(), bias=False)
            diff = sum(float(numpy.sum((up[x] - up[y]) ** 2)) for up in unit_planes)
            magn[x, y] = 1 - numpy.sqrt(diff)

    def cut(values, k, lo, hi=0.95):
        return max(min(hi, values.mean() + values.std() * k), lo) if values.size else lo

    mask = magn > cut(magn[magn > 0], params.magnitude_threshold, 0.75)
    finite = kurt[~numpy.isnan(kurt)]
    if finite.size:
        kurt[numpy.isnan(kurt)] = finite.min()
        mask &= kurt > finite.mean() + finite.std() * params.peak_threshold
    mask &= form > cut(form[form > 0], params.form_threshold, 0.8)
    mask &= mask.T                       # symmetric boundary is not symmetric in (x, y)
    numpy.fill_diagonal(mask, False)

    sets, visited = [], set()
    for x in range(N):                   # greedy clique growth from every unvisited node
        if x in visited:
            continue
        clique, rejected, stack = {x}, set(), [x]
        while stack:
            cur = stack.pop()
            for y in numpy.flatnonzero(mask[cur]):
                y = int(y)
                if y in clique or y in rejected:
                    continue
                if all(mask[y, z] for z in clique):
                    clique.add(y)
                    stack.append(y)
                else:
                    rejected.add(y)
        if len(clique) > 1:
            sets.append(clique)
            visited |= clique
    return sets


class SimilarWeights2D(Weights2D):
    """Weights2D showing only the kernels that have near-duplicates."""

    def __init__(self, workflow, **kwargs):
        kwargs["split_channels"] = False
        super().__init__(workflow, **kwargs)
        self.form_threshold = kwargs.get("form_threshold", 1.1)
        self.peak_threshold = kwargs.get("peak_threshold", 0.5)
        self.magnitude_threshold = kwargs.get("magnitude_threshold", 0.65)
        self.similar_sets = []

    def prepare_pics(self, inp, transposed):
        inp = inp.reshape(inp.shape[0], -1)
        if transposed:
            inp = inp.transpose()
        n_channels, _, _ = self.get_number_of_channels(inp)
        if n_channels is None:
            return None
        self.similar_sets = get_similar_kernels(
            inp, n_channels, SimilarityCalculationParameters(
                self.form_threshold, self.peak_threshold, self.magnitude_threshold))
        rows = [inp[s] for group in self.similar_sets for s in sorted(group)]
        if not rows:
            return []
        return super().prepare_pics(numpy.stack(rows), False)
