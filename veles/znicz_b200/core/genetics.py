"""Hyper-parameter search over ``Range`` markers in the config tree.

The reference marks tunable values with ``Range(default, min, max)`` in sample configs
(/root/reference/samples/MNIST/mnist_config.py:56-86) and leaves the search to the core's
genetic optimiser (SURVEY §2.6 "Hyper-parameter search parallelism"). This module provides
both halves:

* ``Range`` / ``process_config`` — markers and their discovery; ``fix_config`` collapses
  them to their defaults for a normal run.
* ``GeneticsOptimizer`` — a small real-coded GA (tournament selection, blend crossover,
  gaussian mutation, elitism), deterministic for a given seed. Each chromosome is a
  vector of values for the discovered markers; ``evaluate(values) -> fitness`` is supplied
  by the caller (the launcher builds + runs the workflow and returns its fitness
  metric). With ``torchrun`` each rank evaluates a slice of the population and fitness
  values are all-gathered, so the search scales across the GPUs of a node.
"""
from __future__ import annotations

import numpy

from .config import Config


class Range(object):
    """A tunable value: ``Range(default, min, max)``; ints stay ints. ``Range(default,
    choices...)`` with non-numeric items enumerates choices."""

    def __init__(self, default, *bounds):
        self.default = default
        if len(bounds) == 2 and all(isinstance(b, (int, float)) for b in bounds) and \
                isinstance(default, (int, float)) and not isinstance(default, bool):
            self.min_value, self.max_value = bounds
            self.choices = None
            if not self.min_value <= default <= self.max_value:
                raise ValueError("Range default %r is outside [%r, %r]" % (
                    default, self.min_value, self.max_value))
        else:
            self.choices = [default] + [b for b in bounds if b != default]
            self.min_value, self.max_value = 0, len(self.choices) - 1

    @property
    def is_int(self):
        return self.choices is not None or (
            isinstance(self.default, int) and isinstance(self.min_value, int) and
            isinstance(self.max_value, int))

    def decode(self, gene):
        gene = min(max(gene, self.min_value), self.max_value)
        if self.choices is not None:
            return self.choices[int(round(gene))]
        return int(round(gene)) if self.is_int else float(gene)

    def encode_default(self):
        return 0.0 if self.choices is not None else float(self.default)

    def __repr__(self):
        if self.choices is not None:
            return "Range(%s)" % ", ".join(repr(c) for c in self.choices)
        return "Range(%r, %r, %r)" % (self.default, self.min_value, self.max_value)


def _walk(node, path, visit):
    if isinstance(node, Config):
        for k, v in list(node.__dict__.items()):
            if k == "__path__":
                continue
            r = _walk(v, path + (k,), visit)
            if r is not _KEEP:
                object.__setattr__(node, k, r)
        return _KEEP
    if isinstance(node, dict):
        for k, v in list(node.items()):
            r = _walk(v, path + (k,), visit)
            if r is not _KEEP:
                node[k] = r
        return _KEEP
    if isinstance(node, list):
        for i, v in enumerate(node):
            r = _walk(v, path + (i,), visit)
            if r is not _KEEP:
                node[i] = r
        return _KEEP
    if isinstance(node, Range):
        return visit(path, node)
    return _KEEP


_KEEP = object()


def process_config(cfg):
    """→ list of (path tuple, Range) found under ``cfg`` (Config, dict or list)."""
    found = []

    def visit(path, rng):
        found.append((path, rng))
        return _KEEP
    _walk(cfg, (), visit)
    return found


def fix_config(cfg):
    """Replace every ``Range`` by its default value in place."""
    _walk(cfg, (), lambda path, rng: rng.default)
    return cfg


def apply_values(cfg, markers, values):
    """Write decoded chromosome ``values`` at the marker paths (markers from
    ``process_config`` taken *before* any substitution)."""
    for (path, rng), gene in zip(markers, values):
        node = cfg
        for key in path[:-1]:
            node = getattr(node, key) if isinstance(node, Config) else node[key]
        val = rng.decode(gene)
        if isinstance(node, Config):
            object.__setattr__(node, path[-1], val)
        else:
            node[path[-1]] = val


class GeneticsOptimizer(object):
    def __init__(self, markers, evaluate, population_size=12, generations=5, seed=1,
                 elite=2, mutation_rate=0.25, mutation_scale=0.15, tournament=3,
                 rank=0, world_size=1, log=None):
        if not markers:
            raise ValueError("no Range markers to optimise")
        self.markers = markers
        self.evaluate = evaluate
        self.population_size = population_size
        self.generations = generations
        self.rs = numpy.random.RandomState(seed)
        self.elite = min(elite, population_size)
        self.mutation_rate = mutation_rate
        self.mutation_scale = mutation_scale
        self.tournament = tournament
        self.rank, self.world_size = rank, world_size
        self.log = log or (lambda *a: None)
        self.lo = numpy.array([m[1].min_value for m in markers], dtype=numpy.float64)
        self.hi = numpy.array([m[1].max_value for m in markers], dtype=numpy.float64)
        self.history = []
        self.best_values = None
        self.best_fitness = -numpy.inf
        self._cache = {}

    def decode(self, chromo):
        return [m[1].decode(g) for m, g in zip(self.markers, chromo)]

    def _fitness_all(self, pop):
        fit = numpy.full(len(pop), numpy.nan)
        for i, chromo in enumerate(pop):
            key = tuple(self.decode(chromo))
            if key in self._cache:
                fit[i] = self._cache[key]
            elif i % self.world_size == self.rank:
                fit[i] = float(self.evaluate(list(chromo)))
        if self.world_size > 1:
            import torch
            import torch.distributed as dist
            t = torch.from_numpy(numpy.nan_to_num(fit, nan=-numpy.inf))
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            fit = t.cpu().numpy()
        for chromo, f in zip(pop, fit):
            self._cache[tuple(self.decode(chromo))] = f
        return fit

    def run(self):
        n, d = self.population_size, len(self.markers)
        pop = self.lo + self.rs.rand(n, d) * (self.hi - self.lo)
        pop[0] = [m[1].encode_default() for m in self.markers]     # defaults compete too
        for gen in range(self.generations):
            fit = self._fitness_all(pop)
            order = numpy.argsort(-fit)
            pop, fit = pop[order], fit[order]
            if fit[0] > self.best_fitness:
                self.best_fitness, self.best_values = float(fit[0]), self.decode(pop[0])
            self.history.append((gen, float(fit[0]), float(numpy.mean(fit))))
            self.log("generation %d: best %.6g mean %.6g", gen, fit[0], numpy.mean(fit))
            if gen == self.generations - 1:
                break
            nxt = [pop[i].copy() for i in range(self.elite)]
            while len(nxt) < n:
                a = pop[min(self.rs.randint(0, n, self.tournament))]
                b = pop[min(self.rs.randint(0, n, self.tournament))]
                w = self.rs.rand(d) * 1.5 - 0.25                  # BLX-0.25 blend
                child = a * w + b * (1 - w)
                mut = self.rs.rand(d) < self.mutation_rate
                child = child + mut * self.rs.randn(d) * self.mutation_scale * (self.hi - self.lo)
                nxt.append(numpy.clip(child, self.lo, self.hi))
            pop = numpy.array(nxt)
        return self.best_values, self.best_fitness
