"""Dataset normalizers (``veles.normalization``).

Consumed by loaders (``normalization_type`` in configs, e.g.
/root/reference/samples/CIFAR10/cifar_caffe_config.py "internal_mean",
/root/reference/samples/Lines/lines_config.py "mean_disp") and by EvaluatorMSE,
which needs ``denormalize`` + ``coefficients`` (/root/reference/evaluator.py:381,434-436,
/root/reference/cuda/denormalization.jcu:13).

Every normalizer is ``y = x * mul + add`` with (mul, add) either scalars, per-feature
arrays, or per-sample values, so ``coefficients`` returns that pair and the device
evaluator can denormalise in-kernel.
"""
from __future__ import annotations

import numpy

from .registry import make_registry

NormalizerRegistry = make_registry("normalizers")


class NormalizerBase(object, metaclass=NormalizerRegistry):
    MAPPING = None
    stateless = False

    def __init__(self, state=None, **kwargs):
        self._initialized = False
        if state is not None:
            self.__dict__.update(state)
            self._initialized = True

    @property
    def is_initialized(self):
        return self._initialized or self.stateless

    def analyze(self, data):
        """Accumulate dataset statistics (may be called several times)."""
        self._initialized = True

    def normalize(self, data):
        raise NotImplementedError

    def denormalize(self, data):
        raise NotImplementedError

    @property
    def coefficients(self):
        """(mul, add) such that normalized = raw * mul + add, or None if per-sample."""
        return None

    def analyze_and_normalize(self, data):
        self.analyze(data)
        return self.normalize(data)

    @property
    def state(self):
        return {k: v for k, v in self.__dict__.items()}

    def reset(self):
        self._initialized = False


class NoneNormalizer(NormalizerBase):
    MAPPING = "none"
    stateless = True

    def normalize(self, data):
        return data

    def denormalize(self, data):
        return data

    @property
    def coefficients(self):
        return 1.0, 0.0


class LinearNormalizer(NormalizerBase):
    """Per-sample linear map of [min, max] onto ``interval`` (default [-1, 1])."""
    MAPPING = "linear"
    stateless = True

    def __init__(self, state=None, **kwargs):
        super().__init__(state, **kwargs)
        self.interval = tuple(kwargs.get("interval", (-1.0, 1.0)))

    def normalize(self, data):
        flat = data.reshape(data.shape[0], -1)
        mn = flat.min(axis=1, keepdims=True)
        mx = flat.max(axis=1, keepdims=True)
        diff = numpy.where(mx - mn == 0, 1, mx - mn)
        lo, hi = self.interval
        flat[...] = (flat - mn) * ((hi - lo) / diff) + lo
        return data

    def denormalize(self, data):
        raise ValueError("linear (per-sample) normalization is not invertible")


class RangeLinearNormalizer(NormalizerBase):
    """Dataset-wide linear map of [min, max] onto ``interval``."""
    MAPPING = "range_linear"

    def __init__(self, state=None, **kwargs):
        self.interval = tuple(kwargs.get("interval", (-1.0, 1.0)))
        self.vmin = None
        self.vmax = None
        super().__init__(state, **kwargs)

    def analyze(self, data):
        mn, mx = float(data.min()), float(data.max())
        self.vmin = mn if self.vmin is None else min(self.vmin, mn)
        self.vmax = mx if self.vmax is None else max(self.vmax, mx)
        self._initialized = True

    @property
    def coefficients(self):
        lo, hi = self.interval
        diff = (self.vmax - self.vmin) or 1.0
        mul = (hi - lo) / diff
        return mul, lo - self.vmin * mul

    def normalize(self, data):
        mul, add = self.coefficients
        data *= mul
        data += add
        return data

    def denormalize(self, data):
        mul, add = self.coefficients
        return (data - add) / mul


class MeanDispersionNormalizer(NormalizerBase):
    """``(x - mean) * rdisp`` per feature, rdisp = 1 / (max - min)."""
    MAPPING = "mean_disp"

    def __init__(self, state=None, **kwargs):
        self.sum = None
        self.count = 0
        self.fmin = None
        self.fmax = None
        super().__init__(state, **kwargs)

    def analyze(self, data):
        d = data.astype(numpy.float64)
        s = d.sum(axis=0)
        self.sum = s if self.sum is None else self.sum + s
        self.count += data.shape[0]
        mn, mx = d.min(axis=0), d.max(axis=0)
        self.fmin = mn if self.fmin is None else numpy.minimum(self.fmin, mn)
        self.fmax = mx if self.fmax is None else numpy.maximum(self.fmax, mx)
        self._initialized = True

    @property
    def mean(self):
        return self.sum / max(self.count, 1)

    @property
    def rdisp(self):
        disp = self.fmax - self.fmin
        return 1.0 / numpy.where(disp == 0, 1.0, disp)

    @property
    def coefficients(self):
        r = self.rdisp
        return r, -self.mean * r

    def normalize(self, data):
        data -= self.mean.astype(data.dtype)
        data *= self.rdisp.astype(data.dtype)
        return data

    def denormalize(self, data):
        return data / self.rdisp + self.mean


class PointwiseNormalizer(NormalizerBase):
    """Per-feature linear map of the dataset [min, max] onto [-1, 1]."""
    MAPPING = "pointwise"

    def __init__(self, state=None, **kwargs):
        self.fmin = None
        self.fmax = None
        super().__init__(state, **kwargs)

    def analyze(self, data):
        mn, mx = data.min(axis=0), data.max(axis=0)
        self.fmin = mn if self.fmin is None else numpy.minimum(self.fmin, mn)
        self.fmax = mx if self.fmax is None else numpy.maximum(self.fmax, mx)
        self._initialized = True

    @property
    def coefficients(self):
        disp = (self.fmax - self.fmin).astype(numpy.float64)
        mul = numpy.where(disp == 0, 0.0, 2.0 / numpy.where(disp == 0, 1.0, disp))
        add = numpy.where(disp == 0, 0.0, -1.0 - self.fmin * mul)
        return mul, add

    def normalize(self, data):
        mul, add = self.coefficients
        data *= mul.astype(data.dtype)
        data += add.astype(data.dtype)
        return data

    def denormalize(self, data):
        mul, add = self.coefficients
        safe = numpy.where(mul == 0, 1.0, mul)
        return (data - add) / safe


class InternalMeanNormalizer(NormalizerBase):
    """Subtract the per-feature dataset mean, then multiply by ``scale``."""
    MAPPING = "internal_mean"

    def __init__(self, state=None, **kwargs):
        self.sum = None
        self.count = 0
        self.scale = kwargs.get("scale", 1.0)
        super().__init__(state, **kwargs)

    def analyze(self, data):
        s = data.astype(numpy.float64).sum(axis=0)
        self.sum = s if self.sum is None else self.sum + s
        self.count += data.shape[0]
        self._initialized = True

    @property
    def mean(self):
        return self.sum / max(self.count, 1)

    @property
    def coefficients(self):
        return self.scale, -self.mean * self.scale

    def normalize(self, data):
        data -= self.mean.astype(data.dtype)
        if self.scale != 1.0:
            data *= self.scale
        return data

    def denormalize(self, data):
        return data / self.scale + self.mean


class ExternalMeanNormalizer(InternalMeanNormalizer):
    """Mean comes from ``mean_source`` (array or .npy path) instead of the data."""
    MAPPING = "external_mean"

    def __init__(self, state=None, **kwargs):
        super().__init__(state, **kwargs)
        src = kwargs.get("mean_source")
        if src is not None:
            if isinstance(src, str):
                src = numpy.load(src)
            self.sum = numpy.asarray(src, dtype=numpy.float64)
            self.count = 1
            self._initialized = True

    def analyze(self, data):
        if self.sum is None:
            raise ValueError("external_mean normalizer needs mean_source")
        self._initialized = True


def make_normalizer(name, **params):
    try:
        cls = NormalizerRegistry.registry[name]
    except KeyError:
        raise ValueError("Unknown normalization type %r (known: %s)" % (
            name, sorted(NormalizerRegistry.registry)))
    return cls(**params)
