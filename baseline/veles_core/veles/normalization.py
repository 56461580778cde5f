"""``veles.normalization``: the normalizer family the loaders / evaluators refer to."""
import numpy

from veles.mapped_object_registry import MappedObjectsRegistry


class NormalizerRegistry(MappedObjectsRegistry):
    mapping = "normalizers"
    base = object


class NormalizerBase(object, metaclass=NormalizerRegistry):
    MAPPING = None

    def __init__(self, state=None, **kwargs):
        self._initialized = False

    @property
    def is_initialized(self):
        return self._initialized

    def analyze(self, data):
        self._initialized = True

    def normalize(self, data):
        return data

    def denormalize(self, data, **kwargs):
        return data

    def analyze_and_normalize(self, data):
        self.analyze(data)
        return self.normalize(data)

    @property
    def coefficients(self):
        return None


class NoneNormalizer(NormalizerBase):
    MAPPING = "none"


class LinearNormalizer(NormalizerBase):
    MAPPING = "linear"

    def __init__(self, state=None, **kwargs):
        super(LinearNormalizer, self).__init__(state, **kwargs)
        self.interval = kwargs.get("interval", (-1, 1))

    def normalize(self, data):
        lo = data.min(axis=tuple(range(1, data.ndim)), keepdims=True)
        hi = data.max(axis=tuple(range(1, data.ndim)), keepdims=True)
        d = numpy.where(hi - lo == 0, 1, hi - lo)
        a, b = self.interval
        data[...] = (data - lo) / d * (b - a) + a
        return data


class InternalMeanNormalizer(NormalizerBase):
    MAPPING = "internal_mean"

    def __init__(self, state=None, **kwargs):
        super(InternalMeanNormalizer, self).__init__(state, **kwargs)
        self.scale = kwargs.get("scale", 1)
        self._sum = None
        self._count = 0

    def analyze(self, data):
        s = data.sum(axis=0, dtype=numpy.float64)
        self._sum = s if self._sum is None else self._sum + s
        self._count += data.shape[0]
        self._initialized = True

    @property
    def mean(self):
        return (self._sum / max(self._count, 1))

    def normalize(self, data):
        data -= self.mean.astype(data.dtype)
        if self.scale != 1:
            data *= self.scale
        return data

    def denormalize(self, data, **kwargs):
        return data / self.scale + self.mean.astype(data.dtype)

    @property
    def coefficients(self):
        return self.mean, self.scale


class MeanDispersionNormalizer(NormalizerBase):
    MAPPING = "mean_disp"

    def __init__(self, state=None, **kwargs):
        super(MeanDispersionNormalizer, self).__init__(state, **kwargs)
        self._sum = None
        self._min = None
        self._max = None
        self._count = 0

    def analyze(self, data):
        s = data.sum(axis=0, dtype=numpy.float64)
        self._sum = s if self._sum is None else self._sum + s
        mn, mx = data.min(axis=0), data.max(axis=0)
        self._min = mn if self._min is None else numpy.minimum(self._min, mn)
        self._max = mx if self._max is None else numpy.maximum(self._max, mx)
        self._count += data.shape[0]
        self._initialized = True

    def normalize(self, data):
        mean = (self._sum / self._count).astype(data.dtype)
        disp = (self._max - self._min).astype(data.dtype)
        disp[disp == 0] = 1
        data -= mean
        data /= disp
        return data


def factory(name, **kwargs):
    return NormalizerRegistry.normalizers[name](**kwargs)
