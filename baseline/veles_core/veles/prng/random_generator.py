import numpy

_gens = {}


class RandomGenerator(object):
    def __init__(self, key=None, seed=None):
        self._key = key
        self.seed(1234 if seed is None else seed)

    def seed(self, seed, dtype=None, count=None):
        if isinstance(seed, numpy.ndarray):
            seed = int(numpy.frombuffer(seed.tobytes()[:4].ljust(4, b"\0"), numpy.uint32)[0])
        elif isinstance(seed, (bytes, bytearray)):
            seed = int.from_bytes(bytes(seed[:4]).ljust(4, b"\0"), "little")
        self._seed = seed
        self.state = numpy.random.RandomState(int(seed) & 0xFFFFFFFF)

    def fill(self, arr, vle_min=-1.0, vle_max=1.0):
        arr[...] = self.state.uniform(vle_min, vle_max, arr.shape).astype(arr.dtype)

    def fill_normal_real(self, arr, mean, stddev, clip_to_sigma=5.0):
        v = self.state.normal(mean, stddev, arr.shape) if stddev > 0 else numpy.full(arr.shape, mean)
        if stddev > 0 and clip_to_sigma:
            numpy.clip(v, mean - clip_to_sigma * stddev, mean + clip_to_sigma * stddev, out=v)
        arr[...] = v.astype(arr.dtype)

    def normal(self, loc=0.0, scale=1.0, size=None):
        return self.state.normal(loc, scale, size)

    def rand(self, *shape):
        return self.state.rand(*shape)

    def random_sample(self, size=None):
        return self.state.random_sample(size)

    def randint(self, low, high=None, size=None):
        return self.state.randint(low, high, size)

    def shuffle(self, arr):
        self.state.shuffle(arr)

    def permutation(self, x):
        return self.state.permutation(x)

    def choice(self, *a, **kw):
        return self.state.choice(*a, **kw)

    def bytes(self, n):
        return self.state.bytes(n)


def get(key=1):
    g = _gens.get(key)
    if g is None:
        g = _gens[key] = RandomGenerator(key, seed=1234 + 7919 * int(key))
    return g
