"""``veles.memory``: Array = numpy host copy + device copy with explicit map/unmap coherence
(the protocol every reference unit follows: ``unmap_vectors`` before a kernel, ``map_read`` /
``map_write`` / ``map_invalidate`` before touching ``.mem``)."""
import numpy


def roundup(num, align):
    d = num % align
    return num if d == 0 else num + (align - d)


def reshape(a, shape):
    if a.shape == tuple(shape):
        return a
    return a.reshape(shape)


def ravel(a):
    return a.reshape(-1) if a.ndim != 1 else a


def reshape_transposed(w):
    """Interpret a [A, B] array as the transposed storage of a [B, A] matrix."""
    a = w.reshape(tuple(reversed(w.shape))[-2:]) if w.ndim == 2 else w.reshape(w.shape[1], w.shape[0])
    return a.transpose()


def transpose(a):
    return a.transpose()


def interleave(arr):
    """NCHW -> NHWC (or CHW -> HWC)."""
    if arr.ndim == 4:
        return numpy.ascontiguousarray(arr.transpose(0, 2, 3, 1))
    if arr.ndim == 3:
        return numpy.ascontiguousarray(arr.transpose(1, 2, 0))
    raise ValueError("unsupported number of dimensions")


def eq_addr(a, b):
    return a.__array_interface__["data"][0] == b.__array_interface__["data"][0]


def assert_addr(a, b):
    if not eq_addr(a, b):
        raise ValueError("addresses of two arrays differ")


MAP_NONE, MAP_READ, MAP_WRITE, MAP_INVALIDATE = 0, 1, 2, 3


class Array(object):
    """States: MAP_NONE = the device copy is authoritative (after unmap / initialize);
    MAP_READ = host copy valid, device untouched; MAP_WRITE / MAP_INVALIDATE = host copy
    modified, uploaded by the next ``unmap()``."""

    def __init__(self, data=None, shallow_pickle=False):
        self._mem = None
        self.device = None
        self.devmem = None
        self.map_flags = MAP_WRITE
        self.shallow_pickle = shallow_pickle
        self.max_supposed = 1.0
        if data is not None:
            self.mem = data

    # -- host array --------------------------------------------------------------------------
    @property
    def mem(self):
        return self._mem

    @mem.setter
    def mem(self, value):
        if self.devmem is not None and value is not None and \
                (self._mem is None or value.nbytes != self._mem.nbytes):
            raise ValueError("cannot resize an initialized Array (use reset())")
        self._mem = value

    @property
    def v(self):
        return self._mem

    def reset(self, new_mem=None):
        self.devmem = None
        self.device = None
        self.map_flags = MAP_WRITE
        self._mem = new_mem

    def __bool__(self):
        return self._mem is not None and self._mem.size > 0

    __nonzero__ = __bool__

    def __len__(self):
        return len(self._mem)

    def __getitem__(self, key):
        return self._mem[key]

    def __setitem__(self, key, value):
        self._mem[key] = value

    shape = property(lambda self: self._mem.shape,
                     lambda self, v: setattr(self._mem, "shape", v))
    size = property(lambda self: self._mem.size)
    dtype = property(lambda self: self._mem.dtype)
    itemsize = property(lambda self: self._mem.itemsize)
    nbytes = property(lambda self: self._mem.nbytes)

    @property
    def sample_size(self):
        return self._mem.size // self._mem.shape[0]

    @property
    def plain(self):
        return ravel(self._mem)

    @property
    def matrix(self):
        return reshape(self._mem, (self._mem.shape[0], self.sample_size))

    def __repr__(self):
        return "<Array %s %s>" % ("empty" if self._mem is None else self._mem.shape,
                                  "dev" if self.devmem is not None else "host")

    # -- device copy -------------------------------------------------------------------------
    def initialize(self, device):
        if self._mem is None or device is None or not getattr(device, "exists", False):
            return
        if self.devmem is not None and self.device is device and \
                self.devmem.size == self._mem.nbytes:
            return
        self._mem = numpy.ascontiguousarray(self._mem)
        self.device = device
        self.devmem = device.allocate(self._mem)
        self.devmem.to_device(self._mem)
        self.map_flags = MAP_NONE

    def map_read(self):
        if self.devmem is None:
            return
        if self.map_flags == MAP_NONE:
            self.device.download(self.devmem, self._mem)
            self.map_flags = MAP_READ

    def map_write(self):
        if self.devmem is None:
            return
        if self.map_flags == MAP_NONE:
            self.device.download(self.devmem, self._mem)
        self.map_flags = MAP_WRITE

    def map_invalidate(self):
        if self.devmem is None:
            return
        self.map_flags = MAP_INVALIDATE

    def unmap(self):
        if self.devmem is None:
            return
        if self.map_flags in (MAP_WRITE, MAP_INVALIDATE):
            self.device.upload(self.devmem, self._mem)
        self.map_flags = MAP_NONE

    # -- pickling ----------------------------------------------------------------------------
    def __getstate__(self):
        self.map_read()
        mem = self._mem
        if self.shallow_pickle and mem is not None:
            mem = numpy.zeros(mem.shape, mem.dtype)
        return {"mem": mem, "shallow_pickle": self.shallow_pickle,
                "max_supposed": self.max_supposed}

    def __setstate__(self, state):
        self._mem = state["mem"]
        self.shallow_pickle = state["shallow_pickle"]
        self.max_supposed = state.get("max_supposed", 1.0)
        self.device = None
        self.devmem = None
        self.map_flags = MAP_WRITE


Vector = Array
