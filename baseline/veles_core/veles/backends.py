"""``veles.backends``: NumpyDevice and CUDADevice (context + cuBLAS handle + temp buffer)."""
import os

import numpy


class Device(object):
    backend_name = None
    exists = False

    def __init__(self, *args, **kwargs):
        self._temp_request = 0
        self._temp = None

    @property
    def is_async(self):
        return False

    def sync(self):
        pass

    def assign_backend_methods(self, obj, backend_methods):
        for suffix in backend_methods:
            name = "%s_%s" % (self.backend_name, suffix)
            fn = getattr(obj, name, None)
            if fn is None and suffix == "init":
                fn = _nothing
            if fn is None:
                raise NotImplementedError("%s lacks %s()" % (obj, name))
            setattr(obj, "_backend_%s_" % suffix, fn)

    def request_temp_buffer(self, size):
        self._temp_request = max(self._temp_request, int(size))

    def get_temp_buffer(self):
        return None


def _nothing(*args, **kwargs):
    return None


class NumpyDevice(Device):
    backend_name = "numpy"
    exists = False


class _Skip(object):
    def __init__(self, n):
        self.n = n


class CUDADevice(Device):
    backend_name = "cuda"
    exists = True

    def __init__(self, index=None, **kwargs):
        super(CUDADevice, self).__init__()
        import cuda4py as cu
        import cuda4py.blas as cublas
        if index is None:
            index = int(os.environ.get("LOCAL_RANK", "0"))
        self._cu = cu
        self.devices = cu.Devices()
        self.device_info = self.devices[index]
        self.index = index
        self.context = cu.Context(self.device_info)
        self.blas = cublas.CUBLAS(self.context)
        self.pinned = bool(kwargs.get("pinned", False))
        self._registered = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @property
    def is_async(self):
        return True

    def skip(self, n=1):
        return self._cu.skip(n)

    def sync(self):
        self.context.synchronize()

    def allocate(self, host_array):
        if self.pinned:
            self.pin(host_array)
        return self.context.mem_alloc(host_array.nbytes)

    def pin(self, host_array):
        """Page-lock the host side of an Array (cuMemHostRegister) so that its per-step
        uploads / read-backs are genuine pinned-memory transfers."""
        if self._cu.DRY:
            return
        from cuda.bindings import driver
        addr = host_array.ctypes.data
        if addr in self._registered or host_array.nbytes == 0:
            return
        res = driver.cuMemHostRegister(addr, host_array.nbytes, 0)
        if int(res[0]) == 0:
            self._registered[addr] = host_array.nbytes

    def upload(self, devmem, host_array):
        devmem.to_device(host_array)
        self.h2d_bytes += host_array.nbytes

    def download(self, devmem, host_array):
        devmem.to_host(host_array)
        self.d2h_bytes += host_array.nbytes

    def suggest_block_size(self, krn):
        return krn.max_potential_block_size()[1]

    def get_temp_buffer(self):
        if self._temp is None or self._temp.size < self._temp_request:
            self._temp = self.context.mem_alloc(max(self._temp_request, 4))
        return self._temp

    @property
    def compute_capability(self):
        return self.device_info.compute_capability


class Device_auto(object):
    pass


def get_device(backend=None):
    from veles.config import root
    backend = backend or root.common.engine.backend
    if backend in ("cuda", "auto"):
        return CUDADevice()
    return NumpyDevice()
