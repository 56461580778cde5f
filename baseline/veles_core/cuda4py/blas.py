"""cuda4py.blas surface: cuBLAS S/DGEMM through ctypes (column-major, like the library)."""
import ctypes
import os

import numpy

CUBLAS_OP_N = 0
CUBLAS_OP_T = 1
CUBLAS_OP_C = 2

_lib = None
from cuda4py import dry_stats as _stats  # noqa: E402


def _load():
    global _lib
    if _lib is not None:
        return _lib
    # libcublas and libcublasLt must come from the SAME directory: with LD_LIBRARY_PATH pointing
    # at the toolkit, the wheel's libcublas.so.12 (12.8) otherwise binds to the toolkit's
    # libcublasLt.so.12 (12.9) and SGEMM fails with CUBLAS_STATUS_INVALID_VALUE for valid
    # arguments ("non-default emulation strategy ..." in the cublasLt log). Load Lt first, by path.
    dirs = []
    try:
        import nvidia.cublas.lib as _nl          # the wheel torch depends on
        dirs.append(os.path.dirname(_nl.__file__) if getattr(_nl, "__file__", None)
                    else list(_nl.__path__)[0])
    except Exception:
        pass
    dirs += ["/usr/local/cuda/lib64", ""]
    err = None
    for d in dirs:
        try:
            ctypes.CDLL(os.path.join(d, "libcublasLt.so.12"), mode=ctypes.RTLD_GLOBAL)
            _lib = ctypes.CDLL(os.path.join(d, "libcublas.so.12"), mode=ctypes.RTLD_GLOBAL)
            break
        except OSError as e:
            err = e
    if _lib is None:
        raise OSError("cannot load libcublas: %s" % err)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    for fn in ("cublasSgemm_v2", "cublasDgemm_v2"):
        f = getattr(_lib, fn)
        f.restype = ci
        f.argtypes = [vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, ci, vp, vp, ci]
    _lib.cublasCreate_v2.restype = ci
    _lib.cublasCreate_v2.argtypes = [ctypes.POINTER(vp)]
    _lib.cublasDestroy_v2.restype = ci
    _lib.cublasDestroy_v2.argtypes = [vp]
    return _lib


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "devmem"):
        x = x.devmem
    return ctypes.c_void_p(int(x))


class CUBLAS(object):
    def __init__(self, context):
        self.context = context
        import cuda4py
        self._dry = cuda4py.DRY
        if self._dry:
            self.handle = self._lib = None
            return
        context.set_current()
        lib = _load()
        h = ctypes.c_void_p()
        st = lib.cublasCreate_v2(ctypes.byref(h))
        if st != 0:
            raise RuntimeError("cublasCreate failed: %d" % st)
        self.handle = h
        self._lib = lib

    @staticmethod
    def gemm(dtype):
        dtype = numpy.dtype(dtype)
        if dtype == numpy.float32:
            return CUBLAS.sgemm
        if dtype == numpy.float64:
            return CUBLAS.dgemm
        raise ValueError("unsupported GEMM dtype %s" % dtype)

    def _gemm(self, fn, ctype, transA, transB, rowsCountA, columnCountB, commonSideLength,
              alpha, A, B, beta, C, strideA, strideB, strideC):
        if self._dry:
            import cuda4py
            cuda4py.dry_stats["gemms"] += 1
            assert rowsCountA > 0 and columnCountB > 0 and commonSideLength > 0
            assert A is not None and B is not None and C is not None
            return
        if not strideA:
            strideA = commonSideLength if transA != CUBLAS_OP_N else rowsCountA
        if not strideB:
            strideB = columnCountB if transB != CUBLAS_OP_N else commonSideLength
        if not strideC:
            strideC = rowsCountA
        _stats["gemms"] += 1
        a = ctype(float(numpy.asarray(alpha).ravel()[0]))
        b = ctype(float(numpy.asarray(beta).ravel()[0]))
        st = fn(self.handle, transA, transB, rowsCountA, columnCountB, commonSideLength,
                ctypes.byref(a), _ptr(A), strideA, _ptr(B), strideB, ctypes.byref(b), _ptr(C),
                strideC)
        if st != 0:
            raise RuntimeError("cuBLAS gemm failed: status %d (transA %s transB %s m %s n %s k %s "
                               "lda %s ldb %s ldc %s A %s B %s C %s)" % (
                                   st, transA, transB, rowsCountA, columnCountB, commonSideLength,
                                   strideA, strideB, strideC, _ptr(A), _ptr(B), _ptr(C)))

    def sgemm(self, transA, transB, rowsCountA, columnCountB, commonSideLength, alpha, A, B,
              beta, C, strideA=0, strideB=0, strideC=0):
        self._gemm(self._lib and self._lib.cublasSgemm_v2, ctypes.c_float, transA, transB, rowsCountA,
                   columnCountB, commonSideLength, alpha, A, B, beta, C, strideA, strideB, strideC)

    def dgemm(self, transA, transB, rowsCountA, columnCountB, commonSideLength, alpha, A, B,
              beta, C, strideA=0, strideB=0, strideC=0):
        self._gemm(self._lib and self._lib.cublasDgemm_v2, ctypes.c_double, transA, transB, rowsCountA,
                   columnCountB, commonSideLength, alpha, A, B, beta, C, strideA, strideB, strideC)

    def __del__(self):
        try:
            self._lib.cublasDestroy_v2(self.handle)
        except Exception:
            pass
