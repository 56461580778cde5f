import numpy

dtypes = {"float": numpy.float32, "double": numpy.float64}
cl_defines = {"float": {"dtype": "float"}, "double": {"dtype": "double"}}
itypes = {"float": numpy.int32, "double": numpy.int64}


def numpy_dtype_to_opencl(dtype):
    dtype = numpy.dtype(dtype)
    return {numpy.dtype(numpy.float32): "float", numpy.dtype(numpy.float64): "double",
            numpy.dtype(numpy.int32): "int", numpy.dtype(numpy.int64): "long",
            numpy.dtype(numpy.uint8): "uchar", numpy.dtype(numpy.int8): "char",
            numpy.dtype(numpy.int16): "short", numpy.dtype(numpy.uint16): "ushort",
            numpy.dtype(numpy.uint32): "uint", numpy.dtype(numpy.uint64): "ulong"}[dtype]
