import time


def timeit(fn, *args, **kwargs):
    t0 = time.time()
    res = fn(*args, **kwargs)
    return res, time.time() - t0
