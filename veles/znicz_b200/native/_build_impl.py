"""Builds libznicz_native.so, znicz_native_test and znicz_infer next to this file."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "src")
KBUILD = os.path.join(os.path.dirname(HERE), "kernels", "build")
LIB = os.path.join(HERE, "libznicz_native.so")
TEST_BIN = os.path.join(HERE, "znicz_native_test")
INFER_BIN = os.path.join(HERE, "znicz_infer")
KERNEL_OBJS = ("gemm_simt.o", "pooling.o", "elementwise.o", "softmax_eval.o",
               # tensor-core path of the CUDA executor (split-bf16 operands)
               "gemm_umma.o", "gemm_pair.o", "split.o")


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed:\n  %s\n%s" % (" ".join(cmd), r.stdout))


def build(verbose=True, force=False):
    from ..kernels import build as kbuild
    kbuild.build(verbose=False)          # the CUDA executor links the training kernels' objects
    cuda_home = os.path.dirname(os.path.dirname(kbuild._nvcc()))
    srcs = [os.path.join(SRC, f) for f in ("package.cc", "engine.cc", "engine_cuda.cc")]
    objs = [os.path.join(KBUILD, o) for o in KERNEL_OBJS]
    newest = max(os.path.getmtime(p) for p in srcs + objs + [os.path.join(SRC, "znicz_native.h")])
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest and \
            os.path.exists(TEST_BIN) and os.path.exists(INFER_BIN):
        return LIB
    inc = ["-I", os.path.join(cuda_home, "include")]
    libs = ["-L", os.path.join(cuda_home, "lib64"), "-Wl,-rpath," + os.path.join(cuda_home, "lib64"),
            "-lcudart", "-lz"]
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared"] + inc + srcs + objs + libs + ["-o", LIB])
    rpath = ["-Wl,-rpath," + HERE, "-L", HERE, "-lznicz_native"]
    _run(["g++", "-O2", "-std=c++17"] + inc + [os.path.join(HERE, "tests", "test_native.cc")] +
         rpath + libs + ["-o", TEST_BIN])
    _run(["g++", "-O2", "-std=c++17"] + inc + [os.path.join(SRC, "infer_main.cc")] + rpath + libs +
         ["-o", INFER_BIN])
    if verbose:
        print("znicz_b200 native runtime: built %s" % LIB)
    return LIB
