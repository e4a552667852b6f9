"""In-tree build of the native code (sm_100a only).

  gaustar_b200/lib/libgstar_raster.so   CUDA kernels + C ABI (include/gstar_raster.h); no torch
  gaustar_b200/_C*.so                   torch shim exposing the reference's three pybind functions

Both are plain nvcc / g++ invocations (no JIT cache): the built files travel with the repo snapshot
to the GPU box.  ``python gaustar_b200/build.py`` builds everything (run as a script: importing the package needs the
built extension).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgstar_raster.so")
EXT = os.path.join(HERE, "_C.so")

CU_SOURCES = ["preprocess.cu", "binning.cu", "blend.cu", "api.cu", "knn.cu", "sugar_prologue.cu"]
CU_HEADERS = ["gstar_common.cuh", "gstar_kernels.h", "recent_calls.h", os.path.join(ROOT, "include", "gstar_raster.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "--extended-lambda"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_lib(force=False, verbose=False, extra_flags=()):
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in CU_HEADERS]
    if not (force or _newer(LIB, deps)):
        return LIB
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(LIBDIR, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["nvcc", *NVCC_FLAGS, *extra_flags, "-I", os.path.join(ROOT, "include"), "-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    for pr in procs:
        if pr.wait() != 0:
            raise RuntimeError("nvcc failed")
    _run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs], verbose)
    return LIB


def build_ext(force=False, verbose=False):
    import torch
    from torch.utils import cpp_extension as ce

    src = os.path.join(CSRC, "torch_ext.cpp")
    if not (force or _newer(EXT, [src, os.path.join(ROOT, "include", "gstar_raster.h"), LIB])):
        return EXT
    inc = []
    for p in ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include"]:
        inc += ["-isystem", p]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *inc, src, "-o", EXT,
           "-L", LIBDIR, "-lgstar_raster", "-Wl,-rpath,$ORIGIN/lib",
           "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", f"-Wl,-rpath,{tlib}"]
    _run(cmd, verbose)
    return EXT


def build_all(force=False, verbose=False):
    build_lib(force, verbose)
    build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
