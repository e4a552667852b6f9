"""Build the UNMODIFIED reference rasterizer for sm_100a into oracle/_ref/ (test infrastructure).

Only runs where /root/reference exists (the build container); the GPU box uses the prebuilt files
that travel with the repo snapshot (oracle/_ref/ is git-ignored but not gpurun-ignored).

Sources are compiled where they lie -- nothing is copied into this repository:
  DGR/cuda_rasterizer/{rasterizer_impl,forward,backward}.cu   the reference kernels
  DGR/rasterize_points.cu, DGR/ext.cpp                        the reference torch binding
with DGR = /root/reference/gaussian_splatting/submodules/diff-gaussian-rasterization.
The only flag beyond the reference's own setup.py (-I third_party/glm) is ``-include cstdint``:
rasterizer_impl.h:24,40-61 uses std::uintptr_t / uint32_t without including <cstdint>, which gcc 13
no longer provides transitively.  No source edits.

Outputs:
  oracle/_ref/libref_dgr.so   reference kernels + oracle/ref_shim.cu (C ABI, exposes intermediates)
  oracle/_ref/ref_dgr_C.so    the reference's own pybind module under the name ``ref_dgr_C``
                              (stock binding: used by bench.py --impl reference)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DGR = os.path.join(REF_ROOT, "gaussian_splatting", "submodules", "diff-gaussian-rasterization")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libref_dgr.so")
EXT = os.path.join(OUT, "ref_dgr_C.so")

NVCC = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-include", "cstdint",
        "-I", os.path.join(DGR, "third_party", "glm"), "-I", os.path.join(DGR, "cuda_rasterizer"), "-I", DGR, "-w"]
KERNEL_SRCS = [os.path.join(DGR, "cuda_rasterizer", f) for f in ("rasterizer_impl.cu", "forward.cu", "backward.cu")]


def available() -> bool:
    return os.path.isdir(DGR)


def _obj(src, tag=""):
    return os.path.join(OUT, tag + os.path.basename(src) + ".o")


def build(force=False, verbose=False, with_torch_ext=True):
    if not available():
        return False
    os.makedirs(OUT, exist_ok=True)
    run = lambda cmd: (print(" ".join(cmd), flush=True) if verbose else None, subprocess.check_call(cmd))
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "ref_shim.cu")):
        procs = []
        objs = []
        for s in KERNEL_SRCS + [os.path.join(HERE, "ref_shim.cu")]:
            o = _obj(s)
            objs.append(o)
            procs.append(subprocess.Popen(NVCC + ["-c", s, "-o", o]))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed on the reference sources")
        run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs])
    if with_torch_ext and (force or not os.path.exists(EXT)):
        import torch
        from torch.utils import cpp_extension as ce
        inc = []
        for p in ce.include_paths() + [sysconfig.get_paths()["include"]]:
            inc += ["-isystem", p]
        defs = ["-DTORCH_EXTENSION_NAME=ref_dgr_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
                f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
        o1 = _obj(os.path.join(DGR, "rasterize_points.cu"), "t_")
        run(NVCC + defs + inc + ["-c", os.path.join(DGR, "rasterize_points.cu"), "-o", o1])
        o2 = os.path.join(OUT, "t_ext.cpp.o")
        run(["g++", "-O2", "-std=c++17", "-fPIC", "-w", *defs, *inc, "-isystem", "/usr/local/cuda/include", "-I", DGR,
             "-c", os.path.join(DGR, "ext.cpp"), "-o", o2])
        tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
        run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", EXT, o1, o2, *[_obj(s) for s in KERNEL_SRCS],
             "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
             "-Xlinker", f"-rpath={tlib}"])
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("reference built" if ok else "reference sources not present; nothing built")
