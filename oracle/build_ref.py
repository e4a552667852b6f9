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
  oracle/_ref/libref_knn.so   the reference simple_knn (gaussian_splatting/submodules/simple-knn/simple_knn.cu, unmodified) +
                              oracle/ref_knn_shim.cu: pins the simple_knn._C.distCUDA2 shim (tests/test_simple_knn.py)
  oracle/_ref/pyref/...       the reference's Python CALLERS of the operator, byte-compiled where they lie (sourceless
                              byte-code files *.refpyc, the Python analogue of the .so above; imported through oracle/pyref.py): ``gaussian_renderer.render()``, the
                              ``GaussianModel`` it renders, the stock operator wrapper, and the SuGaR model/camera modules.
                              tests/test_zz_reference_callers_gpu.py runs them unchanged against this repository's op.
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
SKNN = os.path.join(REF_ROOT, "gaussian_splatting", "submodules", "simple-knn")
KNN_LIB = os.path.join(OUT, "libref_knn.so")

NVCC = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-include", "cstdint",
        "-I", os.path.join(DGR, "third_party", "glm"), "-I", os.path.join(DGR, "cuda_rasterizer"), "-I", DGR, "-w"]
KERNEL_SRCS = [os.path.join(DGR, "cuda_rasterizer", f) for f in ("rasterizer_impl.cu", "forward.cu", "backward.cu")]


PYREF = os.path.join(OUT, "pyref")
PYC_SUFFIX = ".refpyc"  # (not ".pyc": gpurun's snapshot drops *.pyc; oracle/pyref.py imports these)
GS = os.path.join(REF_ROOT, "gaussian_splatting")
# (source file, path below oracle/_ref/pyref/ without the .pyc suffix)
PY_CALLERS = [
    (os.path.join(GS, "gaussian_renderer", "__init__.py"), "gaussian_splatting/gaussian_renderer/__init__"),
    (os.path.join(GS, "scene", "gaussian_model.py"), "gaussian_splatting/scene/gaussian_model"),
    (os.path.join(GS, "utils", "general_utils.py"), "gaussian_splatting/utils/general_utils"),
    (os.path.join(GS, "utils", "sh_utils.py"), "gaussian_splatting/utils/sh_utils"),
    (os.path.join(GS, "utils", "graphics_utils.py"), "gaussian_splatting/utils/graphics_utils"),
    (os.path.join(GS, "utils", "system_utils.py"), "gaussian_splatting/utils/system_utils"),
    (os.path.join(DGR, "diff_gaussian_rasterization", "__init__.py"), "ref_operator/diff_gaussian_rasterization/__init__"),
    (os.path.join(REF_ROOT, "gaustar_scene", "sugar_model.py"), "gaustar/gaustar_scene/sugar_model"),
    (os.path.join(REF_ROOT, "gaustar_scene", "cameras.py"), "gaustar/gaustar_scene/cameras"),
    (os.path.join(REF_ROOT, "gaustar_scene", "gs_model.py"), "gaustar/gaustar_scene/gs_model"),
    (os.path.join(REF_ROOT, "gaustar_scene", "sugar_optimizer.py"), "gaustar/gaustar_scene/sugar_optimizer"),
    (os.path.join(REF_ROOT, "gaustar_utils", "loss_utils.py"), "gaustar/gaustar_utils/loss_utils"),
    (os.path.join(REF_ROOT, "gaustar_utils", "spherical_harmonics.py"), "gaustar/gaustar_utils/spherical_harmonics"),
    (os.path.join(REF_ROOT, "gaustar_utils", "graphics_utils.py"), "gaustar/gaustar_utils/graphics_utils"),
    (os.path.join(REF_ROOT, "gaustar_utils", "general_utils.py"), "gaustar/gaustar_utils/general_utils"),
]


def available() -> bool:
    return os.path.isdir(DGR)


def build_py(force=False, verbose=False) -> bool:
    """Byte-compile the reference's Python callers into oracle/_ref/pyref (sourceless .pyc; nothing is copied as source)."""
    if not available():
        return False
    import py_compile
    for src, rel in PY_CALLERS:
        if not os.path.exists(src):
            continue
        dst = os.path.join(PYREF, rel + PYC_SUFFIX)
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=src, doraise=True)
        if verbose:
            print("byte-compiled", src, "->", dst, flush=True)
    return True


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
    knn_shim = os.path.join(HERE, "ref_knn_shim.cu")
    if os.path.isdir(SKNN) and (force or not os.path.exists(KNN_LIB) or os.path.getmtime(KNN_LIB) < os.path.getmtime(knn_shim)):
        # -include cfloat: simple_knn.cu uses FLT_MAX without including <cfloat> (the reference's setup.py builds it on older toolchains
        # where another header brought it in); no source edits
        run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-include", "cfloat", "-w",
             "-I", SKNN, "-shared", "-o", KNN_LIB, os.path.join(SKNN, "simple_knn.cu"), knn_shim])
    build_py(force, verbose)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv, verbose=True)
    print("reference built" if ok else "reference sources not present; nothing built")
