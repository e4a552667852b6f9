"""Developer tool: build a kernel-variant copy of libgstar_raster.so with extra -D flags (for A/B timing on the GPU box).

    python tools/build_variant.py <name> -DGSTAR_FWD_NB=64 -DGSTAR_FWD_CTAS=4 ...
    GSTAR_LIB_PATH=gaustar_b200/lib/variants/<name>.so python tools/stage_times.py
"""
import importlib.util, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_gstar_build", os.path.join(ROOT, "gaustar_b200", "build.py"))
B = importlib.util.module_from_spec(spec); spec.loader.exec_module(B)
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.LIBDIR, "variants"); os.makedirs(out_dir, exist_ok=True)
objs, procs = [], []
for s in B.CU_SOURCES:
    o = os.path.join(out_dir, f"{name}.{s}.o"); objs.append(o)
    procs.append(subprocess.Popen(["nvcc", *B.NVCC_FLAGS, *flags, "-I", os.path.join(ROOT, "include"), "-c", os.path.join(B.CSRC, s), "-o", o]))
assert all(p.wait() == 0 for p in procs)
subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", os.path.join(out_dir, name + ".so"), *objs])
for o in objs: os.remove(o)
print("built", os.path.join(out_dir, name + ".so"))
