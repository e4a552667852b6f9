"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/gstar_raster.h; the Python surface matches the reference wrapper
(DGR/diff_gaussian_rasterization/__init__.py:157-220); there is no CPU fallback."""
import os
import re

import pytest
import torch

from gaustar_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gstar_raster.h")).read()
    return sorted(set(re.findall(r"GSTAR_API\s+[\w\s\*]+?\b(gstar_\w+)\s*\(", src)))


def test_header_symbols_all_exported():
    L = capi.lib()
    names = declared_symbols()
    assert len(names) >= 13, names
    for n in names:
        assert hasattr(L, n), f"{n} declared in gstar_raster.h but not exported by libgstar_raster.so"
    assert sorted(capi.EXPORTED) == names


def test_abi_version_and_stage_names():
    L = capi.lib()
    assert L.gstar_abi_version() == 4
    assert [L.gstar_stage_name(i).decode() for i in range(len(capi.STAGES))] == capi.STAGES
    assert L.gstar_stage_name(99) == b""


def test_buffer_sizes_monotone():
    L = capi.lib()
    assert L.gstar_geom_bytes(0) >= 128
    assert L.gstar_geom_bytes(1000) >= 1000 * 64
    assert L.gstar_geom_bytes(2000) > L.gstar_geom_bytes(1000)
    a, b = L.gstar_image_bytes(64, 64), L.gstar_image_bytes(1920, 1080)
    assert b > a >= 64 * 64 * 8
    assert L.gstar_binning_bytes(10) >= 120
    assert L.gstar_binning_bytes(1 << 20) >= 12 << 20


def test_null_arguments_are_errors_not_crashes():
    L = capi.lib()
    assert L.gstar_raster_backward(None, None) < 0
    assert b"null" in L.gstar_last_error()
    assert L.gstar_mark_visible(0, None, None, None, None, None) == 0
    assert L.gstar_mark_visible(5, None, None, None, None, None) < 0


def test_dropin_python_surface():
    import diff_gaussian_rasterization as d
    assert d.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix", "sh_degree", "campos",
        "prefiltered", "debug")
    for n in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert hasattr(d._C, n)
    rs = d.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    r = d.GaussianRasterizer(rs)
    m = torch.zeros(4, 3)
    # DGR/__init__.py:191-195: exactly one of (shs | colors_precomp), (scales+rotations | cov3D_precomp)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(m, m, torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4), cov3D_precomp=torch.zeros(4, 6))


def test_no_cpu_fallback():
    """CPU tensors must be refused loudly (the product has no CPU path)."""
    import diff_gaussian_rasterization as d
    rs = d.GaussianRasterizationSettings(8, 8, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="no CPU path"):
        d.GaussianRasterizer(rs)(m, m, torch.zeros(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="means3D must have dimensions"):
        d._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), m, torch.zeros(4, 1), m, torch.zeros(4, 4), 1.0, torch.Tensor([]),
                                 torch.eye(4), torch.eye(4), 1.0, 1.0, 8, 8, torch.Tensor([]), 0, torch.zeros(3), False, False)


def test_product_does_not_import_oracle():
    """The product package must never reference oracle/ (checked textually over its sources)."""
    pkg = os.path.join(ROOT, "gaustar_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("# oracle/", ""), f


def test_ctypes_structs_match_the_header(tmp_path):
    """The ctypes mirror of gstar_fwd_args / gstar_bwd_args / gstar_reblend_args (gaustar_b200/capi.py) must have the C compiler's layout of
    include/gstar_raster.h: sizes and the offsets of the trailing fields (the ones that were appended over time)."""
    import ctypes
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gstar_raster.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gstar_fwd_args), offsetof(gstar_fwd_args, forward_only), offsetof(gstar_fwd_args, out_color),\n'
                   '       sizeof(gstar_bwd_args), offsetof(gstar_bwd_args, accumulate_param_grads), offsetof(gstar_bwd_args, blend_grad_scratch),\n'
                   '       sizeof(gstar_reblend_args), offsetof(gstar_reblend_args, out_color), offsetof(gstar_reblend_args, forward_only));\n'
                   'return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    F, B, Rb = capi.FwdArgs, capi.BwdArgs, capi.ReblendArgs
    want = [ctypes.sizeof(F), F.forward_only.offset, F.out_color.offset, ctypes.sizeof(B), B.accumulate_param_grads.offset,
            B.blend_grad_scratch.offset, ctypes.sizeof(Rb), Rb.out_color.offset, Rb.forward_only.offset]
    assert got == want, (got, want)
