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
    assert L.gstar_abi_version() == 5
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
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(gstar_fwd_args), offsetof(gstar_fwd_args, forward_only), offsetof(gstar_fwd_args, out_color),\n'
                   '       sizeof(gstar_bwd_args), offsetof(gstar_bwd_args, accumulate_param_grads), offsetof(gstar_bwd_args, blend_grad_scratch),\n'
                   '       sizeof(gstar_reblend_args), offsetof(gstar_reblend_args, out_color), offsetof(gstar_reblend_args, forward_only),\n'
                   '       offsetof(gstar_bwd_args, blend_only), offsetof(gstar_reblend_args, projmatrix));\n'
                   'return 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    F, B, Rb = capi.FwdArgs, capi.BwdArgs, capi.ReblendArgs
    want = [ctypes.sizeof(F), F.forward_only.offset, F.out_color.offset, ctypes.sizeof(B), B.accumulate_param_grads.offset,
            B.blend_grad_scratch.offset, ctypes.sizeof(Rb), Rb.out_color.offset, Rb.forward_only.offset, B.blend_only.offset,
            Rb.projmatrix.offset]
    assert got == want, (got, want)


def test_shared_geometry_host_logic():
    """Which calls may re-blend a remembered forward (gaustar_b200/rasterizer.py) -- the decision is pure host
    logic over (data_ptr, _version, shape, device, dtype) keys and is checked here without a GPU."""
    import torch
    from gaustar_b200 import rasterizer as R
    P = 10
    mk = lambda: (torch.rand(P, 3), torch.rand(P, 1), torch.rand(P, 3), torch.rand(P, 4), torch.Tensor([]), torch.eye(4), torch.eye(4))
    rs = R.GaussianRasterizationSettings(32, 48, 0.5, 0.6, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False)
    t = mk()
    k0 = R._geom_key(t, rs, "cpu")
    # the cache: identity only
    assert R._reblend_allowed(R._geom_key(t, rs, "cpu"), k0, False, True)
    assert not R._reblend_allowed(R._geom_key(t, rs, "cpu"), k0, False, False)  # cache off
    clone = tuple(x.clone() for x in t)
    assert not R._reblend_allowed(R._geom_key(clone, rs, "cpu"), k0, False, True)  # equal values, other memory
    t[0].add_(1.0)  # an optimizer step: same memory, new version
    assert not R._reblend_allowed(R._geom_key(t, rs, "cpu"), k0, False, True)
    # inside shared_geometry(): the caller vouches for the values, the configuration must still agree
    assert R._reblend_allowed(R._geom_key(clone, rs, "cpu"), k0, True, False)
    assert not R._reblend_allowed(R._geom_key(clone, rs._replace(image_width=64), "cpu"), k0, True, False)
    assert not R._reblend_allowed(R._geom_key(clone, rs._replace(scale_modifier=0.5), "cpu"), k0, True, False)
    fewer = tuple(x[:-1].clone() if x.dim() == 2 and x.shape[0] == P else x for x in clone)
    assert not R._reblend_allowed(R._geom_key(fewer, rs, "cpu"), k0, True, False)
    cov = clone[:2] + (torch.Tensor([]), torch.Tensor([]), torch.rand(P, 6)) + clone[5:]  # cov3D_precomp instead of scales/rotations
    assert not R._reblend_allowed(R._geom_key(cov, rs, "cpu"), k0, True, False)
    # the block clears what it remembered on both ends, and nests
    R._tls.src = "stale"
    with R.shared_geometry():
        assert R._tls.src is None and R._tls.scope == {"check": False}
        with R.shared_geometry(check=True):
            assert R._tls.scope == {"check": True}
        assert R._tls.scope == {"check": False}
        R._tls.src = "inside"
    assert R._tls.src is None and R._tls.scope is None
    assert R.set_geometry_cache(True) in (False, True) and R.set_geometry_cache(False) is True


def test_recent_calls_replacement_is_least_recently_used(tmp_path):
    """gaustar_b200/csrc/recent_calls.h (host-side memory of the last calls, what gstar_raster_reblend looks its source up
    in), exercised through a g++ harness: a refreshed or looked-up entry survives the calls that follow -- the scenario
    a FIFO cursor got wrong (the source of a multi-pass call evicted by its own first re-blend) -- and the least recently
    used entry is the one replaced."""
    import subprocess
    src = tmp_path / "rc.cpp"
    src.write_text(r'''
#include <stdio.h>
#include "recent_calls.h"
static const char* A(int i) { return (const char*)(size_t)(0x1000 * (i + 1)); }
int main() {
    RecentCalls rc;
    if (rc.find(A(0)) || rc.find(nullptr)) return 1;                 // empty: nothing is found, not even NULL
    for (int i = 0; i < RecentCalls::N; i++) rc.remember(A(i), 100 + i, 0, i, 7, 8, 9);
    // the oldest address comes back from the allocator (a new forward at A(0)) ...
    rc.remember(A(0), 555, 66, 77, 1, 2, 3);
    // ... and the calls that follow (its re-blends, at fresh addresses) must not evict it
    rc.remember(A(100), 1, 0, 0, 1, 2, 3);
    const RecentCalls::Entry* s = rc.find(A(0));
    if (!s || s->cap != 555 || s->log_slots != 66 || s->R != 77 || s->P != 1 || s->W != 2 || s->H != 3) return 2;
    if (rc.find(A(1))) return 3;                                      // the least recently used one went instead
    // a source that is looked up again and again outlives any number of newer calls
    for (int i = 0; i < 5 * RecentCalls::N; i++) {
        rc.remember(A(200 + i), 1, 0, 0, 1, 2, 3);
        if (!rc.find(A(0))) return 4;
    }
    // without look-ups an entry goes after N newer ones
    for (int i = 0; i < RecentCalls::N; i++) rc.remember(A(1000 + i), 1, 0, 0, 1, 2, 3);
    if (rc.find(A(0))) return 5;
    int live = 0;
    for (int i = 0; i < RecentCalls::N; i++) live += rc.find(A(1000 + i)) != nullptr;
    if (live != RecentCalls::N) return 6;
    printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "rc"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "gaustar_b200", "csrc"), str(src), "-o", str(exe)])
    assert subprocess.check_output([str(exe)]).strip() == b"ok"
