"""ctypes front-end of oracle/_ref/libref_dgr.so -- the UNMODIFIED reference CUDA rasterizer.

TEST INFRASTRUCTURE ONLY (GPU box).  Built by oracle/build_ref.py from /root/reference; used by the
``-m gpu`` parity tests, by tests/golden/make_golden.py (to generate the committed golden vectors)
and by bench.py --impl reference (through the stock pybind module ref_dgr_C.so).
Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
LIB_PATH = os.path.join(REF_DIR, "libref_dgr.so")
EXT_PATH = os.path.join(REF_DIR, "ref_dgr_C.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
    return _lib


def stock_module():
    """The reference's own pybind module (DGR/ext.cpp) built under the name ref_dgr_C."""
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import ref_dgr_C  # noqa
    return ref_dgr_C


def _p(t):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _f(t, dev):
    return None if t is None else t.to(device=dev, dtype=torch.float32).contiguous()


def forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, W, H, shs=None, colors_precomp=None, scales=None,
            rotations=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=0, intermediates=True):
    dev = means3D.device
    torch.cuda.synchronize(dev)
    m3, op, vm, pm, cp, bgc, sh, col, sc, rot, cov = [_f(x, dev) for x in (means3D, opacities, viewmatrix, projmatrix, campos, bg, shs,
                                                                            colors_precomp, scales, rotations, cov3D_precomp)]
    P = m3.shape[0]
    M = 0 if sh is None or sh.numel() == 0 else sh.shape[1]
    out_color = torch.zeros(3, H, W, device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    L = lib()
    with torch.cuda.device(dev):
        R = L.ref_forward(P, sh_degree, M, _p(bgc), W, H, _p(m3), _p(sh), _p(col), _p(op), _p(sc), C.c_float(scale_modifier), _p(rot), _p(cov),
                          _p(vm), _p(pm), _p(cp), C.c_float(tan_fovx), C.c_float(tan_fovy), 0, _p(out_color), _p(radii), 0)
    if R < 0:
        raise RuntimeError(f"reference forward failed ({R})")
    out = dict(num_rendered=R, out_color=out_color, radii=radii)
    if intermediates:
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)
        g = dict(depths=z(P), means2D=z(P, 2), cov3D=z(P, 6), conic_opacity=z(P, 4), rgb=z(P, 3), tiles_touched=z(P, dt=torch.int32),
                 point_offsets=z(P, dt=torch.int32), clamped=z(P, 3, dt=torch.uint8))
        L.ref_get_geometry(_p(g["depths"]), _p(g["means2D"]), _p(g["cov3D"]), _p(g["conic_opacity"]), _p(g["rgb"]), _p(g["tiles_touched"]),
                           _p(g["point_offsets"]), _p(g["clamped"]))
        b = dict(keys_unsorted=z(R, dt=torch.int64), values_unsorted=z(R, dt=torch.int32), keys_sorted=z(R, dt=torch.int64),
                 point_list=z(R, dt=torch.int32))
        if R > 0:
            L.ref_get_binning(_p(b["keys_unsorted"]), _p(b["values_unsorted"]), _p(b["keys_sorted"]), _p(b["point_list"]))
        T = ((W + 15) // 16) * ((H + 15) // 16)
        i = dict(final_T=z(H * W), n_contrib=z(H * W, dt=torch.int32), ranges=z(T, 2, dt=torch.int32))
        L.ref_get_image(_p(i["final_T"]), _p(i["n_contrib"]), _p(i["ranges"]))
        out.update(g); out.update(b); out.update(i)
    return out


def backward(fwd, dL_dout_color, means3D, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, shs=None, colors_precomp=None, scales=None,
             rotations=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=0):
    """Must directly follow the matching forward() (the shim keeps the opaque buffers)."""
    dev = means3D.device
    m3, vm, pm, cp, bgc, sh, col, sc, rot, cov, dpix = [_f(x, dev) for x in (means3D, viewmatrix, projmatrix, campos, bg, shs, colors_precomp,
                                                                             scales, rotations, cov3D_precomp, dL_dout_color)]
    P = m3.shape[0]
    M = 0 if sh is None or sh.numel() == 0 else sh.shape[1]
    H, W = dpix.shape[1], dpix.shape[2]
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    g = dict(dL_dmeans2D=z(P, 3), dL_dconic=z(P, 4), dL_dopacity=z(P, 1), dL_dcolors=z(P, 3), dL_dmeans3D=z(P, 3), dL_dcov3D=z(P, 6),
             dL_dsh=z(P, M, 3), dL_dscales=z(P, 3), dL_drotations=z(P, 4))
    torch.cuda.synchronize(dev)
    with torch.cuda.device(dev):
        rc = lib().ref_backward(P, sh_degree, M, int(fwd["num_rendered"]), _p(bgc), W, H, _p(m3), _p(sh), _p(col), _p(sc), C.c_float(scale_modifier),
                                _p(rot), _p(cov), _p(vm), _p(pm), _p(cp), C.c_float(tan_fovx), C.c_float(tan_fovy), _p(fwd["radii"]), _p(dpix),
                                _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dopacity"]), _p(g["dL_dcolors"]), _p(g["dL_dmeans3D"]),
                                _p(g["dL_dcov3D"]), _p(g["dL_dsh"]), _p(g["dL_dscales"]), _p(g["dL_drotations"]), 0)
    if rc < 0:
        raise RuntimeError(f"reference backward failed ({rc})")
    return g


# ---- the reference simple_knn (oracle/_ref/libref_knn.so; oracle/ref_knn_shim.cu) ----
KNN_LIB_PATH = os.path.join(REF_DIR, "libref_knn.so")
_knn_lib = None


def knn_available() -> bool:
    return os.path.exists(KNN_LIB_PATH)


def knn_mean_dist2(points: torch.Tensor) -> torch.Tensor:
    """SimpleKNN::knn of the unmodified reference on a [P,3] fp32 CUDA tensor: mean squared distance to the 3 nearest other points."""
    global _knn_lib
    if _knn_lib is None:
        _knn_lib = C.CDLL(KNN_LIB_PATH)
        _knn_lib.ref_knn_mean_dist2.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _knn_lib.ref_knn_mean_dist2.restype = C.c_int
    pts = points.detach().to(dtype=torch.float32).contiguous()
    out = torch.zeros(pts.shape[0], device=pts.device)
    torch.cuda.synchronize(pts.device)
    rc = _knn_lib.ref_knn_mean_dist2(pts.shape[0], C.c_void_p(pts.data_ptr()), C.c_void_p(out.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"reference simple_knn failed: CUDA error {rc}")
    return out
