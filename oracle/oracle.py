"""numpy/ctypes front-end of the CPU oracle (oracle/gstar_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gstar_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the
product package ``gaustar_b200``.

The functions mirror the stages of the reference forward/backward
(DGR/cuda_rasterizer/rasterizer_impl.cu:198-336 and :340-434) and expose every
intermediate the parity tests compare.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgstar_oracle.so")
_SRC = os.path.join(_HERE, "gstar_oracle.c")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (no contraction, OpenMP)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


class _Scene(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("W", C.c_int), ("H", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("scale_modifier", C.c_float),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p), ("bg", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_scan.restype = C.c_int64
        _lib.orc_higher_msb.restype = C.c_uint32
    return _lib


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Inputs:
    """Arguments of _C.rasterize_gaussians (DGR/rasterize_points.h:18-38), host side."""
    means3D: np.ndarray
    opacities: np.ndarray
    viewmatrix: np.ndarray
    projmatrix: np.ndarray
    campos: np.ndarray
    bg: np.ndarray
    tan_fovx: float
    tan_fovy: float
    W: int
    H: int
    shs: Optional[np.ndarray] = None
    colors_precomp: Optional[np.ndarray] = None
    scales: Optional[np.ndarray] = None
    rotations: Optional[np.ndarray] = None
    cov3D_precomp: Optional[np.ndarray] = None
    scale_modifier: float = 1.0
    sh_degree: int = 0
    _keep: list = field(default_factory=list, repr=False)

    def c_scene(self) -> _Scene:
        self.means3D = _f32(self.means3D).reshape(-1, 3)
        self.opacities = _f32(self.opacities).reshape(-1)
        self.viewmatrix = _f32(self.viewmatrix).reshape(16)
        self.projmatrix = _f32(self.projmatrix).reshape(16)
        self.campos = _f32(self.campos).reshape(3)
        self.bg = _f32(self.bg).reshape(3)
        self.shs = _f32(self.shs)
        self.colors_precomp = _f32(self.colors_precomp)
        self.scales = _f32(self.scales)
        self.rotations = _f32(self.rotations)
        self.cov3D_precomp = _f32(self.cov3D_precomp)
        P = self.means3D.shape[0]
        M = 0 if self.shs is None else self.shs.shape[1]
        return _Scene(P, self.sh_degree, M, self.W, self.H, _p(self.means3D), _p(self.shs), _p(self.colors_precomp),
                      _p(self.opacities), _p(self.scales), _p(self.rotations), _p(self.cov3D_precomp),
                      self.scale_modifier, _p(self.viewmatrix), _p(self.projmatrix), _p(self.campos), _p(self.bg),
                      self.tan_fovx, self.tan_fovy)


@dataclass
class Forward:
    depths: np.ndarray
    radii: np.ndarray
    means2D: np.ndarray
    cov3D: np.ndarray
    conic_opacity: np.ndarray
    rgb: np.ndarray
    clamped: np.ndarray
    tiles_touched: np.ndarray
    point_offsets: np.ndarray
    num_rendered: int
    keys_unsorted: np.ndarray
    values_unsorted: np.ndarray
    keys_sorted: np.ndarray
    point_list: np.ndarray
    ranges: np.ndarray
    out_color: np.ndarray
    final_T: np.ndarray
    n_contrib: np.ndarray


def tile_grid(W, H):
    return (W + 15) // 16, (H + 15) // 16


def preprocess(inp: Inputs):
    s = inp.c_scene()
    P = s.P
    out = dict(
        depths=np.zeros(P, np.float32), radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32),
        cov3D=np.zeros((P, 6), np.float32), conic_opacity=np.zeros((P, 4), np.float32), rgb=np.zeros((P, 3), np.float32),
        clamped=np.zeros((P, 3), np.uint8), tiles_touched=np.zeros(P, np.uint32))
    lib().orc_preprocess(C.byref(s), *[_p(out[k]) for k in
                                       ("depths", "radii", "means2D", "cov3D", "conic_opacity", "rgb", "clamped", "tiles_touched")])
    return out


def forward(inp: Inputs, blend: bool = True) -> Forward:
    """Full reference forward, stage by stage (rasterizer_impl.cu:198-336)."""
    L = lib()
    pre = preprocess(inp)
    P = inp.means3D.shape[0]
    W, H = inp.W, inp.H
    gx, gy = tile_grid(W, H)
    offsets = np.zeros(P, np.uint32)
    R = int(L.orc_scan(P, _p(pre["tiles_touched"]), _p(offsets)))
    keys = np.zeros(R, np.uint64)
    vals = np.zeros(R, np.uint32)
    L.orc_duplicate_with_keys(P, W, H, _p(pre["means2D"]), _p(pre["depths"]), _p(offsets), _p(pre["radii"]), _p(keys), _p(vals))
    bit = int(L.orc_higher_msb(C.c_uint32(gx * gy)))
    keys_s = np.zeros(R, np.uint64)
    vals_s = np.zeros(R, np.uint32)
    L.orc_sort_pairs(C.c_int64(R), _p(keys), _p(vals), _p(keys_s), _p(vals_s), 32 + bit)
    ranges = np.zeros((gx * gy, 2), np.uint32)
    L.orc_tile_ranges(C.c_int64(R), _p(keys_s), gx * gy, _p(ranges))
    out_color = np.zeros((3, H, W), np.float32)
    final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32)
    if blend:
        colors = inp.colors_precomp if inp.colors_precomp is not None else pre["rgb"]
        colors = _f32(colors).reshape(-1, 3)
        L.orc_blend_forward(W, H, _p(ranges), _p(vals_s), _p(pre["means2D"]), _p(colors), _p(pre["conic_opacity"]), _p(inp.bg),
                            _p(out_color), _p(final_T), _p(n_contrib))
    return Forward(point_offsets=offsets, num_rendered=R, keys_unsorted=keys, values_unsorted=vals, keys_sorted=keys_s,
                   point_list=vals_s, ranges=ranges, out_color=out_color, final_T=final_T, n_contrib=n_contrib, **pre)


@dataclass
class Backward:
    dL_dmeans2D: np.ndarray
    dL_dconic: np.ndarray
    dL_dopacity: np.ndarray
    dL_dcolors: np.ndarray
    dL_dmeans3D: np.ndarray
    dL_dcov3D: np.ndarray
    dL_dsh: np.ndarray
    dL_dscales: np.ndarray
    dL_drotations: np.ndarray


def backward(inp: Inputs, fwd: Forward, dL_dout_color: np.ndarray) -> Backward:
    """Reference backward (rasterizer_impl.cu:340-434); blend sums in fp64."""
    L = lib()
    s = inp.c_scene()
    P, M = s.P, s.M
    W, H = inp.W, inp.H
    g = _f32(dL_dout_color).reshape(3, H, W)
    colors = _f32(inp.colors_precomp if inp.colors_precomp is not None else fwd.rgb).reshape(-1, 3)
    d_m2 = np.zeros((P, 3), np.float64)
    d_con = np.zeros((P, 4), np.float64)
    d_op = np.zeros(P, np.float64)
    d_col = np.zeros((P, 3), np.float64)
    L.orc_blend_backward(W, H, _p(fwd.ranges), _p(fwd.point_list), _p(inp.bg), _p(fwd.means2D), _p(fwd.conic_opacity), _p(colors),
                         _p(fwd.final_T), _p(fwd.n_contrib), _p(g), _p(d_m2), _p(d_con), _p(d_op), _p(d_col))
    m2f, conf, colf = d_m2.astype(np.float32), d_con.astype(np.float32), d_col.astype(np.float32)
    cov3D = _f32(inp.cov3D_precomp if inp.cov3D_precomp is not None else fwd.cov3D).reshape(-1, 6)
    d_mean3 = np.zeros((P, 3), np.float32)
    d_cov3 = np.zeros((P, 6), np.float32)
    d_sh = np.zeros((P, M, 3), np.float32)
    d_sc = np.zeros((P, 3), np.float32)
    d_rot = np.zeros((P, 4), np.float32)
    L.orc_preprocess_backward(C.byref(s), _p(fwd.radii), _p(cov3D), _p(fwd.clamped), _p(m2f), _p(conf), _p(colf), _p(d_mean3),
                              _p(d_cov3), _p(d_sh) if M else None, _p(d_sc) if inp.scales is not None else None,
                              _p(d_rot) if inp.scales is not None else None)
    return Backward(dL_dmeans2D=m2f, dL_dconic=conf, dL_dopacity=d_op.astype(np.float32).reshape(P, 1), dL_dcolors=colf,
                    dL_dmeans3D=d_mean3, dL_dcov3D=d_cov3, dL_dsh=d_sh, dL_dscales=d_sc, dL_drotations=d_rot)


def mark_visible(means3D, viewmatrix):
    m = _f32(means3D).reshape(-1, 3)
    v = _f32(viewmatrix).reshape(16)
    out = np.zeros(m.shape[0], np.uint8)
    lib().orc_mark_visible(m.shape[0], _p(m), _p(v), _p(out))
    return out.astype(bool)
