"""ctypes binding of the C ABI (include/gstar_raster.h) with torch tensors as device memory.

This is the "reference-side binding" in Python form: what a maintainer of a Python host would write
to call libgstar_raster.so directly (see INTEGRATION.md for the C++/pybind form that replaces
DGR/rasterize_points.cu).  The parity tests and bench.py go through this module so that they
exercise exactly the exported C symbols.  No compute happens in Python.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSTAR_LIB_PATH") or os.path.join(_HERE, "lib", "libgstar_raster.so")  # override: kernel-variant experiments

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

# every symbol declared in include/gstar_raster.h
EXPORTED = [
    "gstar_raster_forward", "gstar_raster_backward", "gstar_mark_visible", "gstar_last_error", "gstar_abi_version",
    "gstar_geom_bytes", "gstar_image_bytes", "gstar_binning_bytes", "gstar_geom_unpack", "gstar_image_views",
    "gstar_binning_views", "gstar_profile_stage", "gstar_stage_name", "gstar_set_hit_log", "gstar_set_deterministic", "gstar_hit_log_state", "gstar_debug_header", "gstar_knn3_mean_dist2",
    "gstar_raster_reblend", "gstar_sugar_prologue_forward", "gstar_sugar_prologue_backward",
]
STAGES = ["preprocess_fwd", "tile_scan", "emit", "tile_sort", "blend_fwd", "blend_bwd", "preprocess_bwd"]


class FwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int),
        ("background", C.c_void_p), ("width", C.c_int), ("height", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("scales", C.c_void_p), ("scale_modifier", C.c_float), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("cam_pos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float), ("prefiltered", C.c_int),
        ("out_color", C.c_void_p), ("radii", C.c_void_p), ("debug", C.c_int), ("forward_only", C.c_int),
        ("colors2", C.c_void_p), ("background2", C.c_void_p), ("out_color2", C.c_void_p), ("channels2", C.c_int),
    ]


class BwdArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("R", C.c_int),
        ("background", C.c_void_p), ("width", C.c_int), ("height", C.c_int),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("scales", C.c_void_p),
        ("scale_modifier", C.c_float), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", C.c_void_p), ("geom_buffer", C.c_void_p), ("binning_buffer", C.c_void_p), ("image_buffer", C.c_void_p),
        ("dL_dpix", C.c_void_p),
        ("dL_dmean2D", C.c_void_p), ("dL_dconic", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dcolor", C.c_void_p),
        ("dL_dmean3D", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p), ("dL_dscale", C.c_void_p),
        ("dL_drot", C.c_void_p), ("blend_grad_scratch", C.c_void_p), ("debug", C.c_int), ("accumulate_param_grads", C.c_int),
        ("blend_only", C.c_int),
        ("dL_dpix2", C.c_void_p), ("background2", C.c_void_p), ("colors2", C.c_void_p), ("channels2", C.c_int), ("blend_grad_scratch2", C.c_void_p),
    ]


class ReblendArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("width", C.c_int), ("height", C.c_int),
        ("background", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("src_binning_buffer", C.c_void_p), ("src_image_buffer", C.c_void_p),
        ("out_color", C.c_void_p), ("debug", C.c_int), ("forward_only", C.c_int),
        ("src_viewmatrix", C.c_void_p), ("src_projmatrix", C.c_void_p), ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p),
    ]


class SugarArgs(C.Structure):
    """gstar_sugar_args (include/gstar_raster.h)."""
    _fields_ = [
        ("P", C.c_int), ("K", C.c_int),
        ("verts", C.c_void_p), ("faces32", C.c_void_p), ("faces64", C.c_void_p), ("bary", C.c_void_p),
        ("scales", C.c_void_p), ("cplx", C.c_void_p), ("dens", C.c_void_p),
        ("thickness", C.c_float), ("min_scale", C.c_float), ("max_scale", C.c_float), ("has_min", C.c_int), ("has_max", C.c_int),
        ("points", C.c_void_p), ("scaling", C.c_void_p), ("quats", C.c_void_p), ("opac", C.c_void_p),
        ("g_points", C.c_void_p), ("g_scaling", C.c_void_p), ("g_quats", C.c_void_p), ("g_opac", C.c_void_p),
        ("d_verts", C.c_void_p), ("d_scales", C.c_void_p), ("d_cplx", C.c_void_p), ("d_dens", C.c_void_p),
    ]


_lib = None


def lib():
    """Load libgstar_raster.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built; run `python gaustar_b200/build.py`")
        L = C.CDLL(LIB_PATH)
        L.gstar_last_error.restype = C.c_char_p
        L.gstar_stage_name.restype = C.c_char_p
        L.gstar_stage_name.argtypes = [C.c_int]
        for n in ("gstar_geom_bytes", "gstar_image_bytes", "gstar_binning_bytes"):
            getattr(L, n).restype = C.c_size_t
        L.gstar_geom_bytes.argtypes = [C.c_int]
        L.gstar_image_bytes.argtypes = [C.c_int, C.c_int]
        L.gstar_binning_bytes.argtypes = [C.c_size_t]
        L.gstar_raster_forward.argtypes = [C.POINTER(FwdArgs), ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, C.c_void_p]
        L.gstar_raster_backward.argtypes = [C.POINTER(BwdArgs), C.c_void_p]
        L.gstar_raster_reblend.argtypes = [C.POINTER(ReblendArgs), ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, C.c_void_p]
        L.gstar_mark_visible.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gstar_geom_unpack.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.gstar_image_views.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.gstar_binning_views.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.gstar_profile_stage.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.gstar_set_hit_log.argtypes = [C.c_int]
        L.gstar_set_deterministic.argtypes = [C.c_int]
        L.gstar_knn3_mean_dist2.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                            C.c_float, C.c_void_p, C.c_void_p]
        L.gstar_sugar_prologue_forward.argtypes = [C.POINTER(SugarArgs), C.c_void_p]
        L.gstar_sugar_prologue_backward.argtypes = [C.POINTER(SugarArgs), C.c_void_p]
        L.gstar_debug_header.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.gstar_hit_log_state.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        _lib = L
    return _lib


class GstarError(RuntimeError):
    pass


def _check(rc):
    if rc < 0:
        raise GstarError(f"gstar error {rc}: {lib().gstar_last_error().decode()}")
    return rc


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else C.c_void_p(t.data_ptr())


def _pad4(t):
    """[P,c] -> contiguous [P,4] (zeros in the unused channels): the second pass's colours are one float4 per Gaussian."""
    if t.shape[1] == 4:
        return t.contiguous()
    out = torch.zeros(t.shape[0], 4, dtype=t.dtype, device=t.device)
    out[:, :t.shape[1]] = t
    return out


def _f32(t, dev):
    if t is None:
        return None
    return t.to(device=dev, dtype=torch.float32).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _resizable(dev):
    """A torch uint8 tensor grown through the gstar_alloc_fn callback (cf. resizeFunctional, rasterize_points.cu:27-33).
    Returns (holder, callback); holder[0] is the tensor.  No reference cycle: the buffers are released as soon as the
    caller drops them (a cycle would park ~300 MB per call until Python's cyclic GC happens to run)."""
    holder = [torch.empty(0, dtype=torch.uint8, device=dev)]

    def alloc(_user, nbytes):
        holder[0].resize_(int(nbytes))
        return holder[0].data_ptr()

    return holder, ALLOC_FN(alloc)


def forward(means3D, opacities, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, W, H, shs=None, colors_precomp=None,
            scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=0, prefiltered=False, debug=False,
            forward_only=False, colors2=None, bg2=None):
    """gstar_raster_forward.  Returns dict(num_rendered, out_color, radii, geom, binning, image).
    forward_only: no backward will follow (the forward then skips the hit log).
    colors2 [P,c], bg2 [c], c = 1..4: a second feature pass blended in the same kernel (gstar_fwd_args::colors2); the dict then also
    holds out_color2 [c,H,W] (per channel what a call with that channel as colours and background renders, bit for bit)."""
    L = lib()
    dev = means3D.device
    assert dev.type == "cuda", "gaustar_b200 has no CPU path"
    keep = [_f32(x, dev) for x in (means3D, opacities, viewmatrix, projmatrix, campos, bg, shs, colors_precomp, scales, rotations, cov3D_precomp)]
    m3, op, vm, pm, cp, bgc, sh, col, sc, rot, cov = keep
    P = m3.shape[0]
    M = 0 if sh is None or sh.numel() == 0 else sh.shape[1]
    out_color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    (geom, geom_cb), (binning, binning_cb), (image, image_cb) = _resizable(dev), _resizable(dev), _resizable(dev)
    a = FwdArgs(P, sh_degree, M, _ptr(bgc), W, H, _ptr(m3), _ptr(sh), _ptr(col), _ptr(op), _ptr(sc), scale_modifier, _ptr(rot), _ptr(cov),
                _ptr(vm), _ptr(pm), _ptr(cp), tan_fovx, tan_fovy, int(prefiltered), _ptr(out_color), _ptr(radii), int(debug), int(forward_only))
    out_color2 = None
    if colors2 is not None:
        c2 = int(colors2.shape[1]) if P else max(int(bg2.numel()), 1)
        keep += [_pad4(_f32(colors2, dev)), _f32(bg2, dev)]
        out_color2 = torch.empty(c2, H, W, dtype=torch.float32, device=dev)
        a.colors2, a.background2, a.out_color2, a.channels2 = _ptr(keep[-2]), _ptr(keep[-1]), _ptr(out_color2), c2
    with torch.cuda.device(dev):
        R = _check(L.gstar_raster_forward(C.byref(a), geom_cb, None, binning_cb, None, image_cb, None, _stream(dev)))
    if P == 0:
        out_color.zero_()
        if out_color2 is not None:
            out_color2.zero_()
    out = dict(num_rendered=R, out_color=out_color, radii=radii, geom=geom[0], binning=binning[0], image=image[0], viewmatrix=vm, projmatrix=pm,
               _keep=keep)
    if out_color2 is not None:
        out["out_color2"] = out_color2
    return out


def reblend(src, colors_precomp, bg, W, H, debug=False, forward_only=False, viewmatrix=None, projmatrix=None):
    """gstar_raster_reblend: a second pass over the Gaussians/camera of forward (or re-blend) call `src` with other
    per-Gaussian colours (and background).  Returns a dict shaped like forward()'s: the geometry buffer and radii are the
    source call's (shared), binning and image are new -- pass it to backward() with colors_precomp=<these colours>.
    viewmatrix/projmatrix: this call's camera; if given it is compared with the source call's on the device and a
    mismatch yields a NaN image (header overflow word 2) instead of a picture of the wrong view."""
    L = lib()
    dev = src["geom"].device
    keep = [_f32(colors_precomp, dev), _f32(bg, dev), _f32(viewmatrix, dev), _f32(projmatrix, dev)]
    col, bgc, vm, pm = keep
    guard = vm is not None and pm is not None
    P = col.shape[0]
    out_color = torch.empty(3, H, W, dtype=torch.float32, device=dev)
    (binning, binning_cb), (image, image_cb) = _resizable(dev), _resizable(dev)
    a = ReblendArgs(P, W, H, _ptr(bgc), _ptr(col), _ptr(src["binning"]), _ptr(src["image"]), _ptr(out_color), int(debug), int(forward_only),
                    _ptr(src["viewmatrix"]) if guard else None, _ptr(src["projmatrix"]) if guard else None, _ptr(vm) if guard else None,
                    _ptr(pm) if guard else None)
    with torch.cuda.device(dev):
        R = _check(L.gstar_raster_reblend(C.byref(a), binning_cb, None, image_cb, None, _stream(dev)))
    return dict(num_rendered=R, out_color=out_color, radii=src["radii"], geom=src["geom"], binning=binning[0], image=image[0],
                viewmatrix=src["viewmatrix"], projmatrix=src["projmatrix"], _keep=keep + [src])


def backward(fwd, dL_dout_color, means3D, viewmatrix, projmatrix, campos, bg, tan_fovx, tan_fovy, shs=None, colors_precomp=None,
             scales=None, rotations=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=0, debug=False, accumulate_into=None, lean=False,
             scratch=None, dL_dout_color2=None, colors2=None, bg2=None, atomic_accumulate=False):
    """gstar_raster_backward.  Returns dict of the nine gradient tensors (reference layouts).
    scratch: optional [P,12] moment buffer pre-loaded by blend_moments() calls of other passes over the same geometry (the
    per-Gaussian stage then serves all of them at once); default: a fresh zero-filled one.
    lean: do not materialise the intermediates dL_dconic and -- for inputs that were not given -- dL_dcolors / dL_dcov3D
    (NULL in the C ABI; the dict then holds empty tensors for them).
    accumulate_into: optional dict with dL_dmeans3D/dL_dscales/dL_drotations/dL_dopacity/dL_dsh tensors; the
    parameter gradients are then ADDED to them inside the kernel (accumulate_param_grads=1; atomic_accumulate: =2, reductions at
    L2, so that calls running concurrently on different streams may add into the same tensors).
    dL_dout_color2, colors2, bg2: backward of a two-pass forward (forward(colors2=..., bg2=...)); the dict then also holds
    dL_dcolors2 [P,3] (floats 9..11 of the moment scratch)."""
    L = lib()
    dev = means3D.device
    keep = [_f32(x, dev) for x in (means3D, viewmatrix, projmatrix, campos, bg, shs, colors_precomp, scales, rotations, cov3D_precomp, dL_dout_color)]
    m3, vm, pm, cp, bgc, sh, col, sc, rot, cov, dpix = keep
    P = m3.shape[0]
    M = 0 if sh is None or sh.numel() == 0 else sh.shape[1]
    H, W = dpix.shape[1], dpix.shape[2]
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    g = dict(dL_dmeans2D=e(P, 3), dL_dconic=e(P, 4), dL_dopacity=e(P, 1), dL_dcolors=e(P, 3), dL_dmeans3D=e(P, 3), dL_dcov3D=e(P, 6),
             dL_dsh=e(P, M, 3), dL_dscales=e(P, 3), dL_drotations=e(P, 4))
    if lean:
        g["dL_dconic"] = e(0, 4)
        if col is None:
            g["dL_dcolors"] = e(0, 3)
        if cov is None:
            g["dL_dcov3D"] = e(0, 6)
    if accumulate_into is not None:
        for k in ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh"):
            t = accumulate_into[k]
            assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == g[k].numel(), k
            g[k] = t
    if scratch is None:
        scratch = torch.zeros(P, 12, dtype=torch.float32, device=dev)
    assert scratch.shape == (P, 12) and scratch.dtype == torch.float32 and scratch.is_contiguous()
    a = BwdArgs(P, sh_degree, M, int(fwd["num_rendered"]), _ptr(bgc), W, H, _ptr(m3), _ptr(sh), _ptr(col), _ptr(sc), scale_modifier, _ptr(rot),
                _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cp), tan_fovx, tan_fovy, _ptr(fwd["radii"]), _ptr(fwd["geom"]), _ptr(fwd["binning"]),
                _ptr(fwd["image"]), _ptr(dpix), _ptr(g["dL_dmeans2D"]), _ptr(g["dL_dconic"]), _ptr(g["dL_dopacity"]), _ptr(g["dL_dcolors"]),
                _ptr(g["dL_dmeans3D"]), _ptr(g["dL_dcov3D"]), _ptr(g["dL_dsh"]), _ptr(g["dL_dscales"]), _ptr(g["dL_drotations"]),
                _ptr(scratch), int(debug), (2 if atomic_accumulate else 1) if accumulate_into is not None else 0)
    scratch2 = None
    if dL_dout_color2 is not None:
        c2 = int(dL_dout_color2.shape[0])
        keep += [_f32(dL_dout_color2, dev), _f32(bg2, dev), _pad4(_f32(colors2, dev))]
        a.dL_dpix2, a.background2, a.colors2, a.channels2 = _ptr(keep[-3]), _ptr(keep[-2]), _ptr(keep[-1]), c2
        if c2 == 4:
            scratch2 = torch.zeros(P, dtype=torch.float32, device=dev)
            a.blend_grad_scratch2 = _ptr(scratch2)
    with torch.cuda.device(dev):
        _check(L.gstar_raster_backward(C.byref(a), _stream(dev)))
    if dL_dout_color2 is not None:
        g["dL_dcolors2"] = scratch[:, 9:9 + min(c2, 3)].clone() if c2 < 4 else torch.cat([scratch[:, 9:12], scratch2[:, None]], 1)
    g["_keep"] = keep + [scratch]
    return g


def blend_moments(fwd, dL_dout_color, bg, scratch):
    """gstar_raster_backward with blend_only=1: ADD the raw blend-stage moments of pass `fwd` (a forward or re-blend result)
    into `scratch` [P,12]; columns 6..8 are then this pass's dL_dcolors, columns 0..5 the geometric moments that add over
    the passes of one view.  Nothing else is computed."""
    L = lib()
    dev = scratch.device
    keep = [_f32(dL_dout_color, dev), _f32(bg, dev)]
    dpix, bgc = keep
    P = scratch.shape[0]
    assert scratch.shape == (P, 12) and scratch.dtype == torch.float32 and scratch.is_contiguous()
    a = BwdArgs()
    a.P, a.R, a.width, a.height = P, int(fwd["num_rendered"]), dpix.shape[2], dpix.shape[1]
    a.background, a.dL_dpix, a.blend_grad_scratch = _ptr(bgc), _ptr(dpix), _ptr(scratch)
    a.radii, a.geom_buffer, a.binning_buffer, a.image_buffer = _ptr(fwd["radii"]), _ptr(fwd["geom"]), _ptr(fwd["binning"]), _ptr(fwd["image"])
    a.blend_only = 1
    with torch.cuda.device(dev):
        _check(L.gstar_raster_backward(C.byref(a), _stream(dev)))
    return keep


def mark_visible(means3D, viewmatrix, projmatrix):
    dev = means3D.device
    m3, vm, pm = _f32(means3D, dev), _f32(viewmatrix, dev), _f32(projmatrix, dev)
    present = torch.zeros(m3.shape[0], dtype=torch.bool, device=dev)
    with torch.cuda.device(dev):
        _check(lib().gstar_mark_visible(m3.shape[0], _ptr(m3), _ptr(vm), _ptr(pm), _ptr(present), _stream(dev)))
    return present


def unpack_geometry(fwd, P):
    """Per-Gaussian intermediates in the reference's GeometryState layouts (rasterizer_impl.h:33-47)."""
    dev = fwd["geom"].device
    out = dict(depths=torch.zeros(P, device=dev), means2D=torch.zeros(P, 2, device=dev), conic_opacity=torch.zeros(P, 4, device=dev),
               rgb=torch.zeros(P, 3, device=dev), tiles_touched=torch.zeros(P, dtype=torch.int32, device=dev),
               clamped=torch.zeros(P, 3, dtype=torch.uint8, device=dev))
    with torch.cuda.device(dev):
        _check(lib().gstar_geom_unpack(_ptr(fwd["geom"]), P, _ptr(out["depths"]), _ptr(out["means2D"]), _ptr(out["conic_opacity"]),
                                       _ptr(out["rgb"]), _ptr(out["tiles_touched"]), _ptr(out["clamped"]), _stream(dev)))
    return out


def _view(base: torch.Tensor, addr: int, nbytes: int, dtype):
    off = addr - base.data_ptr()
    return base[off:off + nbytes].view(dtype)


def image_state(fwd, W, H):
    """final_T[H*W], n_contrib[H*W] (int32 view of uint32), ranges[T,2] -- views into the image buffer."""
    ft, nc, rg = C.c_void_p(), C.c_void_p(), C.c_void_p()
    _check(lib().gstar_image_views(_ptr(fwd["image"]), W, H, C.byref(ft), C.byref(nc), C.byref(rg)))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    img = fwd["image"]
    return dict(final_T=_view(img, ft.value, W * H * 4, torch.float32), n_contrib=_view(img, nc.value, W * H * 4, torch.int32),
                ranges=_view(img, rg.value, T * 8, torch.int32).view(T, 2))


def point_list(fwd):
    """Sorted instance list (BinningState::point_list): int32 view of the first num_rendered entries."""
    R = int(fwd["num_rendered"])
    if R == 0:
        return torch.zeros(0, dtype=torch.int32, device=fwd["geom"].device)
    pl = C.c_void_p()
    cap = C.c_uint64()
    _check(lib().gstar_binning_views(_ptr(fwd["binning"]), _ptr(fwd["image"]), C.byref(pl), C.byref(cap)))
    return _view(fwd["binning"], pl.value, R * 4, torch.int32)


def profile_stage(stage: int, start: Optional[torch.cuda.Event] = None, stop: Optional[torch.cuda.Event] = None):
    """Record start/stop around kernel stage `stage` of every later call on this thread (stage < 0: off)."""
    s = C.c_void_p(start.cuda_event) if start is not None else None
    e = C.c_void_p(stop.cuda_event) if stop is not None else None
    _check(lib().gstar_profile_stage(stage, s, e))


def set_hit_log(mode: int) -> int:
    """gstar_set_hit_log: 0 = walk-back backward only, 1 = hit log when it fits (default).  Returns the previous mode."""
    return int(lib().gstar_set_hit_log(int(mode)))


def set_deterministic(on) -> bool:
    """gstar_set_deterministic: fixed-order gradient sums in the blend backward (bit-identical gradients run to run; test mode,
    needs the hit log).  Returns the previous mode; pass -1 to query."""
    return bool(lib().gstar_set_deterministic(int(on)))


def hit_log_state(fwd):
    """(slots_needed, slots_capacity, in_use) of one forward call -- synchronous, for tests."""
    need, cap, used = C.c_uint64(), C.c_uint64(), C.c_int()
    _check(lib().gstar_hit_log_state(_ptr(fwd["image"]), C.byref(need), C.byref(cap), C.byref(used)))
    return int(need.value), int(cap.value), bool(used.value)


def debug_header(fwd):
    """The forward call's device header as a dict (private layout; diagnostics only)."""
    w = (C.c_uint32 * 24)()
    _check(lib().gstar_debug_header(_ptr(fwd["image"]), w))
    w = list(w)
    return dict(num_rendered=w[0], capacity=w[1], overflow=w[2], max_tile=w[3], log_overflow=w[4], sort_fallback_tiles=w[5], sort_max_fine=w[6],
                log_cursor=w[8] | (w[9] << 32), log_capacity=w[10] | (w[11] << 32), cls_end=w[16:20], cls_cursor=w[20:24])
