"""Operator boundary: the Python surface of ``diff_gaussian_rasterization``.

Same names, argument order, return values and error behaviour as the reference
wrapper (DGR/diff_gaussian_rasterization/__init__.py:21-220, DGR =
gaussian_splatting/submodules/diff-gaussian-rasterization), so that
``gaustar_scene/sugar_model.py:1173-1293`` and
``gaussian_splatting/gaussian_renderer/__init__.py:36-93`` run unchanged.  All
compute happens in ``_C`` (gaustar_b200/csrc/torch_ext.cpp -> C ABI -> CUDA).
"""
import contextlib
import os
import threading
from typing import NamedTuple

import torch
import torch.nn as nn

try:
    from . import _C  # built in-tree by gaustar_b200/build.py
except ImportError as e:  # fail loudly: there is no fallback implementation
    raise ImportError(
        "gaustar_b200._C (the sm_100a CUDA extension) is not built or cannot be loaded: "
        f"{e}.  Run `python gaustar_b200/build.py`."
    ) from e


_GRAD_FUSION = False
_GRAD_FUSION_ATOMIC = False


def set_grad_accumulation_fusion(enabled: bool, atomic: bool = False) -> bool:
    """Gradient-accumulation fusion for multi-view steps (off by default; returns the previous setting).

    When on, and every differentiable parameter input of a call (means3D, opacities, scales, rotations, shs) is a LEAF
    whose ``.grad`` is already allocated (contiguous float32, the parameter's shape), the backward kernel adds the
    view's gradients straight into those ``.grad`` tensors and the autograd function returns ``None`` for them --
    instead of writing five fresh tensors that AccumulateGrad then re-reads and sums (236 MB per view at 1 M Gaussians /
    SH degree 3).  Point the ``.grad`` tensors at views of one flat buffer (``gaustar_b200.dist.FlatGrads``) and a
    multi-view step needs a single all-reduce.  Any call that does not qualify takes the ordinary path, so results are
    the same either way; tensor hooks on those leaves are not run for fused calls.

    atomic=True: the kernel adds with reductions at L2 instead of a plain read-modify-write, so backward passes running at the
    same time on different CUDA streams may share ONE set of ``.grad`` tensors (several sets of leaves over the same storage whose
    ``.grad`` all point at the same buffer) -- no per-stream buffers, no merge before the all-reduce."""
    global _GRAD_FUSION, _GRAD_FUSION_ATOMIC
    old, _GRAD_FUSION, _GRAD_FUSION_ATOMIC = _GRAD_FUSION, bool(enabled), bool(atomic)
    return old


_DET = {"explicit": False, "current": False}


def set_deterministic_backward(enabled: bool) -> bool:
    """Fixed-order gradient sums in the blend backward: gradients BIT-IDENTICAL from run to run (include/gstar_raster.h,
    gstar_set_deterministic; the reference sums with fp32 atomics in scheduling order, backward.cu:523-554, and has no such mode).
    Also taken whenever ``torch.use_deterministic_algorithms(True)`` is in force.  A test mode: roughly twice the blend-backward
    time, R x 48 bytes of scratch per backward, needs the hit log (default).  Returns the previous explicit setting."""
    old, _DET["explicit"] = _DET["explicit"], bool(enabled)
    return old


def _sync_deterministic():
    want = _DET["explicit"] or torch.are_deterministic_algorithms_enabled()
    if want != _DET["current"]:
        from . import capi  # same libgstar_raster.so as _C links: one process-wide switch
        capi.set_deterministic(want)
        _DET["current"] = want


def _fusable(t, needs):
    """t: an input of the autograd function.  Empty (absent) inputs are trivially fine."""
    if t.numel() == 0:
        return True
    if not (needs and t.is_leaf):
        return False
    g = t.grad
    return (g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()
            and g.shape == t.shape and g.device == t.device)


# ---- shared-geometry re-blend ----
# GauSTAR rasterizes the same Gaussians from the same camera twice per training step -- RGB, then depth as three equal
# colour channels (gaustar_trainers/refine.py:552-564 and :607-616) -- and three times in refined_mesh.py:733-774.  The
# reference repeats preprocess / duplicateWithKeys / radix sort every time.  Here a call that passes `colors_precomp`
# and whose geometry inputs and camera are those of the previous full forward starts from that call's sorted record
# stream (gstar_raster_reblend): only the blend runs again.  Two ways to say "same geometry":
#   set_geometry_cache(True)   automatic and strict: the geometry tensors and camera matrices must be the very same
#                              tensor objects' memory, unmodified (data_ptr + _version), same stream, same thread.
#                              CAVEAT: writes through `param.data` (some optimizers / densifiers) do not bump `_version`;
#                              code that edits parameters that way must use shared_geometry() blocks instead (or call
#                              release_shared_geometry() after the edit);
#   with shared_geometry():    asserted by the caller for the calls inside the block (GauSTAR's properties rebuild
#                              `points`/`scaling`/`quaternions` and the camera matrices on every call, so the tensors are
#                              equal in value but not in identity); shared_geometry(check=True) verifies the values
#                              (device->host syncs; for debugging).
# Off by default: the source call's buffers (geometry, binning incl. hit log, image) stay alive until the next full
# forward replaces them.
_GEOM_CACHE = os.environ.get("GSTAR_GEOM_CACHE", "0") not in ("", "0")
_tls = threading.local()


class _GeomSource:
    """What a later re-blend needs of a full forward call."""
    __slots__ = ("key", "tensors", "geom", "binning", "image", "radii", "num_rendered", "cam")


def set_geometry_cache(enabled: bool) -> bool:
    """Automatic (identity-keyed) shared-geometry re-blend; returns the previous setting."""
    global _GEOM_CACHE
    old, _GEOM_CACHE = _GEOM_CACHE, bool(enabled)
    if not enabled:
        _tls.src = None
    return old


def release_shared_geometry() -> None:
    """Forget the remembered source call of this thread (its geometry / binning incl. hit log / image buffers -- up to a few GB --
    are otherwise kept alive until the next full forward replaces them)."""
    _tls.src = None


@contextlib.contextmanager
def shared_geometry(check: bool = False):
    """Every rasterizer call inside the block sees the same Gaussians (positions, opacities, scales/rotations or
    covariances) through the same camera and image size; calls after the first that pass ``colors_precomp`` only
    re-blend.  Put it around the renders of ONE camera, not around a loop over cameras.  The camera matrices of a
    re-blend are verified on the device (a mismatch yields a NaN image); the per-Gaussian tensors are not --
    ``check=True`` compares all values with the first call's (slow, synchronizes; raises on mismatch)."""
    old = getattr(_tls, "scope", None)
    _tls.scope = {"check": bool(check)}
    _tls.src = None
    try:
        yield
    finally:
        _tls.scope = old
        _tls.src = None


def _tkey(t):
    if not isinstance(t, torch.Tensor) or t.numel() == 0:
        return None
    try:
        version = t._version
    except RuntimeError:  # inference tensors do not track versions: never equal to a remembered key -> a full forward
        version = object()
    return (t.data_ptr(), version, tuple(t.shape), str(t.device), t.dtype)


def _geom_inputs(means3D, opacities, scales, rotations, cov3Ds_precomp, rs):
    return (means3D, opacities, scales, rotations, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix)


def _geom_key(tensors, rs, dev):
    stream = torch.cuda.current_stream(dev).cuda_stream if torch.device(dev).type == "cuda" else 0
    return (tuple(_tkey(t) for t in tensors), int(rs.image_height), int(rs.image_width), float(rs.tanfovx), float(rs.tanfovy),
            float(rs.scale_modifier), bool(rs.prefiltered), stream)


def _reblend_allowed(key, src_key, scope_active: bool, cache_on: bool) -> bool:
    """May a call with geometry key `key` re-blend the remembered call `src_key`?  (pure host logic)
    Outside a shared_geometry() block: only with the cache on and IDENTICAL keys (same memory, same versions, same sizes,
    scalars and stream).  Inside a block the caller vouches for the values: sizes, scalars, stream and every tensor's
    presence / shape / device / dtype must still agree -- anything else is a different configuration, i.e. a full forward."""
    if not scope_active:
        return cache_on and key == src_key
    if key[1:] != src_key[1:]:
        return False
    return all((a is None) == (b is None) and (a is None or a[2:] == b[2:]) for a, b in zip(key[0], src_key[0]))


def _remember_source(means3D, opacities, scales, rotations, cov3Ds_precomp, rs, geom, binning, image, radii, num_rendered):
    if not (_GEOM_CACHE or getattr(_tls, "scope", None) is not None) or means3D.shape[0] == 0:
        return
    src = _GeomSource()
    src.tensors = _geom_inputs(means3D, opacities, scales, rotations, cov3Ds_precomp, rs)  # kept alive: their addresses cannot be recycled
    src.key = _geom_key(src.tensors, rs, means3D.device)
    src.geom, src.binning, src.image, src.radii, src.num_rendered = geom, binning, image, radii, num_rendered
    # the camera AS IT WAS when this call ran (2 x 16 floats, copied on the call's stream): the device-side guard of a re-blend
    # compares against this snapshot -- a matrix overwritten in place between the two calls compares equal to itself
    src.cam = (rs.viewmatrix.detach().to(device=means3D.device, dtype=torch.float32).reshape(-1).clone(),
               rs.projmatrix.detach().to(device=means3D.device, dtype=torch.float32).reshape(-1).clone())
    _tls.src = src


def _matching_source(means3D, opacities, scales, rotations, cov3Ds_precomp, colors_precomp, sh, rs):
    """The remembered full forward this call may re-blend, or None."""
    src = getattr(_tls, "src", None)
    if src is None or rs.debug or sh.numel() != 0 or colors_precomp.numel() == 0 or not means3D.is_cuda:
        return None
    if not (colors_precomp.is_cuda and colors_precomp.dim() == 2 and colors_precomp.shape == (means3D.shape[0], 3)
            and colors_precomp.dtype == torch.float32):
        return None
    tensors = _geom_inputs(means3D, opacities, scales, rotations, cov3Ds_precomp, rs)
    key = _geom_key(tensors, rs, means3D.device)
    scope = getattr(_tls, "scope", None)
    if not _reblend_allowed(key, src.key, scope is not None, _GEOM_CACHE):
        return None
    if scope is not None and scope["check"]:
        for name, a, b in zip(("means3D", "opacities", "scales", "rotations", "cov3D_precomp", "viewmatrix", "projmatrix"), tensors, src.tensors):
            if a.numel() and not torch.equal(a.to(b.device), b):
                raise RuntimeError(f"shared_geometry(check=True): `{name}` differs from the first call of the block")
    return src


def _cpu_snapshot(args):
    # DGR/__init__.py:17-19: inputs are copied before the call so a crash can be replayed
    return tuple(a.cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    """DGR/__init__.py:21-42."""
    src = _matching_source(means3D, opacities, scales, rotations, cov3Ds_precomp, colors_precomp, sh, raster_settings)
    if src is not None:
        return _ReblendGaussians.apply(means3D, means2D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, src)
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    """DGR/__init__.py:44-155.  Gradients come back in input order:
    (means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, None)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                rs.debug)
        if rs.debug:
            snapshot = _cpu_snapshot(args)
            try:
                num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            except Exception:
                torch.save(snapshot, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            # inference (no input requires grad, e.g. the renders of refined_mesh.py): tell the forward that no backward follows
            fwd_only = not any(ctx.needs_input_grad)
            if fwd_only:
                _C.set_forward_only(True)
            try:
                num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            finally:
                if fwd_only:
                    _C.set_forward_only(False)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer)
        # the caller's own tensor objects (their .grad is the accumulation target of a fused backward)
        ctx.param_inputs = (means3D, sh, opacities, scales, rotations) if _GRAD_FUSION else None
        ctx.mark_non_differentiable(radii)
        _remember_source(means3D, opacities, scales, rotations, cov3Ds_precomp, rs, geomBuffer, binningBuffer, imgBuffer, radii, num_rendered)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        _sync_deterministic()
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, grad_out_color, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer,
                imgBuffer, rs.debug)
        pin = ctx.param_inputs
        if (_GRAD_FUSION and pin is not None and not rs.debug and cov3Ds_precomp.numel() == 0
                and all(_fusable(t, ctx.needs_input_grad[i]) for t, i in zip(pin, (0, 2, 4, 5, 6)))):
            m3, shp, op, sc, rot = pin
            e = torch.Tensor([])
            out = _C.rasterize_gaussians_backward_fused(*args, m3.grad, shp.grad if shp.numel() else e, op.grad, sc.grad, rot.grad, _GRAD_FUSION_ATOMIC)
            grad_means2D, grad_colors_precomp, _go, _gm, grad_cov3Ds_precomp, _gs, _gsc, _gr = out
            return (None, grad_means2D, None, grad_colors_precomp, None, None, None, None, None)
        if rs.debug:
            snapshot = _cpu_snapshot(args)
            try:
                out = _C.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(snapshot, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            out = _C.rasterize_gaussians_backward(*args)
        grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations = out
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, None)


class _ReblendGaussians(torch.autograd.Function):
    """Second pass over the Gaussians and camera of full forward `src` with other per-Gaussian colours
    (gstar_raster_reblend).  Same outputs and gradients as _RasterizeGaussians on the same inputs; the geometry buffer
    and radii are the source call's.  Gradients in input order:
    (means3D, means2D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, None, None)."""

    @staticmethod
    def forward(ctx, means3D, means2D, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, src):
        rs = raster_settings
        fwd_only = not any(ctx.needs_input_grad)
        if fwd_only:
            _C.set_forward_only(True)
        try:
            # the camera of this call is compared with the source call's on the device: a mismatch (a shared_geometry() block
            # put around a loop over cameras) gives a NaN image, not a plausible picture of the wrong view
            num_rendered, color, binningBuffer, imgBuffer = _C.rasterize_gaussians_reblend(
                rs.bg, colors_precomp, rs.image_height, rs.image_width, src.binning, src.image, rs.debug, src.cam[0], src.cam[1],
                rs.viewmatrix, rs.projmatrix)
        finally:
            if fwd_only:
                _C.set_forward_only(False)
        radii = src.radii.detach()  # a tensor object of this call's own (same memory)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, src.geom, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        _sync_deterministic()
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        out = _C.rasterize_gaussians_backward(rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                                              rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, torch.Tensor([]),
                                              rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer, rs.debug)
        grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, _grad_sh, grad_scales, grad_rotations = out
        return (grad_means3D, grad_means2D, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations, grad_cov3Ds_precomp,
                None, None)


_PASS_FUSION = True
_ERR_NOLOG = -5  # GSTAR_ERR_NOLOG


def set_pass_fusion(enabled: bool) -> bool:
    """forward_passes(): blend the first extra pass in the SAME kernel as the main pass (on by default; returns the previous
    setting).  Off: every extra pass is a re-blend of its own, as in round 1.  Results are the same either way -- images bit for
    bit, gradients up to the order of fp32 sums."""
    global _PASS_FUSION
    old, _PASS_FUSION = _PASS_FUSION, bool(enabled)
    return old


class _RasterizePasses(torch.autograd.Function):
    """ONE autograd node for several feature passes over the same Gaussians and camera (config #5: RGB through SH, depth,
    normals): a forward that blends the main pass and the leading extra passes -- as many as fit FOUR channels together, e.g. depth
    as one channel and a normal as three -- in one kernel (the passes share every pair's alpha and transmittance:
    gstar_fwd_args::colors2), a re-blend per further (colours [P,3], background) pair, and a backward whose blend
    stage serves the two fused passes at once and whose per-Gaussian stage (cov2D / projection / SH / cov3D,
    backward.cu:144-396) runs ONCE for all passes -- the six geometric blend moments of the passes add, the three colour
    moments of an extra pass are its dL_dcolors (gstar_bwd_args.blend_only).  A view whose hit log cannot be had (switched off,
    or larger than GSTAR_HIT_LOG_MAX_MB) re-blends the first extra pass as well.  Inputs: the nine of _RasterizeGaussians, a
    tuple of backgrounds, then the extra colour tensors.  Outputs: (color, radii, *extra_images).  dL_dmeans2D is the sum over
    the passes."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings, extra_bgs,
                *extra_colors):
        rs = raster_settings
        args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
                rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                rs.debug)
        fwd_only = not any(ctx.needs_input_grad) and not rs.debug
        if fwd_only:
            _C.set_forward_only(True)
        dual = 0  # extra passes blended together with the main pass
        col2 = bg2 = None
        try:
            extras, bufs = [], []
            if _PASS_FUSION and extra_colors and means3D.shape[0] > 0:
                nch, chans = 0, []
                for c_k in extra_colors:  # the leading passes that fit four channels together
                    if nch + c_k.shape[1] > 4:
                        break
                    nch += c_k.shape[1]
                    chans.append(c_k.shape[1])
                if chans:
                    col2 = extra_colors[0] if len(chans) == 1 else torch.cat(extra_colors[:len(chans)], 1)
                    bg2 = torch.cat([b.reshape(-1).to(device=means3D.device, dtype=torch.float32) for b in extra_bgs[:len(chans)]])
                    r = _C.rasterize_gaussians_dual(*args, col2, bg2)
                    if r[0] != _ERR_NOLOG:
                        dual = len(chans)
                        num_rendered, color, img2, radii, geomBuffer, binningBuffer, imgBuffer = r
                        extras += list(torch.split(img2, chans, 0)) if dual > 1 else [img2]
            if not dual:
                num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = _C.rasterize_gaussians(*args)
            for bg_k, col_k in list(zip(extra_bgs, extra_colors))[dual:]:
                if means3D.shape[0] == 0:
                    extras.append(torch.zeros(col_k.shape[1], rs.image_height, rs.image_width, dtype=torch.float32, device=means3D.device))
                    continue
                if col_k.shape[1] != 3:
                    raise Exception('extra_passes: a pass that is not blended together with the first one must have 3 channels')
                e = torch.Tensor([])  # same camera by construction: no guard
                _, img_k, bin_k, ib_k = _C.rasterize_gaussians_reblend(bg_k, col_k, rs.image_height, rs.image_width, binningBuffer, imgBuffer,
                                                                       rs.debug, e, e, e, e)
                extras.append(img_k)
                bufs += [bin_k, ib_k]
        finally:
            if fwd_only:
                _C.set_forward_only(False)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.extra_bgs = tuple(extra_bgs)
        ctx.dual = dual
        ctx.dual_chans = [c_k.shape[1] for c_k in extra_colors[:dual]]
        ctx.dual_bg = bg2 if dual else None
        ctx.set_materialize_grads(False)  # an extra image nobody differentiated costs no backward pass
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
                              col2 if dual else torch.empty(0, device=means3D.device), *bufs)
        ctx.mark_non_differentiable(radii)
        return (color, radii, *extras)

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, *grad_extras):
        _sync_deterministic()
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer, colors2, *bufs = ctx.saved_tensors
        P = means3D.shape[0]
        n_extra = len(ctx.extra_bgs)
        first = ctx.dual  # extra passes from here on have binning / image buffers of their own
        extra_grads = [None] * n_extra
        if grad_out_color is None:
            grad_out_color = torch.zeros(3, rs.image_height, rs.image_width, dtype=torch.float32, device=means3D.device)
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, grad_out_color, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered, binningBuffer, imgBuffer,
                rs.debug)
        if P == 0:
            out = _C.rasterize_gaussians_backward(*args)
            extra_grads = [torch.zeros(0, 3, device=means3D.device) if ctx.needs_input_grad[10 + k] else None for k in range(n_extra)]
        else:
            scratch = torch.zeros(P, 12, dtype=torch.float32, device=means3D.device)
            for k, g_k in enumerate(grad_extras):
                if g_k is None or k < first:
                    continue
                _C.rasterize_gaussians_blend_backward(ctx.extra_bgs[k], g_k, radii, geomBuffer, ctx.num_rendered, bufs[2 * (k - first)],
                                                      bufs[2 * (k - first) + 1], scratch, rs.debug)
                if ctx.needs_input_grad[10 + k]:
                    extra_grads[k] = scratch[:, 6:9].clone()
                scratch[:, 6:9].zero_()  # the colour moments are per pass; the geometric ones (columns 0..5) add
            if ctx.dual:
                H, W = rs.image_height, rs.image_width
                parts = [grad_extras[k] if grad_extras[k] is not None else torch.zeros(c, H, W, dtype=torch.float32, device=means3D.device)
                         for k, c in enumerate(ctx.dual_chans)]
                g2 = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
                nch = g2.shape[0]
                scratch2 = torch.zeros(P if nch == 4 else 0, dtype=torch.float32, device=means3D.device)
                out = _C.rasterize_gaussians_backward_dual(*args, scratch, g2, ctx.dual_bg, colors2, scratch2)
                moments = scratch[:, 9:9 + min(nch, 3)] if nch < 4 else torch.cat([scratch[:, 9:12], scratch2[:, None]], 1)  # the fused passes' colour moments
                off = 0
                for k, c in enumerate(ctx.dual_chans):
                    if ctx.needs_input_grad[10 + k]:
                        extra_grads[k] = moments[:, off:off + c].clone()
                    off += c
            else:
                out = _C.rasterize_gaussians_backward_preloaded(*args, scratch)
        grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales, grad_rotations = out
        return (grad_means3D, grad_means2D, grad_sh, grad_colors_precomp, grad_opacities, grad_scales, grad_rotations,
                grad_cov3Ds_precomp, None, None, *extra_grads)


class GaussianRasterizationSettings(NamedTuple):
    """DGR/__init__.py:157-169 -- field order is part of the contract (callers build it by keyword)."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    """DGR/__init__.py:171-220."""

    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def _checked_inputs(self, shs, colors_precomp, scales, rotations, cov3D_precomp):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])  # absent inputs travel as empty CPU tensors -> NULL in the C ABI
        return tuple(empty if t is None else t for t in (shs, colors_precomp, scales, rotations, cov3D_precomp))

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        shs, colors_precomp, scales, rotations, cov3D_precomp = self._checked_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, rs)

    def forward_passes(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                       extra_passes=()):
        """Several feature passes over one geometry in ONE call (no counterpart in the reference).
        The arguments of forward() describe the first pass; ``extra_passes`` is a sequence of ``(colors [P,c], bg [c])`` pairs,
        c = 1..4, each rendered like a forward() call with ``colors_precomp=colors`` and that background.  The leading passes that
        fit four channels together -- depth as ONE channel plus a normal, say -- are blended in the same kernel as the first pass
        (a seven-channel blend); a pass beyond that is a re-blend and must have 3 channels, and so must every extra pass when the
        hit log is switched off.  Returns
        ``(color, radii, [extra images])``.  Preprocess, binning and sort run once, and so does the per-Gaussian stage of
        the backward; ``means2D.grad`` receives the sum over the passes."""
        rs = self.raster_settings
        shs, colors_precomp, scales, rotations, cov3D_precomp = self._checked_inputs(shs, colors_precomp, scales, rotations, cov3D_precomp)
        cols, bgs = [], []
        for col, bg in extra_passes:
            if not (isinstance(col, torch.Tensor) and col.is_cuda and col.dtype == torch.float32 and col.dim() == 2
                    and col.shape[0] == means3D.shape[0] and 1 <= col.shape[1] <= 4 and bg.numel() == col.shape[1]):
                raise Exception('extra_passes: colors must be float32 CUDA tensors of shape (P, c), c = 1..4, with a background of c values')
            cols.append(col)
            bgs.append(bg)
        out = _RasterizePasses.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, rs, tuple(bgs), *cols)
        return out[0], out[1], list(out[2:])
