"""the caller-side per-Gaussian prologue of a mesh-bound SuGaR model as one fused CUDA op.

Every render call of `gaustar_scene/sugar_model.py` rebuilds the rasterizer's per-Gaussian inputs from the mesh with a few dozen
torch kernels -- `points` (:417-435), `scaling` (:457-476), `quaternions` (:478-508, incl. pytorch3d's `matrix_to_quaternion`) and
`strengths` (:443-447) -- and autograd differentiates each of them again; GauSTAR's training step does this twice per iteration
(RGB and depth renders).  With the rasterizer itself at ~0.7 ms per view that torch graph is the larger half of an iteration.

`fused_gaussian_params` computes the four tensors with ONE kernel (`gstar_sugar_prologue_forward`, csrc/sugar_prologue.cu) and
their backward with one more (vertex gradients by atomics); `patch_sugar(model)` installs it behind the four properties of a SuGaR
*instance* -- nothing in GauSTAR is edited, and an un-patched model keeps working -- with a cache keyed on the parameters'
versions and dropped when its graph is back-propagated, so the RGB and depth renders of one iteration share one evaluation (and one
backward).  Not supported (the patch
refuses): models not bound to a surface mesh, `editable` rescaling, loose binding (`_delta_t` / `_delta_r` active).

For the colour half of the prologue (`get_points_rgb`, :674-718) no new op is needed: call the renderer with
`compute_color_in_rasterizer=True` and the SH evaluation (and its backward) happens inside the rasterizer's preprocess kernels.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import capi


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _FusedGaussianParams(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, scales, cplx, dens, faces, bary, K, thickness, min_scale, max_scale, on_backward=None):
        ctx.on_backward = on_backward
        dev = verts.device
        if dev.type != "cuda":
            raise RuntimeError("gaustar_b200.sugar has no CPU path")
        v = verts.detach().to(torch.float32).contiguous()
        sc = scales.detach().to(torch.float32).contiguous()
        cx = cplx.detach().to(torch.float32).contiguous()
        de = dens.detach().to(torch.float32).contiguous().view(-1)
        fa = faces.contiguous()
        if fa.dtype not in (torch.int32, torch.int64):
            fa = fa.to(torch.int64)
        ba = bary.detach().to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
        P = sc.shape[0]
        if fa.shape[0] * K != P or ba.shape[0] != K or cx.shape[0] != P or de.shape[0] != P:
            raise ValueError("fused_gaussian_params: expected P = n_faces * n_gaussians_per_face rows in scales / complex numbers / densities")
        points = torch.empty(P, 3, device=dev)
        scaling = torch.empty(P, 3, device=dev)
        quats = torch.empty(P, 4, device=dev)
        opac = torch.empty(P, 1, device=dev)
        a = capi.SugarArgs()
        a.P, a.K = P, K
        a.verts, a.bary, a.scales, a.cplx, a.dens = v.data_ptr(), ba.data_ptr(), sc.data_ptr(), cx.data_ptr(), de.data_ptr()
        a.faces32 = fa.data_ptr() if fa.dtype == torch.int32 else None
        a.faces64 = fa.data_ptr() if fa.dtype == torch.int64 else None
        a.thickness = float(thickness)
        a.has_min, a.min_scale = (1, float(min_scale)) if min_scale is not None else (0, 0.0)
        a.has_max, a.max_scale = (1, float(max_scale)) if max_scale is not None else (0, 0.0)
        a.points, a.scaling, a.quats, a.opac = points.data_ptr(), scaling.data_ptr(), quats.data_ptr(), opac.data_ptr()
        with torch.cuda.device(dev):
            capi._check(capi.lib().gstar_sugar_prologue_forward(C.byref(a), _stream(dev)))
        ctx.save_for_backward(v, sc, cx, de, fa, ba)
        ctx.meta = (K, a.thickness, min_scale, max_scale, verts.shape, dens.shape)
        return points, scaling, quats, opac

    @staticmethod
    def backward(ctx, g_points, g_scaling, g_quats, g_opac):
        if ctx.on_backward is not None:
            ctx.on_backward()  # (a cached result whose graph has been consumed must not be handed out again)
        v, sc, cx, de, fa, ba = ctx.saved_tensors
        K, thickness, min_scale, max_scale, vshape, dshape = ctx.meta
        dev = v.device
        P = sc.shape[0]
        cont = lambda t: None if t is None else t.to(torch.float32).contiguous()
        g_points, g_scaling, g_quats, g_opac = cont(g_points), cont(g_scaling), cont(g_quats), cont(g_opac)
        need_v, need_s, need_c, need_d = ctx.needs_input_grad[:4]
        d_verts = torch.zeros(v.shape, device=dev) if need_v else None
        d_scales = torch.empty(P, 2, device=dev) if need_s else None
        d_cplx = torch.empty(P, 2, device=dev) if need_c else None
        d_dens = torch.empty(P, device=dev) if need_d else None
        a = capi.SugarArgs()
        a.P, a.K = P, K
        a.verts, a.bary, a.scales, a.cplx, a.dens = v.data_ptr(), ba.data_ptr(), sc.data_ptr(), cx.data_ptr(), de.data_ptr()
        a.faces32 = fa.data_ptr() if fa.dtype == torch.int32 else None
        a.faces64 = fa.data_ptr() if fa.dtype == torch.int64 else None
        a.thickness = thickness
        a.has_min, a.min_scale = (1, float(min_scale)) if min_scale is not None else (0, 0.0)
        a.has_max, a.max_scale = (1, float(max_scale)) if max_scale is not None else (0, 0.0)
        p = lambda t: None if t is None else t.data_ptr()
        a.g_points, a.g_scaling, a.g_quats, a.g_opac = p(g_points), p(g_scaling), p(g_quats), p(g_opac)
        a.d_verts, a.d_scales, a.d_cplx, a.d_dens = p(d_verts), p(d_scales), p(d_cplx), p(d_dens)
        with torch.cuda.device(dev):
            capi._check(capi.lib().gstar_sugar_prologue_backward(C.byref(a), _stream(dev)))
        return (d_verts.view(vshape) if need_v else None, d_scales, d_cplx, d_dens.view(dshape) if need_d else None, None, None, None, None, None, None, None)


def fused_gaussian_params(verts, faces, scales, complex_numbers, densities, bary_coords, thickness, min_scale=None, max_scale=None, _on_backward=None):
    """(points [P,3], scaling [P,3], quaternions [P,4], strengths [P,1]) of P = F*K Gaussians bound K per face to the mesh
    (verts [Nv,3], faces [F,3] int32/int64), exactly as SuGaR's properties compute them (sugar_model.py:417-508); differentiable in
    verts, scales [P,2], complex_numbers [P,2] and densities [P,1]."""
    K = int(bary_coords.reshape(-1, 3).shape[0])
    return _FusedGaussianParams.apply(verts, scales, complex_numbers, densities, faces, bary_coords, K, thickness, min_scale, max_scale, _on_backward)


def patch_sugar(model):
    """Route `model.points / .scaling / .quaternions / .strengths` (a mesh-bound SuGaR instance) through the fused op.  The class and
    every other instance stay untouched; `unpatch_sugar(model)` restores the instance.  One evaluation is shared by all reads until a
    parameter changes (tensor versions), e.g. by the RGB and depth renders of one training iteration."""
    if not getattr(model, "binded_to_surface_mesh", False):
        raise ValueError("patch_sugar: the model is not bound to a surface mesh (sugar_model.py:166)")
    if getattr(model, "editable", False):
        raise ValueError("patch_sugar: `editable` rescaling (sugar_model.py:467-471) is not part of the fused op")
    base = type(model)
    state = {"key": None, "out": None}

    def params(self):
        if getattr(self, "_loose_bind", False) or getattr(self, "return_one_densities", False):
            return None  # fall back to the reference's own code paths
        ps = (self._points, self._scales, self._quaternions, self.all_densities)
        key = tuple((t.data_ptr(), t._version, t.requires_grad) for t in ps) + (torch.is_grad_enabled(),)
        if state["key"] != key:
            th = self.surface_mesh_thickness
            state["out"] = fused_gaussian_params(self._points, self._surface_mesh_faces, self._scales, self._quaternions, self.all_densities,
                                                 self.surface_triangle_bary_coords, float(th.item() if torch.is_tensor(th) else th),
                                                 self.min_gaussian_scale, self.max_gaussian_scale, _on_backward=lambda: state.update(key=None))
            state["key"] = key
        return state["out"]

    def prop(i, name):
        ref = getattr(base, name)

        def get(self):
            out = params(self)
            return ref.fget(self) if out is None else out[i]
        return property(get)

    patched = type(base.__name__, (base,), {"points": prop(0, "points"), "scaling": prop(1, "scaling"), "quaternions": prop(2, "quaternions"),
                                            "strengths": prop(3, "strengths"), "_gaustar_b200_unpatched": base})
    model.__class__ = patched
    return model


def unpatch_sugar(model):
    base = getattr(type(model), "_gaustar_b200_unpatched", None)
    if base is not None:
        model.__class__ = base
    return model
