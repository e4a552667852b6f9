"""Deterministic backward (include/gstar_raster.h: gstar_set_deterministic) -- GPU tests, run late.

The reference's gradients differ from run to run (fp32 atomics in scheduling order, backward.cu:523-554) and so do this library's
by default; that noise is why the live-reference gradient bounds in test_parity_gpu.py carry a measured spread term.  With the
test mode on, every gradient of every run is BIT-IDENTICAL, on both lane layouts of the gather kernel, for plain and shared-geometry
backward passes, through the C ABI and through the operator under torch.use_deterministic_algorithms(True); the values still meet
the bound against the fp64 CPU oracle, and a view that cannot take the mode (no hit log) says so with NaN instead of quietly
summing in scheduling order.
"""
import numpy as np
import pytest
import torch

from gaustar_b200 import capi, scene

import helpers as Hh
from oracle import oracle as O
from test_parity_gpu import GRAD_TOL, SCENES, check_grads

pytestmark = pytest.mark.gpu

KEYS = ("dL_dmeans2D", "dL_dmeans3D", "dL_dopacity", "dL_dscales", "dL_drotations", "dL_dsh", "dL_dcolors", "dL_dcov3D", "dL_dconic")


@pytest.fixture()
def deterministic():
    log = capi.set_hit_log(1)
    old = capi.set_deterministic(True)
    yield
    capi.set_deterministic(old)
    capi.set_hit_log(log)


def _forward_with_log(d):
    kw = Hh.to_torch_kwargs(d)
    fwd = capi.forward(**kw)
    torch.cuda.synchronize()
    if not capi.hit_log_state(fwd)[2]:  # the log is provisioned from the previous view: the second call has it
        fwd = capi.forward(**kw)
        torch.cuda.synchronize()
    assert capi.hit_log_state(fwd)[2]
    return kw, fwd


def _same(a, b):
    return all(torch.equal(a[k], b[k]) for k in a if isinstance(a[k], torch.Tensor))


@pytest.mark.parametrize("name", ["surface_sh3", "surface_precomp", "random_big_sh2", "random_closeup_odd", "surface_mod05"])
def test_gradients_are_bit_identical_run_to_run_and_meet_the_oracle_bound(name, deterministic):
    d = SCENES[name]()
    kw, fwd = _forward_with_log(d)
    dpix = np.random.default_rng(1).normal(0, 1, (3, d["H"], d["W"])).astype(np.float32)
    dp = torch.from_numpy(dpix).cuda()
    runs = [capi.backward(fwd, dp, **Hh.bwd_kwargs(kw)) for _ in range(4)]
    torch.cuda.synchronize()
    for r in runs[1:]:
        assert _same(runs[0], r)
    # a second forward of the same view (other buffers, other hit-log addresses) gives the same bits too
    kw2, fwd2 = _forward_with_log(d)
    again = capi.backward(fwd2, dp, **Hh.bwd_kwargs(kw2))
    torch.cuda.synchronize()
    assert _same(runs[0], again)
    # the values: same bound against the fp64 oracle as the default path
    inp = Hh.oracle_inputs_from_dict(d)
    of = O.forward(inp)
    st = capi.image_state(fwd, d["W"], d["H"])
    of.n_contrib = st["n_contrib"].cpu().numpy().astype(np.uint32).reshape(of.n_contrib.shape)
    of.final_T = st["final_T"].cpu().numpy().reshape(of.final_T.shape).copy()
    check_grads(runs[0], O.backward(inp, of, dpix).__dict__, tol=GRAD_TOL)
    # and against the default (atomic) path of this library: same sums in another order
    capi.set_deterministic(False)
    plain = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    capi.set_deterministic(True)
    for k in ("dL_dmeans3D", "dL_dopacity", "dL_dmeans2D"):
        assert Hh.rel_err(plain[k].cpu(), runs[0][k].cpu()) < GRAD_TOL, k


def test_headline_size_is_bit_identical_run_to_run_while_the_default_path_is_not(deterministic):
    """1 M Gaussians at 1920x1080 (BASELINE headline): two deterministic backward runs agree in every bit; the default path, measured
    the same way, does not (that is the noise the mode removes -- if this ever fails the default path became deterministic too)."""
    g = scene.surface_gaussians(1_000_000, 3, seed=0)
    d = Hh.scene_dict(g, scene.dome_cameras(8, 1920, 1080)[5])
    kw, fwd = _forward_with_log(d)
    dp = torch.randn(3, 1080, 1920, device="cuda", generator=torch.Generator("cuda").manual_seed(2)) / (1920 * 1080)
    a = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
    b = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    assert _same(a, b)
    capi.set_deterministic(False)
    c = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
    e = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    capi.set_deterministic(True)
    assert not _same(c, e)
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dopacity"):
        assert Hh.rel_err(c[k].cpu(), a[k].cpu()) < GRAD_TOL, k


def test_view_without_hit_log_gives_nan_not_unordered_sums(deterministic):
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    capi.set_hit_log(0)
    fwd = capi.forward(**kw)
    torch.cuda.synchronize()
    assert not capi.hit_log_state(fwd)[2]
    g = capi.backward(fwd, torch.ones(3, d["H"], d["W"], device="cuda"), **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    vis = fwd["radii"] > 0
    assert torch.isnan(g["dL_dmeans3D"][vis]).all() and torch.isfinite(g["dL_dmeans3D"][~vis]).all()


def test_operator_under_torch_deterministic_algorithms():
    """torch.use_deterministic_algorithms(True) switches the mode on for the operator's backward (and off again afterwards):
    two optimisation-style fwd+bwd passes through GaussianRasterizer give identical .grad bits; so does the two-pass shared-geometry step."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    P = kw["means3D"].shape[0]
    target = torch.rand(3, kw["H"], kw["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(5))

    def settings(bg, deg):
        return dgr.GaussianRasterizationSettings(kw["H"], kw["W"], kw["tan_fovx"], kw["tan_fovy"], bg, 1.0, kw["viewmatrix"].view(4, 4), kw["projmatrix"].view(4, 4),
                                                 deg, kw["campos"], False, False)

    def step():
        leaves = {k: kw[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
        with dgr.shared_geometry():
            rgb, _ = dgr.GaussianRasterizer(settings(kw["bg"], kw["sh_degree"]))(means3D=leaves["means3D"], means2D=torch.zeros(P, 3, device="cuda"), shs=leaves["shs"],
                                                                                opacities=leaves["opacities"], scales=leaves["scales"], rotations=leaves["rotations"])
            depth_cols = (leaves["means3D"] @ kw["viewmatrix"].view(4, 4)[:3, 2:3] + kw["viewmatrix"].view(4, 4)[3, 2]).expand(-1, 3)
            dep, _ = dgr.GaussianRasterizer(settings(torch.full((3,), 10.0, device="cuda"), 0))(means3D=leaves["means3D"], means2D=torch.zeros(P, 3, device="cuda"),
                                                                                             colors_precomp=depth_cols, opacities=leaves["opacities"],
                                                                                             scales=leaves["scales"], rotations=leaves["rotations"])
        ((rgb - target).abs().mean() + 0.1 * dep.mean()).backward()
        return {k: v.grad.clone() for k, v in leaves.items()}

    step()  # provisions the hit log for this view
    was = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True)
    try:
        a, b = step(), step()
        torch.cuda.synchronize()
        assert capi.set_deterministic(-1) is True
    finally:
        torch.use_deterministic_algorithms(was)
    assert all(torch.equal(a[k], b[k]) for k in a) and all(torch.isfinite(v).all() for v in a.values())
    step()
    assert capi.set_deterministic(-1) is False  # back to the default path with the torch switch
