"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(gaustar_b200/capi.py -> libgstar_raster.so) or through the public operator API on top of it.

Bars (SURVEY.md 8c / BASELINE.md 2.5):
  bit-exact : radii, tiles_touched, num_rendered, depth / pixel-centre / conic bits, the sorted
              (tile|depth, gaussian) list, tile ranges, n_contrib            [vs reference & golden]
  fp32 tol. : out_color, final_T: rtol 1e-5 / atol 2e-6 vs the reference (same op order; observed 0);
              atol 2e-5 vs the CPU oracle (glibc expf vs the GPU's ex2.approx-based expf)
  gradients : vs the fp64-accumulating CPU oracle and vs the golden vectors: max-abs error <= 3e-4 of the tensor's
              max magnitude, AND element-wise on every entry above 1 % of that max: the worst relative error must
              stay below max(1e-3, 2 x the worst relative error of the UNMODIFIED REFERENCE against the same oracle
              on the same entries) -- fp32 sums through 1/det^2 cannot meet a flat rtol 1e-3 entry by entry (the
              reference itself shows 1e-2 on dL_dscales), so the bar is "no worse than the reference", measured.
              vs the LIVE reference, whose fp32 atomics are order-nondeterministic: the reference is run TWICE and its
              own run-to-run spread is measured per tensor; the bound is 3e-4 + 4 x that spread -- measured, not
              argued (two runs of the unmodified reference differ by up to 1.4e-3 on dL_dcov3D: tools/grad_noise.py).
"""
import numpy as np
import pytest
import torch

from gaustar_b200 import capi, scene
from oracle import oracle as O
from oracle import refgpu

import helpers as Hh

pytestmark = pytest.mark.gpu

GRAD_TOL = 3e-4        # max-abs error / max |ref|, vs oracle and golden (hit-log backward: a few adds per Gaussian)
# The walk-back backward (no hit log) sums per-(Gaussian, warp) partials with unordered fp32 atomics like the reference does per
# pixel -- thousands of adds for a fat splat -- and shows up to 3.9e-4 on dL_dscales of the fat-splat scene; it is the
# fallback path, held to twice the bar.
GRAD_TOL_WALK = 6e-4
GRAD_RTOL_ELEM = 1e-3  # element-wise, on entries above GRAD_ELEM_FLOOR of the tensor's max
GRAD_ELEM_FLOOR = 1e-2
# dL_dcov3D / dL_dscales / dL_drotations come out of a chain with cancellation (backward.cu:278-341), and the oracle restates that
# chain in the REFERENCE's fp32 operation order (only the blend sums are fp64), so on ill-conditioned entries the oracle is one valid
# fp32 evaluation, not the true value.  Measured over 12 runs per arm (profiles/r2bb_elementwise_noise.txt): the reference's own
# worst entry moves between 7e-4 and 2.2e-3 (rotations, close-up scene), 1.1e-2 .. 1.9e-2 (scales, big splats); this library's
# between 1.7e-3 and 2.2e-3, resp. 2.5e-3 (deterministic mode) .. 1.6e-2.  A single reference run is therefore not a yardstick for
# these three tensors: the check takes the reference's worst over three runs and never asks for less than this floor.
GRAD_RTOL_ELEM_CHAIN = 5e-3
# The reference sums its blend gradients with order-nondeterministic fp32 atomics, and dL_dcov3D / dL_dscales /
# dL_drotations amplify that noise through 1/det^2 (backward.cu:201-212).  Where a test can run the reference itself
# (test_against_live_reference, the full-size tests) the bound comes from the reference's measured spread; the fixed table
# below is only for comparisons that cannot (a CUDA-graph replay against eager calls, the reference callers' own modules).
LIVE_REF_TOL = {"dL_dcov3D": 5e-3, "dL_dscales": 3e-3, "dL_drotations": 3e-3}
LIVE_SPREAD_FACTOR = 4.0


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


@pytest.fixture(params=["hitlog", "walk"], autouse=True)
def backward_path(request):
    """Every test runs twice: with the forward's hit log + instance-parallel backward (the default), and with the
    log switched off (walk-back backward, also the automatic fallback when a view's log does not fit)."""
    old = capi.set_hit_log(1 if request.param == "hitlog" else 0)
    yield request.param
    capi.set_hit_log(old)


def run_mine(d, debug=False):
    kw = Hh.to_torch_kwargs(d)
    fwd = capi.forward(debug=debug, **kw)
    torch.cuda.synchronize()
    if capi.set_hit_log(-1) == 1:
        if not capi.hit_log_state(fwd)[2]:  # first view of a new scene: the log was sized for another one; the hint is in now
            fwd = capi.forward(debug=debug, **kw)
            torch.cuda.synchronize()
        need, cap, used = capi.hit_log_state(fwd)
        assert used and 0 < need <= cap, (need, cap, used)
    else:
        assert not capi.hit_log_state(fwd)[2]
    return kw, fwd


def check_forward_against(fwd, kw, ref, exact_image, n_contrib_slack=0):
    """ref: dict-like with numpy arrays in the reference layouts."""
    P, W, H = kw["means3D"].shape[0], kw["W"], kw["H"]
    g = {k: v.cpu().numpy() for k, v in capi.unpack_geometry(fwd, P).items()}
    st = {k: v.cpu().numpy() for k, v in capi.image_state(fwd, W, H).items()}
    pl = capi.point_list(fwd).cpu().numpy()
    radii = fwd["radii"].cpu().numpy()
    vis = ref["radii"] > 0
    assert fwd["num_rendered"] == int(ref["num_rendered"])
    np.testing.assert_array_equal(radii, ref["radii"])
    np.testing.assert_array_equal(g["tiles_touched"], np.asarray(ref["tiles_touched"]).astype(np.int32))
    np.testing.assert_array_equal(bits(g["depths"])[vis], bits(ref["depths"])[vis])
    np.testing.assert_array_equal(bits(g["means2D"])[vis], bits(ref["means2D"])[vis])
    np.testing.assert_array_equal(bits(g["conic_opacity"])[vis], bits(ref["conic_opacity"])[vis])
    np.testing.assert_array_equal(pl, np.asarray(ref["point_list"]).astype(np.int32))
    np.testing.assert_array_equal(st["ranges"], np.asarray(ref["ranges"]).astype(np.int32))
    nc_ref = np.asarray(ref["n_contrib"]).astype(np.int32).reshape(-1)
    assert int((st["n_contrib"] != nc_ref).sum()) <= n_contrib_slack
    out = fwd["out_color"].cpu().numpy()
    if exact_image:
        np.testing.assert_allclose(out, ref["out_color"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(st["final_T"], np.asarray(ref["final_T"]).reshape(-1), rtol=1e-5, atol=1e-6)
    else:
        same = st["n_contrib"] == nc_ref
        assert np.abs(out - ref["out_color"]).reshape(3, -1)[:, same].max() < 2e-5
    return g, st


def elementwise_worst(m, r):
    """Worst relative error over the entries of r above GRAD_ELEM_FLOOR of its max (0.0 if there are none)."""
    r64, m64 = np.asarray(r, np.float64), np.asarray(m, np.float64).reshape(np.shape(r))
    big = np.abs(r64) > GRAD_ELEM_FLOOR * np.abs(r64).max()
    return float((np.abs(m64 - r64)[big] / np.abs(r64)[big]).max()) if big.any() else 0.0


def default_grad_tol():
    return GRAD_TOL if capi.set_hit_log(-1) == 1 else GRAD_TOL_WALK


def check_grads(mine, ref, tol=None, per_key=None, elementwise_against=None):
    """max-norm bound per tensor (per_key raises it for single tensors).  elementwise_against: the unmodified reference's gradients
    for the same inputs (one dict, or several runs of it) -- our worst element-wise relative error (entries above 1 % of the max) vs
    `ref` must stay below max(GRAD_RTOL_ELEM, 2 x the reference's own worst over those runs); for the three tensors behind the
    cov3D -> scale / rotation chain the floor is GRAD_RTOL_ELEM_CHAIN (see there)."""
    for k in Hh.GRAD_KEYS:
        r = np.asarray(ref[k])
        if r.size == 0:
            continue
        m = mine[k].cpu().numpy().reshape(r.shape)
        assert np.isfinite(m).all(), k
        bound = max(default_grad_tol() if tol is None else tol, (per_key or {}).get(k, 0.0))
        assert Hh.rel_err(m, r) < bound, (k, Hh.rel_err(m, r), bound)
        if elementwise_against is not None:
            runs = elementwise_against if isinstance(elementwise_against, (list, tuple)) else [elementwise_against]
            ours, theirs = elementwise_worst(m, r), max(elementwise_worst(np.asarray(e[k]), r) for e in runs)
            floor = GRAD_RTOL_ELEM_CHAIN if k in ("dL_dcov3D", "dL_dscales", "dL_drotations") else GRAD_RTOL_ELEM
            assert ours < max(floor, 2.0 * theirs), (k, "element-wise", ours, "reference:", theirs)


def reference_spread(ref, dpix, bkw, runs=3):
    """Per-tensor run-to-run spread of the UNMODIFIED reference's backward on identical inputs (max-abs difference between two
    runs over the tensor's max): what its atomics' order costs.  Returns (gradients of the first run, {key: spread})."""
    first = {k: v.cpu().numpy() for k, v in refgpu.backward(ref, dpix, **bkw).items()}
    spread = {k: 0.0 for k in first}
    for _ in range(runs - 1):
        again = refgpu.backward(ref, dpix, **bkw)
        for k, v in again.items():
            if first[k].size:
                spread[k] = max(spread[k], Hh.rel_err(v.cpu().numpy(), first[k]))
    return first, spread


def live_bound(spread):
    """Per-tensor bound against ONE run of the live reference: our own fp32 error budget (GRAD_TOL, the bar against the
    deterministic oracle) plus LIVE_SPREAD_FACTOR x the reference's measured run-to-run spread."""
    return {k: default_grad_tol() + LIVE_SPREAD_FACTOR * v for k, v in spread.items()}


SCENES = {
    "surface_sh3": lambda: Hh.scene_dict(scene.surface_gaussians(20000, 3, seed=1), scene.dome_cameras(6, 400, 225)[2]),
    "surface_precomp": lambda: Hh.scene_dict(scene.surface_gaussians(12000, 3, seed=2), scene.dome_cameras(6, 320, 200)[4], use_sh=False,
                                             bg=(10.0, 10.0, 10.0)),
    "random_big_sh2": lambda: Hh.scene_dict(scene.random_gaussians(5000, 2, seed=5, scale_range=(0.01, 0.4)),
                                            scene.look_at_camera([0.5, 1.3, 4.0], [0, 1, 0], 320, 180, fy_over_H=1.2)),
    "random_closeup_odd": lambda: Hh.scene_dict(scene.random_gaussians(4000, 1, seed=6), scene.look_at_camera([0.2, 1.0, 0.9], [0, 1, 0], 333, 201, fy_over_H=0.9),
                                                bg=(0.3, 0.2, 0.1)),
    "sh_deg1_of_3": lambda: Hh.scene_dict(scene.random_gaussians(3000, 3, seed=8, scale_range=(0.01, 0.1)),
                                          scene.look_at_camera([0.0, 1.0, 3.0], [0, 1, 0], 160, 120), sh_degree=1),
    # inputs the callers can reach (round 2): render(..., scaling_modifier) (gaussian_renderer/__init__.py:18,44), an off-centre
    # principal point (sugar_model.py:1160-1161), campos of shape [1,3] on the SH path (sugar_model.py:1164: get_camera_center())
    "surface_mod05": lambda: dict(Hh.scene_dict(scene.surface_gaussians(16000, 3, seed=3), scene.dome_cameras(6, 384, 216)[3]), scale_modifier=0.5),
    "random_mod2": lambda: dict(Hh.scene_dict(scene.random_gaussians(4000, 2, seed=12, scale_range=(0.01, 0.15)),
                                              scene.look_at_camera([0.4, 1.2, 3.6], [0, 1, 0], 320, 180, fy_over_H=1.2)), scale_modifier=2.0),
    "surface_offcentre": lambda: Hh.scene_dict(scene.surface_gaussians(16000, 3, seed=6),
                                               scene.look_at_camera([1.9, 1.6, 2.2], [0, 1, 0], 400, 225, principal_ndc=(0.31, -0.22))),
    "surface_campos_1x3": lambda: (lambda d: dict(d, campos=d["campos"].reshape(1, 3)))(
        Hh.scene_dict(scene.surface_gaussians(9000, 3, seed=7), scene.dome_cameras(6, 320, 180)[0])),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_against_cpu_oracle(name):
    d = SCENES[name]()
    kw, fwd = run_mine(d)
    inp = Hh.oracle_inputs_from_dict(d)
    of = O.forward(inp)
    ref = dict(of.__dict__)
    npix = d["W"] * d["H"]
    g, st = check_forward_against(fwd, kw, ref, exact_image=False, n_contrib_slack=max(2, npix // 20000))
    if "shs" in d:
        vis = of.radii > 0
        np.testing.assert_allclose(g["rgb"][vis], of.rgb[vis], rtol=1e-5, atol=2e-6)
    dpix = np.random.default_rng(1).normal(0, 1, (3, d["H"], d["W"])).astype(np.float32)
    mine = capi.backward(fwd, torch.from_numpy(dpix).cuda(), **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    # oracle backward on the GPU's own forward state so that a flipped n_contrib does not count as a gradient error
    of.n_contrib = st["n_contrib"].astype(np.uint32)
    of.final_T = st["final_T"].copy()
    ob = O.backward(inp, of, dpix)
    ref_grads = None
    if refgpu.available():  # the reference on the same inputs, three runs (it is not reproducible): the yardstick of the element-wise check
        rf = refgpu.forward(**kw)
        ref_grads = [{k: v.cpu().numpy() for k, v in refgpu.backward(rf, torch.from_numpy(dpix).cuda(), **Hh.bwd_kwargs(kw)).items()} for _ in range(3)]
    check_grads(mine, ob.__dict__, elementwise_against=ref_grads)


@pytest.mark.parametrize("path", Hh.golden_files() or [None])
def test_against_reference_golden(path):
    if path is None:
        pytest.fail("no golden vectors in tests/golden/")
    inp_d, rf, rb = Hh.load_golden(path)
    d = {k: v for k, v in inp_d.items() if k != "dL_dpix"}
    kw, fwd = run_mine(d)
    check_forward_against(fwd, kw, rf, exact_image=True)
    mine = capi.backward(fwd, torch.from_numpy(inp_d["dL_dpix"]).cuda(), **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    # (the golden gradients are ONE run of the reference and carry its atomics' noise: max-norm bound only)
    check_grads(mine, rb)


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref not built (reference sources absent at build time)")
@pytest.mark.parametrize("name", sorted(SCENES))
def test_against_live_reference(name):
    d = SCENES[name]()
    kw, fwd = run_mine(d)
    ref = refgpu.forward(**kw)
    rnp = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in ref.items()}
    check_forward_against(fwd, kw, rnp, exact_image=True)
    dpix = torch.randn(3, d["H"], d["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    mine = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
    rg, spread = reference_spread(ref, dpix, Hh.bwd_kwargs(kw))
    check_grads(mine, rg, per_key=live_bound(spread))


def test_operator_api_autograd_matches_cabi():
    """GaussianRasterizer + autograd (the call sugar_model.py:1285-1293 makes) == direct C-ABI calls,
    including non-contiguous expanded colours (refine.py:605) and the means2D gradient the densifier reads."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_precomp"]()
    kw = Hh.to_torch_kwargs(d)
    P = kw["means3D"].shape[0]
    depth_like = torch.rand(P, 1, device="cuda")
    colors = depth_like.expand(-1, 3)  # non-contiguous
    leaf = {k: kw[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations")}
    col_leaf = depth_like.clone().requires_grad_(True)
    means2D = torch.zeros(P, 3, device="cuda", requires_grad=True)
    rs = dgr.GaussianRasterizationSettings(image_height=kw["H"], image_width=kw["W"], tanfovx=kw["tan_fovx"], tanfovy=kw["tan_fovy"], bg=kw["bg"],
                                           scale_modifier=1.0, viewmatrix=kw["viewmatrix"].view(4, 4), projmatrix=kw["projmatrix"].view(4, 4), sh_degree=0,
                                           campos=kw["campos"].view(1, 3), prefiltered=False, debug=False)
    img, radii = dgr.GaussianRasterizer(rs)(means3D=leaf["means3D"], means2D=means2D, opacities=leaf["opacities"], colors_precomp=col_leaf.expand(-1, 3),
                                           scales=leaf["scales"], rotations=leaf["rotations"])
    assert img.shape == (3, kw["H"], kw["W"]) and radii.dtype == torch.int32 and radii.shape == (P,)
    w = torch.randn_like(img)
    (img * w).sum().backward()
    kw2 = dict(kw); kw2["colors_precomp"] = colors.contiguous()
    fwd = capi.forward(**kw2)
    g = capi.backward(fwd, w, **Hh.bwd_kwargs(kw2))
    torch.cuda.synchronize()
    assert torch.equal(img, fwd["out_color"]) and torch.equal(radii, fwd["radii"])
    assert Hh.rel_err(leaf["means3D"].grad.cpu(), g["dL_dmeans3D"].cpu()) < GRAD_TOL
    assert Hh.rel_err(leaf["scales"].grad.cpu(), g["dL_dscales"].cpu()) < GRAD_TOL
    assert Hh.rel_err(leaf["rotations"].grad.cpu(), g["dL_drotations"].cpu()) < GRAD_TOL
    assert Hh.rel_err(leaf["opacities"].grad.cpu(), g["dL_dopacity"].cpu()) < GRAD_TOL
    assert Hh.rel_err(means2D.grad.cpu(), g["dL_dmeans2D"].cpu()) < GRAD_TOL
    assert Hh.rel_err(col_leaf.grad.cpu(), g["dL_dcolors"].sum(1, keepdim=True).cpu()) < GRAD_TOL


def test_operator_api_with_shs_like_gaussian_renderer():
    """gaussian_renderer.render() passes shs=[P,16,3] (gaussian_renderer/__init__.py:85-93): dL_dsh flows."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    shs = kw["shs"].clone().requires_grad_(True)
    m3 = kw["means3D"].clone().requires_grad_(True)
    rs = dgr.GaussianRasterizationSettings(kw["H"], kw["W"], kw["tan_fovx"], kw["tan_fovy"], kw["bg"], 1.0, kw["viewmatrix"].view(4, 4),
                                           kw["projmatrix"].view(4, 4), 3, kw["campos"], False, True)  # debug=True: per-stage sync+check path
    img, radii = dgr.GaussianRasterizer(rs)(means3D=m3, means2D=torch.zeros_like(m3, requires_grad=True), opacities=kw["opacities"], shs=shs,
                                           scales=kw["scales"], rotations=kw["rotations"])
    img.square().sum().backward()
    assert shs.grad.shape == shs.shape and torch.isfinite(shs.grad).all() and shs.grad.abs().max() > 0
    assert torch.isfinite(m3.grad).all()
    assert (shs.grad[radii == 0] == 0).all()


def test_edge_empty_and_culled():
    import diff_gaussian_rasterization as dgr
    dev = "cuda"
    rs = dgr.GaussianRasterizationSettings(40, 56, 0.5, 0.5, torch.tensor([0.1, 0.2, 0.3], device=dev), 1.0, torch.eye(4, device=dev),
                                           torch.eye(4, device=dev), 0, torch.zeros(3, device=dev), False, False)
    r = dgr.GaussianRasterizer(rs)
    # P == 0: zeros image (rasterize_points.cu:66,81), empty radii
    z3, z4, z1 = torch.zeros(0, 3, device=dev), torch.zeros(0, 4, device=dev), torch.zeros(0, 1, device=dev)
    img, radii = r(z3, z3, z1, colors_precomp=z3, scales=z3, rotations=z4)
    assert img.shape == (3, 40, 56) and float(img.abs().max()) == 0.0 and radii.numel() == 0
    # every Gaussian behind the near plane: image == background, no instances, zero grads
    P = 100
    m = torch.randn(P, 3, device=dev)
    m[:, 2] = -5.0
    m.requires_grad_(True)
    rot = torch.zeros(P, 4, device=dev); rot[:, 0] = 1
    img, radii = r(m, torch.zeros_like(m), torch.full((P, 1), 0.5, device=dev), colors_precomp=torch.rand(P, 3, device=dev),
                   scales=torch.full((P, 3), 0.1, device=dev), rotations=rot)
    assert (radii == 0).all()
    assert torch.allclose(img, torch.tensor([0.1, 0.2, 0.3], device=dev).view(3, 1, 1).expand_as(img))
    img.sum().backward()
    assert float(m.grad.abs().max()) == 0.0


def test_mark_visible():
    d = SCENES["random_closeup_odd"]()
    kw = Hh.to_torch_kwargs(d)
    got = capi.mark_visible(kw["means3D"], kw["viewmatrix"], kw["projmatrix"]).cpu().numpy()
    np.testing.assert_array_equal(got, O.mark_visible(d["means3D"], d["viewmatrix"]))
    assert 0 < got.sum() < len(got)


@pytest.mark.parametrize("P,size", [(7000, 32), (40000, 16)])
def test_long_tile_lists_sort_paths(P, size):
    """Tile lists longer than the small sort kernel (4096) and longer than shared memory (24576)."""
    g = scene.random_gaussians(P, 0, seed=9, scale_range=(0.3, 0.6), extent=0.3)
    g.opacities[:] = 0.02
    cam = scene.look_at_camera([0.0, 1.0, 3.0], [0, 1, 0], size, size, fy_over_H=1.0)
    d = Hh.scene_dict(g, cam, use_sh=False)
    kw, fwd = run_mine(d)
    of = O.forward(Hh.oracle_inputs_from_dict(d))
    assert (of.ranges[:, 1] - of.ranges[:, 0]).max() > (4096 if P < 20000 else 24576)
    check_forward_against(fwd, kw, dict(of.__dict__), exact_image=False, n_contrib_slack=4)


def test_capacity_regrow_path():
    """A call whose instance count exceeds the provision made from the previous call must still be exact."""
    small = SCENES["sh_deg1_of_3"]()
    big = SCENES["random_big_sh2"]()
    for d in (small, big, small, big):
        kw, fwd = run_mine(d)
        of = O.forward(Hh.oracle_inputs_from_dict(d), blend=False)
        assert fwd["num_rendered"] == of.num_rendered
        np.testing.assert_array_equal(capi.point_list(fwd).cpu().numpy(), of.point_list.astype(np.int32))


def test_full_size_properties():
    """BASELINE headline size (1M Gaussians, 1920x1080): size-independent properties."""
    g = scene.surface_gaussians(1_000_000, 3, seed=0)
    cam = scene.dome_cameras(8, 1920, 1080)[5]
    d = Hh.scene_dict(g, cam)
    kw, fwd = run_mine(d)
    P, W, H = g.P, 1920, 1080
    geo = capi.unpack_geometry(fwd, P)
    st = capi.image_state(fwd, W, H)
    pl = capi.point_list(fwd).long()
    R = fwd["num_rendered"]
    assert int(geo["tiles_touched"].sum()) == R == pl.numel()
    rng = st["ranges"].long()
    n = rng[:, 1] - rng[:, 0]
    assert int(n.sum()) == R
    nz = n > 0
    order = rng[nz, 0].argsort()
    starts, lens = rng[nz, 0][order], n[nz][order]
    tile_ids = torch.arange(rng.shape[0], device="cuda")[nz][order]
    assert int(starts[0]) == 0 and torch.equal(starts[1:], (starts + lens)[:-1]) and int((starts + lens)[-1]) == R  # ranges tile the list
    assert bool((tile_ids[1:] > tile_ids[:-1]).all())  # in ascending tile order
    # per-tile lists are sorted by (depth bits, index): check globally with the tile id of every instance
    tile_of = torch.repeat_interleave(tile_ids, lens)
    key = (tile_of << 32) | geo["depths"].view(torch.int32)[pl].long()
    assert bool((key[1:] >= key[:-1]).all())
    tie = key[1:] == key[:-1]
    assert bool((pl[1:][tie] > pl[:-1][tie]).all())
    # every instance lies in a tile of its Gaussian's rect; counts per Gaussian match tiles_touched
    assert torch.equal(torch.bincount(pl, minlength=P).int(), geo["tiles_touched"])
    out = fwd["out_color"]
    assert torch.isfinite(out).all() and float(st["final_T"].min()) >= 0 and float(st["final_T"].max()) <= 1
    # idempotence: a second call is bit-identical (deterministic forward)
    _, fwd2 = run_mine(d)
    assert torch.equal(fwd2["out_color"], out) and torch.equal(capi.point_list(fwd2).long(), pl)
    # linearity of the backward in the upstream gradient
    dpix = torch.randn(3, H, W, device="cuda") / (W * H)
    g1 = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
    g2 = capi.backward(fwd, 2.0 * dpix, **Hh.bwd_kwargs(kw))
    for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity", "dL_dscales"):
        assert torch.isfinite(g1[k]).all()
        assert Hh.rel_err((2.0 * g1[k]).cpu(), g2[k].cpu()) < 1e-4, k


def test_accumulate_param_grads_equals_sum_of_views():
    """accumulate_param_grads=1 (multi-view steps): K8 adds into the flat buffer == sum of per-view gradients."""
    from gaustar_b200 import dist as gdist
    g = scene.surface_gaussians(15000, 3, seed=4)
    cams = scene.dome_cameras(6, 256, 160)
    P, M = g.P, g.shs.shape[1]
    flat = gdist.FlatGrads(P, M, "cuda")
    ref = gdist.FlatGrads(P, M, "cuda")
    for ci in (1, 3, 4):
        kw = Hh.to_torch_kwargs(Hh.scene_dict(g, cams[ci]))
        fwd = capi.forward(**kw)
        dpix = torch.randn(3, 160, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(ci))
        plain = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
        ref.accumulate(plain)
        fused = capi.backward(fwd, dpix, accumulate_into=flat.views, **Hh.bwd_kwargs(kw))
        assert fused["dL_dsh"].data_ptr() == flat.views["dL_dsh"].data_ptr()
    torch.cuda.synchronize()
    for k in gdist.GRAD_FIELDS:
        assert Hh.rel_err(flat.views[k].cpu(), ref.views[k].cpu()) < 1e-5, k


@pytest.mark.parametrize("sh_degree", [3, 2, 0])
def test_atomic_accumulate_from_two_streams_into_one_buffer(sh_degree):
    """accumulate_param_grads=2: backward passes of different views running at the same time on two CUDA streams add into ONE flat
    buffer with reductions at L2 (bench.py's value arm; SH rows of 48 / 27 floats take the vector / scalar store paths, degree 0
    with colors_precomp the no-SH path) == the sum of the per-view gradients."""
    from gaustar_b200 import dist as gdist
    g = scene.surface_gaussians(40000, max(sh_degree, 1), seed=4)
    cams = scene.dome_cameras(8, 320, 200)
    P, M = g.P, (g.shs.shape[1] if sh_degree else 0)
    flat, ref = gdist.FlatGrads(P, M, "cuda"), gdist.FlatGrads(P, M, "cuda")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    views = []
    for ci in range(8):
        kw = Hh.to_torch_kwargs(Hh.scene_dict(g, cams[ci], use_sh=sh_degree > 0))
        fwd = capi.forward(**kw)
        dpix = torch.randn(3, 200, 320, device="cuda", generator=torch.Generator("cuda").manual_seed(ci))
        plain = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
        plain = dict(plain, dL_dsh=plain["dL_dsh"] if M else torch.zeros(P, 0, 3, device="cuda"))
        ref.accumulate(plain)
        views.append((kw, fwd, dpix))
    torch.cuda.synchronize()
    for rep in range(3):  # repeated: a lost update would not show every time
        flat.zero_()
        torch.cuda.synchronize()
        for i, (kw, fwd, dpix) in enumerate(views):
            with torch.cuda.stream(streams[i & 1]):
                capi.backward(fwd, dpix, accumulate_into=flat.views, atomic_accumulate=True, lean=True, **Hh.bwd_kwargs(kw))
        torch.cuda.synchronize()
        for k in gdist.GRAD_FIELDS:
            if flat.views[k].numel():
                assert Hh.rel_err(flat.views[k].cpu(), ref.views[k].cpu()) < 2e-5, (k, rep)


def test_two_streams_autograd_equals_single_stream():
    """bench.py's e2e arm keeps two views in flight on two CUDA streams with one set of autograd leaves per stream
    (same storage, separate .grad buffers).  The summed gradients must equal the single-stream result."""
    import diff_gaussian_rasterization as dgr
    from gaustar_b200 import dist as gdist
    g = scene.surface_gaussians(12000, 3, seed=7)
    cams = scene.dome_cameras(6, 320, 192)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    base = {"means3D": t(g.means3D), "scales": t(g.scales), "rotations": t(g.rotations), "opacities": t(g.opacities), "shs": t(g.shs)}
    name_of = {"means3D": "dL_dmeans3D", "scales": "dL_dscales", "rotations": "dL_drotations", "opacities": "dL_dopacity", "shs": "dL_dsh"}
    bg = torch.tensor([0.0, 1.0, 0.0], device="cuda")
    targets = [torch.rand(3, 192, 320, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) for i in range(4)]

    def render_loss(ls, cam, tgt):
        rs = dgr.GaussianRasterizationSettings(192, 320, cam.tanfovx, cam.tanfovy, bg, 1.0, t(cam.viewmatrix), t(cam.projmatrix), 3, t(cam.campos), False, False)
        img, _ = dgr.GaussianRasterizer(rs)(means3D=ls["means3D"], means2D=torch.zeros_like(ls["means3D"], requires_grad=True), opacities=ls["opacities"],
                                            shs=ls["shs"], scales=ls["scales"], rotations=ls["rotations"])
        torch.nn.functional.l1_loss(img, tgt).backward()

    def make_leaves(flat):
        ls = {k: v.detach().requires_grad_(True) for k, v in base.items()}
        for k, p_ in ls.items():
            p_.grad = flat.views[name_of[k]]
        return ls

    ref = gdist.FlatGrads(g.P, 16, "cuda")
    ls = make_leaves(ref)
    for i in range(4):
        render_loss(ls, cams[i], targets[i])
    torch.cuda.synchronize()

    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    flats = [gdist.FlatGrads(g.P, 16, "cuda"), gdist.FlatGrads(g.P, 16, "cuda")]
    sets = []
    for st, fl in zip(streams, flats):
        with torch.cuda.stream(st):
            sets.append(make_leaves(fl))
    torch.cuda.synchronize()
    for rep in range(3):  # repeat: a race would show up as run-to-run differences
        for fl in flats:
            fl.zero_()
        torch.cuda.synchronize()
        for i in range(4):
            with torch.cuda.stream(streams[i & 1]):
                render_loss(sets[i & 1], cams[i], targets[i])
        torch.cuda.synchronize()
        total = flats[0].flat + flats[1].flat
        assert Hh.rel_err(total.cpu(), ref.flat.cpu()) < 1e-4, rep

    # the same with ONE buffer: both streams' leaves share their .grad tensors and the kernel adds with reductions at L2
    # (set_grad_accumulation_fusion(True, atomic=True): what bench.py's e2e arm does)
    import gaustar_b200
    one = gdist.FlatGrads(g.P, 16, "cuda")
    shared_sets = []
    for st in streams:
        with torch.cuda.stream(st):
            shared_sets.append(make_leaves(one))
    torch.cuda.synchronize()
    old = gaustar_b200.set_grad_accumulation_fusion(True, atomic=True)
    try:
        for rep in range(3):
            one.zero_()
            torch.cuda.synchronize()
            for i in range(4):
                with torch.cuda.stream(streams[i & 1]):
                    render_loss(shared_sets[i & 1], cams[i], targets[i])
            torch.cuda.synchronize()
            assert Hh.rel_err(one.flat.cpu(), ref.flat.cpu()) < 1e-4, rep
    finally:
        gaustar_b200.set_grad_accumulation_fusion(old)


def test_grad_accumulation_fusion_matches_autograd():
    """set_grad_accumulation_fusion(True): leaves with preallocated .grad receive the views' gradients inside the kernel
    (the function returns None for them); the result equals ordinary autograd accumulation, and a call that does not
    qualify (non-leaf input) silently takes the ordinary path."""
    import diff_gaussian_rasterization as dgr
    import gaustar_b200
    g = scene.surface_gaussians(9000, 3, seed=11)
    cams = scene.dome_cameras(5, 256, 160)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    base = {"means3D": t(g.means3D), "scales": t(g.scales), "rotations": t(g.rotations), "opacities": t(g.opacities), "shs": t(g.shs)}
    bg = torch.tensor([0.0, 1.0, 0.0], device="cuda")
    tg = [torch.rand(3, 160, 256, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) for i in range(3)]

    def run(fused, scale_inputs=False):
        old = gaustar_b200.set_grad_accumulation_fusion(fused)
        try:
            ls = {k: v.detach().clone().requires_grad_(True) for k, v in base.items()}
            for p_ in ls.values():
                p_.grad = torch.zeros_like(p_)
            m2d = torch.zeros(g.P, 3, device="cuda", requires_grad=True)
            for i in range(3):
                cam = cams[i + 1]
                rs = dgr.GaussianRasterizationSettings(160, 256, cam.tanfovx, cam.tanfovy, bg, 1.0, t(cam.viewmatrix), t(cam.projmatrix), 3, t(cam.campos),
                                                       False, False)
                m3 = ls["means3D"] * 1.0 if scale_inputs else ls["means3D"]  # non-leaf input: must fall back
                img, _ = dgr.GaussianRasterizer(rs)(means3D=m3, means2D=m2d, opacities=ls["opacities"], shs=ls["shs"], scales=ls["scales"],
                                                    rotations=ls["rotations"])
                torch.nn.functional.l1_loss(img, tg[i]).backward()
            torch.cuda.synchronize()
            return {k: v.grad.clone() for k, v in ls.items()}, m2d.grad.clone()
        finally:
            gaustar_b200.set_grad_accumulation_fusion(old)

    ref, ref2d = run(False)
    fus, fus2d = run(True)
    mixed, _ = run(True, scale_inputs=True)
    for k in ref:
        assert float(ref[k].abs().max()) > 0
        assert Hh.rel_err(fus[k].cpu(), ref[k].cpu()) < 1e-4, k
        assert Hh.rel_err(mixed[k].cpu(), ref[k].cpu()) < 1e-4, k
    assert Hh.rel_err(fus2d.cpu(), ref2d.cpu()) < 1e-4


def test_hit_log_too_small_falls_back_on_device():
    """A view whose hit log does not fit the provision made from the previous (smaller) view must take the walk-back
    backward by itself -- decided on the device, no host round trip -- and still be exact; the next call has a log again."""
    old = capi.set_hit_log(1)
    try:
        small = SCENES["sh_deg1_of_3"]()
        big = SCENES["random_big_sh2"]()
        for _ in range(40):  # the provision shrinks only after 32 consecutive small views
            fs = capi.forward(**Hh.to_torch_kwargs(small))
        torch.cuda.synchronize()
        assert capi.hit_log_state(fs)[2]
        kw = Hh.to_torch_kwargs(big)
        fwd = capi.forward(**kw)
        torch.cuda.synchronize()
        need, cap, used = capi.hit_log_state(fwd)
        assert need > cap and not used, (need, cap, used)
        inp = Hh.oracle_inputs_from_dict(big)
        of = O.forward(inp)
        g, st = check_forward_against(fwd, kw, dict(of.__dict__), exact_image=False, n_contrib_slack=4)
        dpix = np.random.default_rng(5).normal(0, 1, (3, big["H"], big["W"])).astype(np.float32)
        mine = capi.backward(fwd, torch.from_numpy(dpix).cuda(), **Hh.bwd_kwargs(kw))
        torch.cuda.synchronize()
        of.n_contrib = st["n_contrib"].astype(np.uint32)
        of.final_T = st["final_T"].copy()
        check_grads(mine, O.backward(inp, of, dpix).__dict__, tol=GRAD_TOL_WALK)  # (this view took the walk-back backward)
        fwd2 = capi.forward(**kw)
        torch.cuda.synchronize()
        assert capi.hit_log_state(fwd2)[2]  # re-provisioned from the need the first call published
    finally:
        capi.set_hit_log(old)


def test_forward_only_skips_hit_log_and_stays_exact():
    """forward_only (set automatically by the operator when no input requires grad): same image bit for bit, no hit log;
    a backward on those buffers still works (walk-back kernel)."""
    import diff_gaussian_rasterization as dgr
    old = capi.set_hit_log(1)
    try:
        d = SCENES["surface_sh3"]()
        kw = Hh.to_torch_kwargs(d)
        a = capi.forward(**kw)
        b = capi.forward(forward_only=True, **kw)
        torch.cuda.synchronize()
        assert torch.equal(a["out_color"], b["out_color"]) and not capi.hit_log_state(b)[2]
        dpix = torch.randn(3, d["H"], d["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(4))
        ga, gb = capi.backward(a, dpix, **Hh.bwd_kwargs(kw)), capi.backward(b, dpix, **Hh.bwd_kwargs(kw))
        for k in ("dL_dmeans3D", "dL_dsh", "dL_dopacity"):
            assert Hh.rel_err(gb[k].cpu(), ga[k].cpu()) < GRAD_TOL, k
        rs = dgr.GaussianRasterizationSettings(d["H"], d["W"], kw["tan_fovx"], kw["tan_fovy"], kw["bg"], 1.0, kw["viewmatrix"].view(4, 4),
                                               kw["projmatrix"].view(4, 4), 3, kw["campos"], False, False)
        with torch.no_grad():
            img, _ = dgr.GaussianRasterizer(rs)(means3D=kw["means3D"], means2D=torch.zeros_like(kw["means3D"]), opacities=kw["opacities"],
                                                shs=kw["shs"], scales=kw["scales"], rotations=kw["rotations"])
        assert torch.equal(img, a["out_color"])
    finally:
        capi.set_hit_log(old)


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref not built (reference sources absent at build time)")
@pytest.mark.parametrize("shape", ["config2_200k_1080p", "config3_res_1352x1014"])
def test_baseline_config_shapes_against_live_reference(shape):
    """BASELINE.json configs #2 (200 k Gaussians, 1920x1080) and #3's resolution (1352x1014: 84.5 x 63.4 tiles, i.e.
    partial tiles on both borders; 150 k Gaussians here) against the unmodified reference on the same inputs."""
    if shape == "config2_200k_1080p":
        d = Hh.scene_dict(scene.surface_gaussians(200000, 3, seed=0), scene.dome_cameras(16, 1920, 1080)[3])
    else:
        d = Hh.scene_dict(scene.surface_gaussians(150000, 3, seed=3), scene.dome_cameras(20, 1352, 1014)[7])
    kw, fwd = run_mine(d)
    ref = refgpu.forward(**kw)
    rnp = {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in ref.items()}
    check_forward_against(fwd, kw, rnp, exact_image=True)
    dpix = torch.randn(3, d["H"], d["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(6)) / (d["H"] * d["W"])
    mine = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
    rg, spread = reference_spread(ref, dpix, Hh.bwd_kwargs(kw))
    check_grads(mine, rg, per_key=live_bound(spread))


def test_prefiltered_flag_changes_nothing_when_every_gaussian_is_in_front():
    """prefiltered=True is a promise that no Gaussian is behind the near plane (auxiliary.h:154-160 traps on a culled point);
    with every Gaussian in front the outputs equal the prefiltered=False call bit for bit."""
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    a = capi.forward(prefiltered=False, **kw)
    b = capi.forward(prefiltered=True, **kw)
    torch.cuda.synchronize()
    assert a["num_rendered"] == b["num_rendered"] and torch.equal(a["radii"], b["radii"]) and torch.equal(a["out_color"], b["out_color"])


# BASELINE.json's configurations at FULL size against the live reference (SURVEY 8d "configs restated"): the headline
# (1 M surface Gaussians, 1920x1080, SH3), #3 (1 M, 1352x1014, SH3), #4 (500 k, 1352x1014, GauSTAR's sh_levels = 3 -> M = 9),
# #5 (4 M, 3840x2160: 32 400 tiles, 47 sort bits; RGB through SH, then a colors_precomp pass as its depth / normal renders do).
FULL_CONFIGS = {
    "headline_1M_1080p_sh3": dict(P=1_000_000, W=1920, H=1080, deg=3, cam=5, precomp=False),
    "config3_1M_1352x1014_sh3": dict(P=1_000_000, W=1352, H=1014, deg=3, cam=2, precomp=False),
    "config4_500k_1352x1014_sh2": dict(P=500_000, W=1352, H=1014, deg=2, cam=7, precomp=False),
    "config5_4M_4k_sh3": dict(P=4_000_000, W=3840, H=2160, deg=3, cam=1, precomp=False),
    "config5_4M_4k_precomp": dict(P=4_000_000, W=3840, H=2160, deg=3, cam=4, precomp=True),
}


@pytest.mark.skipif(not refgpu.available(), reason="oracle/_ref not built (reference sources absent at build time)")
@pytest.mark.parametrize("name", sorted(FULL_CONFIGS))
def test_full_size_against_live_reference(name, backward_path):
    """Every integer / key quantity and the image bit-identical to the unmodified reference at the full size of the configuration;
    gradients within GRAD_TOL + LIVE_SPREAD_FACTOR x the reference's own measured run-to-run spread."""
    c = FULL_CONFIGS[name]
    if backward_path == "walk" and c["P"] > 1_000_000:
        pytest.skip("the walk-back backward is covered at 1 M; 4 M runs once (memory and time)")
    g = scene.surface_gaussians(c["P"], c["deg"], seed=0)
    cam = scene.dome_cameras(8, c["W"], c["H"])[c["cam"]]
    d = Hh.scene_dict(g, cam, use_sh=not c["precomp"], bg=(10.0, 10.0, 10.0) if c["precomp"] else (0.0, 1.0, 0.0))
    del g
    kw, fwd = run_mine(d)
    ref = refgpu.forward(**kw)
    P, W, H = kw["means3D"].shape[0], kw["W"], kw["H"]
    geo = capi.unpack_geometry(fwd, P)
    st = capi.image_state(fwd, W, H)
    vis = ref["radii"] > 0
    assert fwd["num_rendered"] == ref["num_rendered"] and ref["num_rendered"] > P
    assert torch.equal(fwd["radii"], ref["radii"])
    assert torch.equal(geo["tiles_touched"].int(), ref["tiles_touched"].int())
    for k in ("depths", "means2D", "conic_opacity"):
        assert torch.equal(geo[k].contiguous().view(torch.int32)[vis], ref[k].contiguous().view(torch.int32)[vis]), k
    assert torch.equal(capi.point_list(fwd).int(), ref["point_list"].int())
    assert torch.equal(st["ranges"].int().reshape(-1), ref["ranges"].int().reshape(-1))
    assert torch.equal(st["n_contrib"].int().reshape(-1), ref["n_contrib"].int().reshape(-1))
    assert torch.equal(fwd["out_color"], ref["out_color"])
    assert torch.equal(st["final_T"].reshape(-1), ref["final_T"].reshape(-1))
    del geo, st
    dpix = torch.randn(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(2)) / (W * H)
    mine = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
    torch.cuda.synchronize()
    rg, spread = reference_spread(ref, dpix, Hh.bwd_kwargs(kw))
    check_grads(mine, rg, per_key=live_bound(spread))
