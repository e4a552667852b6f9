"""CPU tests of the oracle itself (no GPU): stage invariants, analytic backward vs finite differences,
edge cases, and -- the pin -- bit/tolerance agreement with the golden vectors produced by the
unmodified reference rasterizer on a B200 (tests/golden/*.npz, see tests/golden/make_golden.py)."""
import numpy as np
import pytest

from gaustar_b200 import scene
from oracle import oracle as O

import helpers as Hh


def small_scene(P=400, W=64, H=48, deg=2, seed=0, big=True):
    g = scene.random_gaussians(P, sh_degree=deg, seed=seed, scale_range=(0.02, 0.3) if big else (0.005, 0.03))
    cam = scene.look_at_camera([0.4, 1.3, 3.2], [0, 1, 0], W, H, fy_over_H=1.1)
    return Hh.scene_dict(g, cam)


def test_stage_invariants():
    d = small_scene()
    f = O.forward(Hh.oracle_inputs_from_dict(d))
    assert f.num_rendered == int(f.tiles_touched.sum()) == len(f.point_list)
    # keys ascending; ties keep gaussian-index order (stable sort, rasterizer_impl.cu:303-308)
    k = f.keys_sorted
    assert np.all(k[1:] >= k[:-1])
    same = k[1:] == k[:-1]
    assert np.all(f.point_list[1:][same] > f.point_list[:-1][same])
    # ranges partition the list by tile id
    tiles = (k >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        s, e = f.ranges[t]
        assert np.all(tiles[s:e] == t) and (s == 0 or tiles[s - 1] != t) and (e == len(k) or tiles[e] != t)
    empty = np.setdiff1d(np.arange(len(f.ranges)), np.unique(tiles))
    assert np.all(f.ranges[empty] == 0)
    # key = tile<<32 | depth bits of the gaussian
    dbits = f.depths.view(np.uint32)[f.point_list].astype(np.uint64)
    assert np.all((k & np.uint64(0xffffffff)) == dbits)
    assert np.all(f.n_contrib <= (f.ranges[:, 1] - f.ranges[:, 0]).max())
    assert np.isfinite(f.out_color).all() and (f.final_T >= 0).all() and (f.final_T <= 1).all()


def test_higher_msb_matches_reference_values():
    # rasterizer_impl.cu:35-50 -- SURVEY 8: 64 tiles -> 7 (39 bits), 8160 -> 13 (45), 5440 -> 13, 32400 -> 15 (47)
    L = O.lib()
    assert [int(L.orc_higher_msb(n)) for n in (64, 8160, 5440, 32400)] == [7, 13, 13, 15]


def test_all_culled_gives_background():
    d = small_scene(P=50)
    d["means3D"] = d["means3D"].copy()
    d["means3D"][:, 2] += 100.0  # behind the camera (camera looks down -z from z=3.2)
    f = O.forward(Hh.oracle_inputs_from_dict(d))
    assert f.num_rendered == 0 and (f.radii == 0).all()
    for c in range(3):
        assert np.all(f.out_color[c] == d["bg"][c])
    assert np.all(f.final_T == 1.0) and np.all(f.n_contrib == 0)


def test_mark_visible():
    d = small_scene(P=200)
    vis = O.mark_visible(d["means3D"], d["viewmatrix"])
    f = O.forward(Hh.oracle_inputs_from_dict(d), blend=False)
    assert np.all(vis[f.radii > 0])


def smooth_scene(use_sh):
    """One 16x16 tile, huge soft Gaussians: no tile-rect, alpha-cutoff or saturation discontinuity is
    crossed by a small parameter change, so finite differences see the same function the analytic
    backward differentiates (the reference's hard cut-offs carry no gradient)."""
    rng = np.random.default_rng(0)
    P = 16
    g = scene.random_gaussians(P, sh_degree=2, seed=4)
    g.means3D[:] = rng.uniform(-0.5, 0.5, (P, 3)).astype(np.float32) + np.array([0, 1, 0], np.float32)
    g.scales[:] = rng.uniform(3.0, 6.0, (P, 3)).astype(np.float32)
    g.opacities[:] = rng.uniform(0.05, 0.3, (P, 1)).astype(np.float32)
    cam = scene.look_at_camera([0.2, 1.2, 3.5], [0, 1, 0], 16, 16, fy_over_H=1.0)
    return Hh.scene_dict(g, cam, use_sh=use_sh)


@pytest.mark.parametrize("use_sh", [True, False])
def test_backward_matches_finite_differences(use_sh):
    """Directional derivative of L = <w, out_color> along random directions in parameter space."""
    d = smooth_scene(use_sh)
    rng = np.random.default_rng(1)
    w = rng.normal(0, 1, (3, 16, 16)).astype(np.float32)
    inp = Hh.oracle_inputs_from_dict(d)
    f = O.forward(inp)
    assert (f.radii > 0).all() and f.final_T.min() > 1e-3 and (f.n_contrib == 16).all()
    b = O.backward(inp, f, w)
    names = {"means3D": b.dL_dmeans3D, "scales": b.dL_dscales, "rotations": b.dL_drotations, "opacities": b.dL_dopacity}
    names["shs" if use_sh else "colors_precomp"] = b.dL_dsh if use_sh else b.dL_dcolors

    def loss(dd):
        return float((O.forward(Hh.oracle_inputs_from_dict(dd)).out_color.astype(np.float64) * w).sum())

    for name, grad in names.items():
        direction = rng.normal(0, 1, d[name].shape).astype(np.float32)
        direction /= np.linalg.norm(direction)
        eps = 1e-2
        dp, dm = dict(d), dict(d)
        dp[name] = (d[name] + eps * direction).astype(np.float32)
        dm[name] = (d[name] - eps * direction).astype(np.float32)
        fd = (loss(dp) - loss(dm)) / (2 * eps)
        an = float((grad.reshape(direction.shape).astype(np.float64) * direction).sum())
        assert abs(fd - an) <= 0.03 * max(abs(fd), abs(an)) + 1e-3, (name, fd, an)


@pytest.mark.parametrize("path", Hh.golden_files() or [None])
def test_oracle_matches_reference_golden(path):
    """THE PIN: oracle vs outputs of the unmodified reference CUDA rasterizer (B200)."""
    if path is None:
        pytest.fail("no golden vectors in tests/golden/ -- run tests/golden/make_golden.py on the GPU box")
    inp_d, fwd, bwd = Hh.load_golden(path)
    inp = Hh.oracle_inputs_from_dict(inp_d)
    f = O.forward(inp)
    vis = fwd["radii"] > 0
    # bit-exact: integer / key / index work
    assert f.num_rendered == int(fwd["num_rendered"])
    np.testing.assert_array_equal(f.radii, fwd["radii"])
    np.testing.assert_array_equal(f.tiles_touched.astype(np.int32), fwd["tiles_touched"])
    np.testing.assert_array_equal(f.point_offsets.astype(np.int32), fwd["point_offsets"])
    np.testing.assert_array_equal(f.depths.view(np.int32)[vis], fwd["depths"].view(np.int32)[vis])
    np.testing.assert_array_equal(f.means2D.view(np.int32)[vis], fwd["means2D"].view(np.int32)[vis])
    np.testing.assert_array_equal(f.conic_opacity.view(np.int32)[vis], fwd["conic_opacity"].view(np.int32)[vis])
    if "scales" in inp_d:
        np.testing.assert_array_equal(f.cov3D.view(np.int32)[vis], fwd["cov3D"].view(np.int32)[vis])
    np.testing.assert_array_equal(f.keys_unsorted.view(np.int64), fwd["keys_unsorted"])
    np.testing.assert_array_equal(f.values_unsorted.astype(np.int32), fwd["values_unsorted"])
    np.testing.assert_array_equal(f.keys_sorted.view(np.int64), fwd["keys_sorted"])
    np.testing.assert_array_equal(f.point_list.astype(np.int32), fwd["point_list"])
    np.testing.assert_array_equal(f.ranges.astype(np.int32), fwd["ranges"])
    np.testing.assert_array_equal(f.n_contrib.astype(np.int32), fwd["n_contrib"])
    # floating point: fp32 tolerance (CPU expf vs the GPU's ex2.approx-based expf)
    if "shs" in inp_d:
        np.testing.assert_allclose(f.rgb[vis], fwd["rgb"][vis], rtol=1e-5, atol=2e-6)
        np.testing.assert_array_equal(f.clamped[vis], fwd["clamped"][vis])
    np.testing.assert_allclose(f.out_color, fwd["out_color"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(f.final_T, fwd["final_T"], rtol=1e-5, atol=1e-6)
    # gradients: the reference sums with fp32 atomics in arbitrary order, the oracle in fp64
    b = O.backward(inp, f, inp_d["dL_dpix"])
    for k in Hh.GRAD_KEYS:
        ref = bwd[k]
        if ref.size == 0:
            continue
        got = getattr(b, k).reshape(ref.shape)
        assert Hh.rel_err(got, ref) < 2e-4, (k, Hh.rel_err(got, ref))
