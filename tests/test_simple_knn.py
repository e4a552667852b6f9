"""simple_knn._C.distCUDA2 shim (SURVEY 8b co-requisite): import surface on CPU, values against the oracle on the GPU."""
import numpy as np
import pytest
import torch


def test_import_surface_and_cpu_refusal():
    from simple_knn._C import distCUDA2  # the import GauSTAR does at module load (sugar_model.py:9, gaussian_model.py:20)
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(4, 3))


def test_oracle_known_answers():
    from oracle import knn_oracle as K
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [10, 10, 10]], np.float32)
    got = K.dist2_mean3(pts)
    np.testing.assert_allclose(got[0], (1 + 4 + 9) / 3.0, rtol=1e-6)
    np.testing.assert_allclose(got[1], (1 + 5 + 10) / 3.0, rtol=1e-6)
    # fewer than four points: the missing entries stay FLT_MAX (simple_knn.cu:152,185): one of them dominates the mean, two overflow
    np.testing.assert_allclose(K.dist2_mean3(pts[:3]), np.float32(3.4028235e38) / 3, rtol=1e-6)
    assert np.isinf(K.dist2_mean3(pts[:2])).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cloud", "surface", "flat", "duplicates", "tiny"])
def test_dist_matches_oracle(kind):
    from simple_knn._C import distCUDA2
    from oracle import knn_oracle as K
    from gaustar_b200 import scene
    rng = np.random.default_rng(3)
    if kind == "cloud":
        pts = rng.normal(0, 1, (20000, 3)).astype(np.float32)
        pts[:50] *= 40.0  # far outliers stretch the grid
    elif kind == "surface":
        pts = scene.surface_gaussians(12000, 0, seed=2).means3D
    elif kind == "flat":
        pts = rng.uniform(-1, 1, (6000, 3)).astype(np.float32)
        pts[:, 1] = 0.25  # zero extent along y
    elif kind == "duplicates":
        pts = np.repeat(rng.uniform(0, 1, (700, 3)).astype(np.float32), 5, axis=0)
    else:
        pts = rng.uniform(0, 1, (4, 3)).astype(np.float32)
    got = distCUDA2(torch.from_numpy(pts).cuda()).cpu().numpy()
    ref = K.dist2_mean3(pts)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-12)
    if kind == "tiny":
        for m in (3, 2, 1):
            np.testing.assert_allclose(distCUDA2(torch.from_numpy(pts[:m]).cuda()).cpu().numpy(), K.dist2_mean3(pts[:m]), rtol=2e-6)
        assert distCUDA2(torch.zeros(0, 3, device="cuda")).numel() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,P", [("cloud", 200_000), ("surface", 1_000_000), ("surface", 50_000), ("flat", 30_000), ("duplicates", 20_000)])
def test_dist_matches_the_live_reference(kind, P):
    """The shim against the UNMODIFIED reference simple_knn compiled into oracle/_ref (SimpleKNN::knn, simple_knn.cu:188-220) on the
    same points, at sizes the brute-force oracle cannot reach -- this is what pins the shim's parity.  Both search exactly; the only
    freedom is the rounding of dx*dx + dy*dy + dz*dz (FMA contraction) and of the sum of the three, hence rtol 2e-6, not bit equality."""
    from oracle import refgpu
    if not refgpu.knn_available():
        pytest.skip("oracle/_ref/libref_knn.so not built (reference sources absent at build time)")
    from simple_knn._C import distCUDA2
    from gaustar_b200 import scene
    rng = np.random.default_rng(11)
    if kind == "cloud":
        pts = rng.normal(0, 1, (P, 3)).astype(np.float32)
        pts[:100] *= 25.0
    elif kind == "surface":
        pts = scene.surface_gaussians(P, 0, seed=4).means3D
    elif kind == "flat":
        pts = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
        pts[:, 2] = -0.5
    else:
        pts = np.repeat(rng.uniform(0, 1, (P // 4, 3)).astype(np.float32), 4, axis=0)
    t = torch.from_numpy(np.ascontiguousarray(pts)).cuda()
    got, ref = distCUDA2(t), refgpu.knn_mean_dist2(t)
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(ref).all()
    np.testing.assert_allclose(got.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6, atol=1e-12)
