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
