"""shims/pytorch3d/transforms (parity unpinned: pytorch3d is not in this image) against identities and
against the numpy restatement the scene generator already uses (gaustar_b200/scene.py, pinned by tests/test_scene_cpu.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shims"))
from pytorch3d import transforms as T  # noqa: E402

from gaustar_b200 import scene  # noqa: E402


def _random_rotations(n, seed):
    q = torch.randn(n, 4, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    return q / q.norm(dim=-1, keepdim=True)


def test_quaternion_matrix_round_trip_and_sign():
    q = _random_rotations(500, 0)
    R = T.quaternion_to_matrix(q)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand_as(R), atol=1e-12)
    assert torch.allclose(torch.linalg.det(R), torch.ones(500, dtype=torch.float64), atol=1e-12)
    q2 = T.matrix_to_quaternion(R)
    assert (q2[:, 0] >= 0).all()
    assert torch.allclose(q2, T.standardize_quaternion(q), atol=1e-10)
    # non-unit quaternions give the same rotation (sugar_model.py normalises lazily)
    assert torch.allclose(T.quaternion_to_matrix(3.7 * q), R, atol=1e-12)
    # batched leading dimensions
    assert T.matrix_to_quaternion(R.reshape(5, 100, 3, 3)).shape == (5, 100, 4)


def test_matrix_to_quaternion_matches_the_numpy_restatement_also_near_180_degrees():
    q = _random_rotations(300, 1)
    q[:100, 0] = 1e-9  # rotations by ~180 degrees: the w-pivot candidate is ill-conditioned, another pivot must be taken
    q = q / q.norm(dim=-1, keepdim=True)
    R = T.quaternion_to_matrix(q)
    got = T.matrix_to_quaternion(R).numpy()
    want = scene._matrix_to_quaternion(R.numpy())
    # w ~ 0: the standardized sign is arbitrary there -- compare as rotations
    same = np.minimum(np.abs(got - want).max(-1), np.abs(got + want).max(-1))
    assert same.max() < 1e-9
    assert np.allclose(T.quaternion_to_matrix(torch.from_numpy(got)).numpy(), R.numpy(), atol=1e-9)


def test_apply_invert_multiply():
    q, p = _random_rotations(200, 2), torch.randn(200, 3, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    R = T.quaternion_to_matrix(q)
    assert torch.allclose(T.quaternion_apply(q, p), (R @ p[..., None])[..., 0], atol=1e-12)
    assert torch.allclose(T.quaternion_apply(T.quaternion_invert(q), T.quaternion_apply(q, p)), p, atol=1e-12)
    q2 = _random_rotations(200, 4)
    assert torch.allclose(T.quaternion_to_matrix(T.quaternion_raw_multiply(q, q2)), R @ T.quaternion_to_matrix(q2), atol=1e-12)
    # broadcasting: one quaternion, many points
    assert torch.allclose(T.quaternion_apply(q[:1], p), p @ R[0].T, atol=1e-12)


def test_differentiable():
    q = _random_rotations(50, 5).requires_grad_(True)
    R = T.quaternion_to_matrix(q)
    (T.matrix_to_quaternion(R) * torch.arange(4.0, dtype=torch.float64)).sum().backward()
    assert torch.isfinite(q.grad).all() and q.grad.abs().max() > 0
