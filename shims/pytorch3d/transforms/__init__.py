"""`pytorch3d.transforms` functions GauSTAR imports (gaustar_scene/sugar_model.py:6, sugar_densifier.py:4,
sugar_compositor.py:3), restated from the published algorithms of pytorch3d 0.7.4 (pinned by the reference's
environment.yml:161; the package itself is not in this image -- PARITY UNPINNED, tests check identities).

Quaternions are (w, x, y, z), real part first; all functions broadcast over leading dimensions and are differentiable.
`matrix_to_quaternion` returns the standardized sign (w >= 0): q and -q are the same rotation, so the rasterizer's
results do not depend on it.
"""
import torch
import torch.nn.functional as F


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    """sqrt(max(0, x)) with a zero subgradient where x is 0."""
    ret = torch.zeros_like(x)
    positive = x > 0
    ret[positive] = torch.sqrt(x[positive])
    return ret


def standardize_quaternion(quaternions: torch.Tensor) -> torch.Tensor:
    return torch.where(quaternions[..., 0:1] < 0, -quaternions, quaternions)


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    """Rotation matrices [..., 3, 3] -> quaternions [..., 4].  Four candidates (one per component taken as the pivot) are
    formed from the matrix; the one with the largest pivot is the best conditioned and is returned."""
    if matrix.size(-1) != 3 or matrix.size(-2) != 3:
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(matrix.reshape(batch_dim + (9,)), dim=-1)
    q_abs = _sqrt_positive_part(torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], dim=-1))
    quat_by_rijk = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1)], dim=-2)
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(floor))
    out = candidates[F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5, :].reshape(batch_dim + (4,))
    return out  # (0.7.4, the version GauSTAR pins, does not standardise the sign here; later releases do)


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """Quaternions [..., 4] (not necessarily unit) -> rotation matrices [..., 3, 3]."""
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))


def quaternion_raw_multiply(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product a * b."""
    aw, ax, ay, az = torch.unbind(a, -1)
    bw, bx, by, bz = torch.unbind(b, -1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)


def quaternion_invert(quaternion: torch.Tensor) -> torch.Tensor:
    """Inverse of a UNIT quaternion (the conjugate)."""
    return quaternion * quaternion.new_tensor([1, -1, -1, -1])


def quaternion_apply(quaternion: torch.Tensor, point: torch.Tensor) -> torch.Tensor:
    """Rotate points [..., 3] by unit quaternions [..., 4]: q (0, p) q^-1."""
    if point.size(-1) != 3:
        raise ValueError(f"Points are not in 3D, {point.shape}.")
    as_quat = torch.cat((point.new_zeros(point.shape[:-1] + (1,)), point), -1)
    return quaternion_raw_multiply(quaternion_raw_multiply(quaternion, as_quat), quaternion_invert(quaternion))[..., 1:]
