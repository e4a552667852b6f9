"""`pytorch3d.renderer.cameras` subset: `FoVPerspectiveCameras`, `_get_sfm_calibration_matrix` (see the package docstring).

Conventions restated from pytorch3d 0.7.4 (renderer/cameras.py): row vectors, X_cam = X_world @ R + T, so the camera centre is
C = -T @ R^-1 (`get_camera_center`: the translation row of the inverse world-to-view transform).
"""
import torch


def _get_sfm_calibration_matrix(N, device, focal_length, principal_point, orthographic: bool = False):
    """fx, fy, px, py -> the 4x4 calibration matrix K (cameras.py, same name):
        perspective:  [[fx, 0, px, 0], [0, fy, py, 0], [0, 0, 0, 1], [0, 0, 1, 0]]
        orthographic: [[fx, 0, 0, px], [0, fy, 0, py], [0, 0, 1, 0], [0, 0, 0, 1]]"""
    if not torch.is_tensor(focal_length):
        focal_length = torch.tensor(focal_length, device=device)
    if focal_length.ndim in (0, 1) or focal_length.shape[1] == 1:
        fx = fy = focal_length
    else:
        fx, fy = focal_length.unbind(1)
    if not torch.is_tensor(principal_point):
        principal_point = torch.tensor(principal_point, device=device)
    px, py = principal_point.unbind(1)
    K = fx.new_zeros(N, 4, 4)
    K[:, 0, 0] = fx
    K[:, 1, 1] = fy
    if orthographic:
        K[:, 0, 3] = px
        K[:, 1, 3] = py
        K[:, 2, 2] = 1.0
        K[:, 3, 3] = 1.0
    else:
        K[:, 0, 2] = px
        K[:, 1, 2] = py
        K[:, 3, 2] = 1.0
        K[:, 2, 3] = 1.0
    return K


def _batched(x, n, device, dtype=torch.float32):
    t = x if torch.is_tensor(x) else torch.tensor(x, dtype=dtype)
    t = t.to(device=device, dtype=dtype)
    if t.dim() == 0:
        t = t[None]
    if t.shape[0] == 1 and n > 1:
        t = t.expand(n, *t.shape[1:]).clone()
    return t


class _RowVectorTransform:
    """The small part of pytorch3d.transforms.Transform3d GauSTAR touches: a batch of 4x4 row-vector matrices."""

    def __init__(self, matrix):
        self._matrix = matrix

    def get_matrix(self):
        return self._matrix

    def inverse(self):
        return _RowVectorTransform(torch.linalg.inv(self._matrix))

    def transform_points(self, points):
        p = points if points.dim() == 3 else points[None]
        out = p @ self._matrix[:, :3, :3] + self._matrix[:, 3:4, :3]
        return out if points.dim() == 3 else out[0]


class FoVPerspectiveCameras:
    def __init__(self, znear=1.0, zfar=100.0, aspect_ratio=1.0, fov=60.0, degrees: bool = True, R=None, T=None, K=None, device="cpu"):
        R = torch.eye(3)[None] if R is None else R
        T = torch.zeros(1, 3) if T is None else T
        n = max(R.shape[0], T.shape[0], 1 if K is None else K.shape[0])
        self.device = torch.device(device)
        self.R = _batched(R, n, self.device)
        self.T = _batched(T, n, self.device)
        self.K = None if K is None else _batched(K, n, self.device)
        self.znear = _batched(znear, n, self.device)
        self.zfar = _batched(zfar, n, self.device)
        self.aspect_ratio = _batched(aspect_ratio, n, self.device)
        self.fov = _batched(fov, n, self.device)
        self.degrees = degrees

    def __len__(self):
        return self.R.shape[0]

    def __getitem__(self, index):
        if isinstance(index, int):
            index = [index]
        if isinstance(index, slice):
            index = list(range(len(self)))[index]
        idx = torch.as_tensor(index, dtype=torch.int64, device=self.device)
        out = object.__new__(FoVPerspectiveCameras)
        out.device, out.degrees = self.device, self.degrees
        for k in ("R", "T", "K", "znear", "zfar", "aspect_ratio", "fov"):
            v = getattr(self, k)
            setattr(out, k, None if v is None else v[idx])
        return out

    def to(self, device):
        out = object.__new__(FoVPerspectiveCameras)
        out.device, out.degrees = torch.device(device), self.degrees
        for k in ("R", "T", "K", "znear", "zfar", "aspect_ratio", "fov"):
            v = getattr(self, k)
            setattr(out, k, None if v is None else v.to(device))
        return out

    def cuda(self):
        return self.to("cuda")

    def get_world_to_view_transform(self):
        """X_view = X_world @ R + T (refine.py:604-605 takes the z column of it as the per-Gaussian depth)."""
        n = len(self)
        M = torch.zeros(n, 4, 4, device=self.device, dtype=self.R.dtype)
        M[:, :3, :3] = self.R
        M[:, 3, :3] = self.T
        M[:, 3, 3] = 1.0
        return _RowVectorTransform(M)

    def get_camera_center(self):
        return self.get_world_to_view_transform().inverse().get_matrix()[:, 3, :3]

    def get_projection_transform(self):
        raise NotImplementedError("shims/pytorch3d: the FoV projection transform is not needed by GauSTAR's calls (K is always given)")
