"""Synthetic inputs for the rasterizer: mesh-bound surface Gaussians + camera dome.

Host-side (numpy) restatement of how GauSTAR's caller builds the operator's
inputs, used by the tests and bench.py:

* 6 Gaussians per mesh triangle at the barycentrics of
  ``gaustar_scene/sugar_model.py:214-226``, in-plane scale
  ``min_edge / (4 + 2*sqrt(3))`` (``:216,357-358``), thickness
  ``spatial_extent * 1e-6`` (``:180``), frame = [face normal | first edge | cross]
  (``:484-508``) converted to a (w,x,y,z) quaternion.
* camera matrices exactly as ``sugar_model.py:1130-1163`` builds them:
  ``viewmatrix = getWorld2View(R,t).T`` and ``projmatrix = viewmatrix @ P.T``
  (``gaustar_utils/graphics_utils.py:38-85``), row-vector convention.

Nothing here touches the GPU; arrays are float32 numpy.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

SH_C0 = 0.28209479177387814

_BARY6 = np.array(
    [[2 / 3, 1 / 6, 1 / 6], [1 / 6, 2 / 3, 1 / 6], [1 / 6, 1 / 6, 2 / 3],
     [1 / 6, 5 / 12, 5 / 12], [5 / 12, 1 / 6, 5 / 12], [5 / 12, 5 / 12, 1 / 6]], dtype=np.float32)
_CIRCLE_RADIUS6 = 1.0 / (4.0 + 2.0 * math.sqrt(3.0))


def _uv_sphere(n_faces):
    """Plain UV sphere (constant longitude count): clusters tiny triangles at the poles. Stress-test only."""
    n_lat = max(3, int(math.ceil(math.sqrt(n_faces / 4.0))) + 1)
    n_lon = max(3, int(math.ceil(n_faces / (2.0 * (n_lat - 1)))))
    while 2 * n_lon * (n_lat - 1) < n_faces:
        n_lon += 1
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    rings = [np.stack([np.sin(t) * np.cos(ph), np.full_like(ph, np.cos(t)), np.sin(t) * np.sin(ph)], -1)
             for t in th for ph in [np.linspace(0, 2 * np.pi, n_lon, endpoint=False)]]
    return rings


def _band_sphere(n_faces):
    """Latitude rings whose vertex count is proportional to sin(latitude): near-uniform triangles everywhere."""
    L = max(4, int(math.ceil(math.sqrt(n_faces / 1.7))))
    while True:
        th = np.linspace(0, np.pi, L + 1)[1:-1]
        counts = np.maximum(5, np.round(1.3333 * L * np.sin(th)).astype(int))
        faces = counts[0] + counts[-1] + int((counts[:-1] + counts[1:]).sum())
        if faces >= n_faces:
            break
        L += 1
    rings = []
    for i, (t, c) in enumerate(zip(th, counts)):
        ph = np.linspace(0, 2 * np.pi, c, endpoint=False) + (0.5 * i) * 2 * np.pi / c
        rings.append(np.stack([np.sin(t) * np.cos(ph), np.full_like(ph, np.cos(t)), np.sin(t) * np.sin(ph)], -1))
    return rings


def _zip_band(off_a, ang_a, off_b, ang_b):
    """Triangulate the band between two rings with different vertex counts by walking both rings in angle order.
    ang_* are the ascending longitudes of the rings' vertices in [0, 2pi)."""
    na, nb = len(ang_a), len(ang_b)
    # start ring b at its first vertex at or after ring a's first vertex
    sb = int(np.searchsorted(ang_b, ang_a[0])) % nb
    ib = (np.arange(nb) + sb) % nb
    pa = np.concatenate([ang_a - ang_a[0], [2 * np.pi]])
    pb = np.mod(ang_b[ib] - ang_a[0], 2 * np.pi)
    pb = np.concatenate([pb, [pb[0] + 2 * np.pi]])
    faces = []
    i = j = 0
    while i < na or j < nb:
        if j >= nb or (i < na and pa[i + 1] <= pb[j + 1]):
            faces.append([off_a + i % na, off_a + (i + 1) % na, off_b + ib[j % nb]])
            i += 1
        else:
            faces.append([off_a + i % na, off_b + ib[(j + 1) % nb], off_b + ib[j % nb]])
            j += 1
    return faces


def capsule_mesh(n_faces: int, seed: int = 0, radius: float = 0.35, stretch_y: float = 1.8, center=(0.0, 1.0, 0.0), uniform: bool = True):
    """Closed genus-0 'capsule' (sphere stretched in y) with exactly n_faces triangles.

    uniform=True (default): latitude bands with vertex count ~ sin(latitude), i.e. near-uniform triangle size like
    a reconstructed body mesh (TSDF + decimation).  uniform=False: plain UV sphere, whose pole singularity piles
    thousands of tiny triangles into a few pixels -- kept as a stress test for very long tile lists."""
    rings = _band_sphere(n_faces) if uniform else _uv_sphere(n_faces)
    verts = [np.array([[0.0, 1.0, 0.0]])] + rings + [np.array([[0.0, -1.0, 0.0]])]
    offs = np.cumsum([0] + [len(v) for v in verts])
    lon = [np.mod(np.arctan2(r[:, 2], r[:, 0]), 2 * np.pi) for r in rings]
    faces = []
    n0 = len(rings[0])
    faces += [[0, offs[1] + (j + 1) % n0, offs[1] + j] for j in range(n0)]
    for k in range(len(rings) - 1):
        oa, ob = np.argsort(lon[k]), np.argsort(lon[k + 1])
        fz = _zip_band(0, lon[k][oa], 1 << 40, lon[k + 1][ob])
        for tri in fz:
            faces.append([offs[k + 2] + ob[v - (1 << 40)] if v >= (1 << 40) else offs[k + 1] + oa[v] for v in tri])
    last = offs[-1] - 1
    nl = len(rings[-1])
    faces += [[last, offs[-3] + j, offs[-3] + (j + 1) % nl] for j in range(nl)]
    verts = np.concatenate(verts, 0)
    faces = np.asarray(faces, dtype=np.int64)
    # consistent outward orientation
    fv = verts[faces]
    nrm = np.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0])
    flip = (nrm * fv.mean(1)).sum(-1) < 0
    faces[flip] = faces[flip][:, [0, 2, 1]]
    rng = np.random.default_rng(seed)
    # drop random surplus faces so the face count is exact (tiny holes; irrelevant to the op)
    if len(faces) > n_faces:
        keep = np.sort(rng.permutation(len(faces))[:n_faces])
        faces = faces[keep]
    verts = verts * np.array([radius, radius * stretch_y, radius])
    fv = verts[faces]
    mean_edge = np.linalg.norm(fv - fv[:, [1, 2, 0]], axis=-1).mean()
    verts = verts + rng.normal(0, 0.1 * mean_edge, verts.shape)
    verts = verts + np.asarray(center)
    return verts.astype(np.float32), faces


def _matrix_to_quaternion(R: np.ndarray) -> np.ndarray:
    """Rotation matrices [N,3,3] -> unit quaternions (w,x,y,z), w >= 0."""
    m00, m01, m02 = R[:, 0, 0], R[:, 0, 1], R[:, 0, 2]
    m10, m11, m12 = R[:, 1, 0], R[:, 1, 1], R[:, 1, 2]
    m20, m21, m22 = R[:, 2, 0], R[:, 2, 1], R[:, 2, 2]
    q_abs = np.sqrt(np.maximum(0.0, np.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)))
    cand = np.stack([
        np.stack([q_abs[:, 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        np.stack([m21 - m12, q_abs[:, 1] ** 2, m10 + m01, m02 + m20], -1),
        np.stack([m02 - m20, m10 + m01, q_abs[:, 2] ** 2, m12 + m21], -1),
        np.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[:, 3] ** 2], -1)], -2)
    cand = cand / (2.0 * np.maximum(q_abs[..., None], 0.1))
    best = np.argmax(q_abs, -1)
    q = cand[np.arange(len(R)), best]
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    q[q[:, 0] < 0] *= -1
    return q


@dataclass
class Gaussians:
    means3D: np.ndarray    # [P,3]
    scales: np.ndarray     # [P,3] (thickness, s1, s2) -- activated
    rotations: np.ndarray  # [P,4] (w,x,y,z), normalised
    opacities: np.ndarray  # [P,1] in (0,1)
    shs: np.ndarray        # [P,M,3]

    @property
    def P(self):
        return self.means3D.shape[0]


def bind_gaussians(verts, faces, sh_degree=3, seed=0, opacity="trained", extent=3.0) -> Gaussians:
    """Surface Gaussians bound to a triangle mesh, SuGaR-style (sugar_model.py:180-226,354-368,417-508)."""
    rng = np.random.default_rng(seed + 1)
    fv = verts[faces].astype(np.float64)  # [F,3,3]
    F = len(faces)
    pts = np.einsum("gk,fkc->fgc", _BARY6.astype(np.float64), fv).reshape(-1, 3)
    min_edge = np.linalg.norm(fv - fv[:, [1, 2, 0]], axis=-1).min(-1)
    plane = np.maximum(min_edge * _CIRCLE_RADIUS6, 1e-7)
    thickness = extent / 1_000_000.0
    scales = np.stack([np.full(F, thickness), plane, plane], -1)
    scales = np.repeat(scales, 6, 0)
    n = np.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0])
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-20)
    e1 = fv[:, 0] - fv[:, 1]
    e1 /= np.maximum(np.linalg.norm(e1, axis=-1, keepdims=True), 1e-20)
    e2 = np.cross(n, e1)
    e2 /= np.maximum(np.linalg.norm(e2, axis=-1, keepdims=True), 1e-20)
    R = np.stack([n, e1, e2], -1)  # columns
    quat = np.repeat(_matrix_to_quaternion(R), 6, 0)
    P = 6 * F
    if opacity == "trained":
        op = rng.uniform(0.8, 0.99, (P, 1))
    elif opacity == "init":
        op = np.full((P, 1), 0.1)
    else:
        op = np.full((P, 1), float(opacity))
    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0, 1, (P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0, 0.05, (P, M - 1, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return Gaussians(f32(pts), f32(scales), f32(quat), f32(op), f32(shs))


def surface_gaussians(P: int, sh_degree=3, seed=0, opacity="trained", uniform=True) -> Gaussians:
    """ceil(P/6) faces -> 6*ceil(P/6) Gaussians."""
    F = (P + 5) // 6
    verts, faces = capsule_mesh(F, seed, uniform=uniform)
    return bind_gaussians(verts, faces, sh_degree, seed, opacity)


def random_gaussians(P: int, sh_degree=3, seed=0, scale_range=(0.005, 0.3), extent=1.5) -> Gaussians:
    """Generic 3DGS-like cloud (anisotropic, wide range of sizes) for stress / parity tests."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-extent, extent, (P, 3)) + np.array([0, 1.0, 0])
    ls = rng.uniform(np.log(scale_range[0]), np.log(scale_range[1]), (P, 3))
    scales = np.exp(ls)
    q = rng.normal(0, 1, (P, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    op = rng.uniform(0.02, 0.99, (P, 1))
    M = (sh_degree + 1) ** 2
    shs = np.zeros((P, M, 3))
    shs[:, 0, :] = (rng.uniform(0, 1, (P, 3)) - 0.5) / SH_C0
    if M > 1:
        shs[:, 1:, :] = rng.normal(0, 0.15, (P, M - 1, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return Gaussians(f32(means), f32(scales), f32(q), f32(op), f32(shs))


@dataclass
class Camera:
    """Exactly the camera-dependent fields of GaussianRasterizationSettings (DGR/__init__.py:157-169)."""
    image_width: int
    image_height: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray  # [4,4] = getWorld2View(R,t).T
    projmatrix: np.ndarray  # [4,4] = viewmatrix @ P.T
    campos: np.ndarray      # [3]


def _projection(znear, zfar, tan_half_x, tan_half_y):
    """graphics_utils.py:65-85 getProjectionMatrix (symmetric frustum)."""
    P = np.zeros((4, 4), dtype=np.float64)
    P[0, 0] = 1.0 / tan_half_x
    P[1, 1] = 1.0 / tan_half_y
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def look_at_camera(eye, target, W, H, fy_over_H=1.6667, znear=0.01, zfar=100.0, principal_ndc=(0.0, 0.0)) -> Camera:
    """principal_ndc: off-centre principal point the way SuGaR hands it to the op -- sugar_model.py:1160-1161 writes
    proj[2,0] = -K[0,0,2], proj[2,1] = -K[0,1,2] into the transposed projection (zero for GauSTAR's centred data)."""
    eye = np.asarray(eye, np.float64)
    target = np.asarray(target, np.float64)
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    if np.linalg.norm(right) < 1e-6:
        right = np.array([1.0, 0, 0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    # COLMAP axes: x right, y down, z forward. Rows of the w2c rotation.
    Rw2c = np.stack([right, down, fwd], 0)
    t = -Rw2c @ eye
    # getWorld2View(R=Rw2c.T, t): Rt[:3,:3] = R.T = Rw2c
    Rt = np.eye(4)
    Rt[:3, :3] = Rw2c
    Rt[:3, 3] = t
    view = np.float32(Rt).T.astype(np.float64)
    fy = fy_over_H * H
    fx = fy
    tan_x, tan_y = W / (2.0 * fx), H / (2.0 * fy)
    proj = _projection(znear, zfar, tan_x, tan_y).astype(np.float32).T.astype(np.float64)
    proj[2, 0] = -float(principal_ndc[0])
    proj[2, 1] = -float(principal_ndc[1])
    full = (view.astype(np.float32) @ proj.astype(np.float32)).astype(np.float32)
    return Camera(W, H, float(tan_x), float(tan_y), np.ascontiguousarray(view, np.float32), np.ascontiguousarray(full, np.float32),
                  eye.astype(np.float32))


def dome_cameras(V: int, W: int, H: int, radius=3.0, target=(0.0, 1.0, 0.0), seed=0):
    """ActorsHQ-like dome: V views on 5 rings, elevation -30..+40 degrees."""
    rings = np.linspace(-30.0, 40.0, 5)
    cams = []
    per = int(math.ceil(V / len(rings)))
    k = 0
    for ri, elev in enumerate(rings):
        for j in range(per):
            if k >= V:
                break
            az = 2 * math.pi * (j + 0.5 * (ri % 2)) / per
            e = math.radians(elev)
            eye = np.asarray(target) + radius * np.array([math.cos(e) * math.sin(az), math.sin(e), math.cos(e) * math.cos(az)])
            cams.append(look_at_camera(eye, target, W, H))
            k += 1
    return cams
