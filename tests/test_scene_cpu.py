"""BASELINE.json config #1 on the CPU (no rasterizer, no GPU): 5k-face synthetic mesh -> SuGaR-bound Gaussians -> SH
evaluation for one 128x128 view.  Checks the scene generator's restatement of the SuGaR parameterisation
(gaustar_scene/sugar_model.py:180-226,354-368,417-508) and pins the oracle's SH->RGB stage against
colours computed by the REFERENCE's own eval_sh (tests/golden/make_config1_golden.py)."""
import os

import numpy as np

from gaustar_b200 import scene
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capsule_mesh_is_a_sphere_with_pinholes():
    """Exactly n faces: the generator triangulates a closed genus-0 surface and drops the few surplus faces at random
    (scene.capsule_mesh), so the result is a manifold sphere with isolated one-triangle holes."""
    verts, faces = scene.capsule_mesh(5000, seed=0)
    assert faces.shape == (5000, 3) and verts.dtype == np.float32
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    uniq, cnt = np.unique(e, axis=0, return_counts=True)
    assert cnt.max() == 2                                        # manifold: no edge with more than two faces
    boundary = int((cnt == 1).sum())                            # edges of the holes the dropped faces left
    assert boundary <= 3 * 0.04 * len(faces)
    # a sphere with k boundary loops has V - E + F = 2 - k, and every loop has at least three edges: no handles
    chi = len(verts) - len(uniq) + len(faces)
    assert 2 - boundary / 3 <= chi <= 2
    assert len(np.unique(faces)) == len(verts)                   # no unused vertices
    c = verts.mean(0)
    assert abs(c[0]) < 0.02 and abs(c[1] - 1.0) < 0.05 and abs(c[2]) < 0.02


def test_sugar_binding_of_config1():
    verts, faces = scene.capsule_mesh(5000, seed=0)
    g = scene.bind_gaussians(verts, faces, sh_degree=3, seed=0)
    F = len(faces)
    assert g.P == 6 * F == 30000 and g.shs.shape == (g.P, 16, 3)
    fv = verts[faces].astype(np.float64)
    n = np.cross(fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0])
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    pts = g.means3D.reshape(F, 6, 3).astype(np.float64)
    # six Gaussians per triangle, in its plane, strictly inside it (barycentric pattern of sugar_model.py:217-226)
    assert np.abs(np.einsum("fgc,fc->fg", pts - fv[:, None, 0], n)).max() < 1e-5
    T = np.stack([fv[:, 1] - fv[:, 0], fv[:, 2] - fv[:, 0]], -1)                 # [F,3,2]
    uv = np.einsum("fij,fgj->fgi", np.linalg.pinv(T), pts - fv[:, None, 0])      # barycentric (b1, b2)
    b = np.concatenate([1 - uv.sum(-1, keepdims=True), uv], -1)
    assert b.min() > 0.05 and np.allclose(b.sum(-1), 1.0)
    assert np.allclose(b.mean(1), 1.0 / 3.0, atol=1e-6)                          # the pattern is centred on the centroid
    # flat discs: thickness = extent * 1e-6 along the face normal, equal in-plane radii = min_edge / (4 + 2 sqrt 3)
    min_edge = np.linalg.norm(fv - fv[:, [1, 2, 0]], axis=-1).min(-1)
    sc = g.scales.reshape(F, 6, 3)
    assert np.allclose(sc[..., 0], 3.0e-6, rtol=1e-5)
    assert np.allclose(sc[..., 1], (min_edge / (4 + 2 * np.sqrt(3)))[:, None], rtol=1e-4) and np.array_equal(sc[..., 1], sc[..., 2])
    # unit quaternions (w,x,y,z) whose first rotation axis -- the thin one -- is the face normal
    q = g.rotations.astype(np.float64)
    assert np.allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-6)
    w, x, y, z = q.T
    axis0 = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y + w * z), 2 * (x * z - w * y)], -1)  # first column of R(q)
    assert np.abs(np.einsum("pc,pc->p", axis0, np.repeat(n, 6, 0))).min() > 1 - 1e-5
    assert 0.8 <= g.opacities.min() and g.opacities.max() <= 0.99


def test_oracle_sh_stage_matches_reference_eval_sh():
    z = np.load(os.path.join(ROOT, "tests", "golden", "config1", "sh_rgb.npz"))
    g = scene.surface_gaussians(30000, sh_degree=3, seed=0)
    cam = scene.dome_cameras(4, 128, 128)[1]
    assert g.P == int(z["P"]) and np.array_equal(cam.campos.astype(np.float32), z["campos"])
    inp = O.Inputs(means3D=g.means3D, opacities=g.opacities, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos,
                   bg=np.zeros(3, np.float32), tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, W=128, H=128, shs=g.shs, scales=g.scales,
                   rotations=g.rotations, sh_degree=3)
    f = O.forward(inp, blend=False)
    idx = z["idx"].astype(np.int64)
    vis = f.radii[idx] > 0
    assert vis.sum() > 0.9 * len(idx)  # the whole body is in front of the camera
    np.testing.assert_allclose(f.rgb[idx][vis], z["rgb"][vis], rtol=2e-5, atol=2e-6)
