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
    # pytorch3d 0.7.4 (the version GauSTAR pins) returns the best-conditioned candidate WITHOUT standardising its sign
    assert torch.allclose(T.standardize_quaternion(q2), T.standardize_quaternion(q), atol=1e-10)
    assert torch.allclose(T.quaternion_to_matrix(q2), R, atol=1e-10)
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


def _cube():
    v = torch.tensor([[x, y, z] for x in (0.0, 1.0) for y in (0.0, 1.0) for z in (0.0, 1.0)], dtype=torch.float64)
    f = torch.tensor([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4],
                      [1, 5, 7], [1, 7, 3]])
    return v, f


def test_meshes_packing_edges_areas_normals():
    from pytorch3d.structures import Meshes
    v, f = _cube()
    m = Meshes(verts=[v, v + 5.0], faces=[f, f])
    assert len(m) == 2 and m.verts_packed().shape == (16, 3) and m.faces_packed().shape == (24, 3)
    assert m.faces_packed()[12:].min() == 8  # the second mesh's indices are offset
    e = m.edges_packed()
    assert e.shape == (36, 2) and (e[:, 0] < e[:, 1]).all()  # 12 cube edges + 6 face diagonals per mesh, each once
    key = e[:, 0] * 16 + e[:, 1]
    assert (key[1:] > key[:-1]).all()  # sorted lexicographically
    f2e, F = m.faces_packed_to_edges_packed(), m.faces_packed()
    for k, (i, j) in enumerate(((1, 2), (2, 0), (0, 1))):  # column k is the edge opposite vertex k
        pair = torch.stack([torch.minimum(F[:, i], F[:, j]), torch.maximum(F[:, i], F[:, j])], 1)
        assert torch.equal(e[f2e[:, k]], pair)
    assert torch.allclose(m.faces_areas_packed(), torch.full((24,), 0.5, dtype=torch.float64))
    n = m.faces_normals_packed()
    centre = torch.tensor([0.5, 0.5, 0.5], dtype=torch.float64)
    fc = v[f].mean(1)
    assert torch.allclose(n.norm(dim=1), torch.ones(24, dtype=torch.float64)) and ((n[:12] * (fc - centre)).sum(1) > 0).all()  # outward winding
    assert [x.shape for x in m.faces_normals_list()] == [(12, 3), (12, 3)] and [x.shape for x in m.verts_list()] == [(8, 3), (8, 3)]


def test_normal_consistency_closed_forms():
    from pytorch3d.loss import mesh_normal_consistency
    from pytorch3d.structures import Meshes
    # a flat strip: every interior edge joins coplanar faces
    gx, gy = torch.meshgrid(torch.arange(4.0), torch.arange(3.0), indexing="ij")
    pv = torch.stack([gx.reshape(-1), gy.reshape(-1), torch.zeros(12)], 1)
    idx = lambda i, j: i * 3 + j
    pf = torch.tensor([t for i in range(3) for j in range(2) for t in ([idx(i, j), idx(i + 1, j), idx(i + 1, j + 1)], [idx(i, j), idx(i + 1, j + 1), idx(i, j + 1)])])
    assert abs(float(mesh_normal_consistency(Meshes([pv], [pf])))) < 1e-6
    # cube: 12 right-angle edges (1 - cos 90 = 1) and 6 coplanar face diagonals (0) -> 12 / 18
    v, f = _cube()
    assert abs(float(mesh_normal_consistency(Meshes([v], [f]))) - 12.0 / 18.0) < 1e-9
    # regular tetrahedron: outward normals of adjacent faces have cosine -1/3 -> 4/3 on each of the 6 edges
    tv = torch.tensor([[1.0, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=torch.float64)
    tf = torch.tensor([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]])
    assert abs(float(mesh_normal_consistency(Meshes([tv], [tf]))) - 4.0 / 3.0) < 1e-9
    # a batch is the mean of its meshes; the loss is differentiable
    both = mesh_normal_consistency(Meshes([v, tv], [f, tf]))
    assert abs(float(both) - 0.5 * (12.0 / 18.0 + 4.0 / 3.0)) < 1e-9
    vv = v.clone().requires_grad_(True)
    mesh_normal_consistency(Meshes([vv], [f])).backward()
    assert torch.isfinite(vv.grad).all()


def test_uniform_laplacian_closed_forms():
    from pytorch3d.loss import mesh_laplacian_smoothing
    from pytorch3d.structures import Meshes
    # regular tetrahedron centred at the origin: every vertex's neighbours average to -v/3, so ||L v|| = (4/3) |v|
    tv = torch.tensor([[1.0, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], dtype=torch.float64)
    tf = torch.tensor([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]])
    assert abs(float(mesh_laplacian_smoothing(Meshes([tv], [tf]), method="uniform")) - 4.0 / 3.0 * 3.0 ** 0.5) < 1e-9
    # scaling the mesh scales the loss; translating it does not change it
    assert abs(float(mesh_laplacian_smoothing(Meshes([2 * tv + 7.0], [tf]))) - 8.0 / 3.0 * 3.0 ** 0.5) < 1e-9


def test_knn_points_matches_brute_force_and_normals_of_a_plane():
    from pytorch3d.ops import estimate_pointcloud_normals, knn_points
    g = torch.Generator().manual_seed(0)
    p = torch.rand(2, 300, 3, generator=g)
    q = torch.rand(2, 70, 3, generator=g)
    out = knn_points(q, p, K=5, return_nn=True)
    d = ((q[:, :, None] - p[:, None]) ** 2).sum(-1)
    dk, ik = torch.topk(d, 5, dim=-1, largest=False)
    assert torch.allclose(out.dists, dk, atol=1e-6) and torch.equal(out.idx, ik)
    assert torch.allclose(out.knn, torch.stack([p[n][ik[n]] for n in range(2)]))
    plane = torch.cat([torch.rand(1, 500, 2, generator=g), torch.zeros(1, 500, 1)], -1)
    n = estimate_pointcloud_normals(plane, neighborhood_size=12)
    assert torch.allclose(n.abs(), torch.tensor([0.0, 0.0, 1.0]).expand_as(n), atol=1e-4)


def test_camera_container_and_calibration_matrix():
    """What GauSTAR does with pytorch3d cameras (cameras.py:229-330,537-548; sugar_model.py:1113-1162): build from R/T/K, index,
    ask for the centres.  Row-vector convention: X_cam = X_world R + T  =>  C R + T = 0."""
    from pytorch3d.renderer import FoVPerspectiveCameras, RasterizationSettings, TexturesVertex
    from pytorch3d.renderer.cameras import _get_sfm_calibration_matrix
    from pytorch3d.transforms import quaternion_to_matrix
    g = torch.Generator().manual_seed(1)
    R = quaternion_to_matrix(torch.nn.functional.normalize(torch.randn(4, 4, generator=g), dim=-1))
    T = torch.randn(4, 3, generator=g)
    K = _get_sfm_calibration_matrix(4, "cpu", torch.tensor([[2.0, 3.0]]).expand(4, -1), torch.tensor([[0.1, -0.2]]).expand(4, -1))
    assert torch.equal(K[0], torch.tensor([[2.0, 0, 0.1, 0], [0, 3.0, -0.2, 0], [0, 0, 0, 1], [0, 0, 1, 0]]))
    cams = FoVPerspectiveCameras(R=R, T=T, K=K, znear=0.0001)
    C = cams.get_camera_center()
    assert torch.allclose(torch.einsum("ni,nij->nj", C, R) + T, torch.zeros(4, 3), atol=1e-5)
    one = cams[2]
    assert len(cams) == 4 and len(one) == 1 and torch.equal(one.R[0], R[2]) and torch.equal(one.K[0], K[2])
    assert abs(one.znear.item() - 1e-4) < 1e-9 and one.zfar.item() == 100.0
    assert torch.allclose(one.get_camera_center()[0], C[2])
    assert RasterizationSettings(image_size=(4, 5)).image_size == (4, 5)
    tv = TexturesVertex(verts_features=torch.rand(1, 7, 3))
    assert tv.verts_features_packed().shape == (7, 3)


def test_regularisers_on_degenerate_meshes_follow_0_7_4():
    """An isolated vertex keeps the -I row of L = D^-1 A - I (its term is ||v_i||); a mesh without an edge shared by two faces gives a
    SCALAR zero normal-consistency loss (`loss.sum() / N`), not a one-element tensor."""
    from pytorch3d.loss import mesh_laplacian_smoothing, mesh_normal_consistency
    from pytorch3d.structures import Meshes
    v = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [3.0, 4.0, 0.0]], dtype=torch.float64, requires_grad=True)  # vertex 3 is in no face
    f = torch.tensor([[0, 1, 2]])
    m = Meshes([v], [f])
    nc = mesh_normal_consistency(m)
    assert nc.dim() == 0 and float(nc) == 0.0
    nc.backward()  # differentiable (zero gradient)
    lap = mesh_laplacian_smoothing(m)
    # the triangle's vertices: ||mean of the two neighbours - v||; the isolated one: ||v_3|| = 5
    tri = [((v[1] + v[2]) / 2 - v[0]).norm(), ((v[0] + v[2]) / 2 - v[1]).norm(), ((v[0] + v[1]) / 2 - v[2]).norm()]
    want = (sum(float(t) for t in tri) + 5.0) / 4.0
    assert abs(float(lap) - want) < 1e-12
