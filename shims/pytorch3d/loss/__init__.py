"""`pytorch3d.loss` regularisers GauSTAR imports (gaustar_trainers/refine.py:6, used at :681-686), restated from the
published definitions of pytorch3d 0.7.4 (PARITY UNPINNED: the package is not in this image).

mesh_normal_consistency: for every PAIR of faces sharing an edge (v0, v1), with a and b the faces' third vertices,
    n0 = (v1 - v0) x (a - v0),  n1 = (v1 - v0) x (b - v0),  term = 1 - cos(n0, -n1)
(coplanar neighbours lie on opposite sides of the edge, so n0 = -n1 and the term is 0); the terms of a mesh are averaged,
the meshes of the batch are averaged.
mesh_laplacian_smoothing(method="uniform"): L = D^-1 A - I on the vertex graph, loss = mean over a mesh's vertices of
||(L V)_i||, averaged over the batch.
"""
import torch


def mesh_normal_consistency(meshes):
    if meshes.isempty():
        return torch.tensor([0.0], dtype=torch.float32, device=meshes.device, requires_grad=True)
    N = len(meshes)
    verts, faces = meshes.verts_packed(), meshes.faces_packed()
    f2e = meshes.faces_packed_to_edges_packed()          # [F,3]: edge opposite vertex k
    E = meshes.edges_packed().shape[0]
    F = faces.shape[0]
    # (edge, face, third vertex) incidences, grouped by edge
    edge_of = f2e.reshape(-1)
    face_of = torch.arange(F, device=faces.device).repeat_interleave(3)
    third = faces.reshape(-1)                             # vertex k is the one opposite edge k
    order = torch.argsort(edge_of, stable=True)
    edge_of, face_of, third = edge_of[order], face_of[order], third[order]
    counts = torch.bincount(edge_of, minlength=E)
    start = torch.cumsum(counts, 0) - counts
    pairs_a, pairs_b = [], []
    for k in range(int(counts.max()) if E else 0):       # all C(k,2) pairs of an edge's faces (2 faces per edge on a manifold)
        for l in range(k + 1, int(counts.max())):
            has = counts > l
            pairs_a.append(start[has] + k)
            pairs_b.append(start[has] + l)
    if not pairs_a:  # no edge with two faces: the sum over no terms, a scalar (0.7.4: `loss.sum() / N`)
        return (verts * 0.0).sum() / N
    ia, ib = torch.cat(pairs_a), torch.cat(pairs_b)
    edges = meshes.edges_packed()[edge_of[ia]]
    v0, v1 = verts[edges[:, 0]], verts[edges[:, 1]]
    a, b = verts[third[ia]], verts[third[ib]]
    n0 = torch.cross(v1 - v0, a - v0, dim=1)
    n1 = torch.cross(v1 - v0, b - v0, dim=1)
    loss = 1.0 - torch.cosine_similarity(n0, -n1, dim=1)
    mesh_idx = meshes.verts_packed_to_mesh_idx()[edges[:, 0]]
    weights = 1.0 / torch.bincount(mesh_idx, minlength=N)[mesh_idx].to(loss.dtype)
    return (loss * weights).sum() / N


def mesh_laplacian_smoothing(meshes, method: str = "uniform"):
    if method != "uniform":
        raise NotImplementedError("shim: only method='uniform' (the one refine.py:121 selects) is provided")
    if meshes.isempty():
        return torch.tensor([0.0], dtype=torch.float32, device=meshes.device, requires_grad=True)
    N = len(meshes)
    verts, edges = meshes.verts_packed(), meshes.edges_packed()
    V = verts.shape[0]
    e0, e1 = edges[:, 0], edges[:, 1]
    deg = torch.zeros(V, dtype=verts.dtype, device=verts.device).index_add(0, e0, torch.ones_like(e0, dtype=verts.dtype)).index_add(
        0, e1, torch.ones_like(e1, dtype=verts.dtype))
    nbr = torch.zeros_like(verts).index_add(0, e0, verts[e1]).index_add(0, e1, verts[e0])
    lap = nbr / deg.clamp(min=1.0)[:, None] - verts  # L = D^-1 A - I: the -I term also for a vertex without neighbours (its row gives -v_i)
    mesh_idx = meshes.verts_packed_to_mesh_idx()
    weights = 1.0 / meshes.num_verts_per_mesh()[mesh_idx].to(verts.dtype)
    return (lap.norm(dim=1) * weights).sum() / N
