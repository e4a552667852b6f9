"""Stand-in for the two `pytorch3d.ops` functions GauSTAR imports (`sugar_model.py:7`, `refine.py`): `knn_points`
and `estimate_pointcloud_normals`.  Restated from the published pytorch3d 0.7.4 behaviour; parity unpinned.

`knn_points(p1, p2, K)` returns squared Euclidean distances and indices of the K nearest `p2` points of every `p1`
point, sorted by distance (pytorch3d/ops/knn.py: `_KNN(dists, idx, knn)`); computed here by chunked brute force on
whatever device the inputs live on -- an initialisation / regularisation helper in GauSTAR, not part of the hot path.
"""
from collections import namedtuple

import torch

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, norm: int = 2, K: int = 1, version: int = -1, return_nn: bool = False,
               return_sorted: bool = True):
    if p1.dim() != 3 or p2.dim() != 3 or p1.shape[0] != p2.shape[0] or p1.shape[2] != p2.shape[2]:
        raise ValueError("pts1 and pts2 must have the same batch dimension and point dimension")
    if norm not in (1, 2):
        raise ValueError("Support for 1 or 2 norm.")
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("shim: heterogeneous batches are not needed by GauSTAR")
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    k = min(K, P2)
    dists = p1.new_zeros(N, P1, K)
    idx = torch.zeros(N, P1, K, dtype=torch.int64, device=p1.device)
    chunk = max(1, min(P1, (1 << 24) // max(P2, 1)))
    for n in range(N):
        for s in range(0, P1, chunk):
            a = p1[n, s:s + chunk]
            if norm == 2:
                d = (a * a).sum(-1, keepdim=True) - 2.0 * a @ p2[n].T + (p2[n] * p2[n]).sum(-1)[None]
                d = d.clamp_min(0.0)
            else:
                d = (a[:, None, :] - p2[n][None]).abs().sum(-1)
            dk, ik = torch.topk(d, k, dim=-1, largest=False, sorted=return_sorted)
            if norm == 2:  # exact distances of the selected neighbours (the expansion above loses digits)
                dk = ((a[:, None, :] - p2[n][ik]) ** 2).sum(-1)
            dists[n, s:s + chunk, :k] = dk
            idx[n, s:s + chunk, :k] = ik
    nn = None
    if return_nn:
        nn = torch.stack([p2[n][idx[n]] for n in range(N)])
    return _KNN(dists=dists, idx=idx, knn=nn)


def estimate_pointcloud_normals(pointclouds, neighborhood_size: int = 50, disambiguate_directions: bool = True, *, use_symeig_workaround: bool = True):
    """Normals as the eigenvector of the smallest eigenvalue of the local covariance (pytorch3d/ops/points_normals.py)."""
    pts = pointclouds
    if pts.dim() != 3:
        raise ValueError("shim: expects a (N, P, 3) tensor")
    k = min(neighborhood_size, pts.shape[1])
    nb = knn_points(pts, pts, K=k, return_nn=True).knn  # N, P, k, 3
    c = nb - nb.mean(dim=2, keepdim=True)
    cov = c.transpose(-1, -2) @ c / k
    _, vecs = torch.linalg.eigh(cov)
    normals = vecs[..., 0]
    if disambiguate_directions:
        # pytorch3d orients by the neighbourhood's centre of mass; any consistent rule serves GauSTAR's use (|n . d|)
        sign = ((nb.mean(dim=2) - pts) * normals).sum(-1, keepdim=True)
        normals = torch.where(sign > 0, -normals, normals)
    return normals
