"""TEST INFRASTRUCTURE ONLY -- CPU restatement of simple_knn's distCUDA2 (the product never imports this).

Follows gaussian_splatting/submodules/simple-knn/simple_knn.cu:139-186 (boxMeanDist): for every point, the three smallest
squared distances to OTHER points (the point's own index is skipped, simple_knn.cu:171-172; coincident points count with
distance 0; with fewer than 4 points the missing entries stay FLT_MAX and dominate -- or, two of them, overflow -- the sum), averaged
(simple_knn.cu:185).  Brute force in fp32, chunked; the Morton boxes of the reference only prune the same search.
"""
import numpy as np

FLT_MAX = np.float32(3.4028234663852886e38)


def dist2_mean3(points: np.ndarray, chunk: int = 512) -> np.ndarray:
    pts = np.ascontiguousarray(points, np.float32)
    P = pts.shape[0]
    out = np.zeros(P, np.float32)
    for s in range(0, P, chunk):
        q = pts[s:s + chunk]
        d = q[:, None, :] - pts[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(np.float32)
        d2[np.arange(q.shape[0]), np.arange(s, s + q.shape[0])] = np.inf  # skip self
        k = min(3, P - 1)
        best = np.full((q.shape[0], 3), FLT_MAX, np.float32)
        if k > 0:
            best[:, :k] = np.sort(np.partition(d2, k - 1, axis=1)[:, :k], axis=1)
        with np.errstate(over="ignore"):
            out[s:s + chunk] = (best[:, 0] + best[:, 1] + best[:, 2]) / np.float32(3.0)
    return out
