"""``simple_knn._C.distCUDA2`` -- mean squared distance of every point to its three nearest neighbours
(replaces gaussian_splatting/submodules/simple-knn/spatial.cu:15-26 + simple_knn.cu:188-220; used by
gaussian_splatting/scene/gaussian_model.py:134 to initialise the Gaussian scales).

Host side of a uniform-grid search: bin the points into cells of about two points each (torch sort / searchsorted are
plumbing here -- this runs once per model, not per view), then one CUDA kernel of libgstar_raster.so (csrc/knn.cu)
walks the cell shells around every point.  CUDA tensors only: there is no CPU path.
"""
import ctypes as C

import torch

from gaustar_b200 import capi

_MAX_CELLS = 1 << 22


def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if not points.is_cuda:
        raise RuntimeError("simple_knn (gaustar_b200): distCUDA2 needs a CUDA tensor (there is no CPU path)")
    pts = points.detach().to(torch.float32).contiguous()
    P = pts.shape[0]
    out = torch.zeros(P, dtype=torch.float32, device=pts.device)  # spatial.cu:20-21
    if P == 0:
        return out
    mn, mx = pts.min(0).values, pts.max(0).values
    ext = (mx - mn).clamp_min(0).double().cpu()
    lo = mn.double().cpu()
    # cells of ~2 points: edge from the volume (or area / length when the cloud is flat in some directions)
    live = [float(e) for e in ext if float(e) > 0.0]
    if not live:
        cell = 1.0
    else:
        vol = 1.0
        for e in live:
            vol *= e
        cell = (vol * 2.0 / max(P, 1)) ** (1.0 / len(live))
        cell = max(cell, max(live) / 1024.0, 1e-30)
    res = [max(1, int(float(e) / cell) + 1) for e in ext]
    while res[0] * res[1] * res[2] > _MAX_CELLS:
        cell *= 1.26
        res = [max(1, int(float(e) / cell) + 1) for e in ext]
    nx, ny, nz = res
    origin = [float(v) for v in lo]
    o32 = torch.tensor(origin, dtype=torch.float32, device=pts.device)
    cell32 = float(torch.tensor(cell, dtype=torch.float32))
    ci = torch.floor((pts - o32) * (1.0 / cell32)).to(torch.int64)  # same expression as the kernel's cell lookup
    ci[:, 0].clamp_(0, nx - 1); ci[:, 1].clamp_(0, ny - 1); ci[:, 2].clamp_(0, nz - 1)
    lin = (ci[:, 2] * ny + ci[:, 1]) * nx + ci[:, 0]
    lin_sorted, order = torch.sort(lin)
    cell_start = torch.searchsorted(lin_sorted, torch.arange(nx * ny * nz + 1, device=pts.device, dtype=torch.int64)).to(torch.int32)
    order32 = order.to(torch.int32)
    L = capi.lib()
    with torch.cuda.device(pts.device):
        rc = L.gstar_knn3_mean_dist2(P, C.c_void_p(pts.data_ptr()), C.c_void_p(order32.data_ptr()), C.c_void_p(cell_start.data_ptr()), nx, ny, nz,
                                     C.c_float(float(o32[0])), C.c_float(float(o32[1])), C.c_float(float(o32[2])), C.c_float(cell32),
                                     C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(pts.device).cuda_stream))
    if rc < 0:
        raise RuntimeError(f"gstar_knn3_mean_dist2 failed ({rc})")
    return out
