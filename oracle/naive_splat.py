"""Naive PyTorch point-splat on the host CPU: the CPU baseline BASELINE.md 2.2 / SURVEY 8d name (TEST / BENCH INFRASTRUCTURE).

"project centres, per-Gaussian bounding box, dense per-pixel alpha evaluation in chunks, sort by depth with torch.sort,
cumprod transmittance; autograd for backward" -- nothing of the rasterizer's machinery (no key sort, no early termination, no
hand-written backward).  The reference rasterizer has no CPU path; this stands in as "what one would write in PyTorch on the CPU".
The per-pixel formula is the reference's (forward.cu:329-358: alpha = min(0.99, o exp(-0.5 d^T conic d)), alpha < 1/255 dropped,
front-to-back compositing), the EWA projection follows forward.cu:74-113 / 118-152; there is no T < 1e-4 cut, so images agree with
the rasterizer to ~1e-4, not bit for bit.  Only bench.py's cpu_baseline leg and the tests use it.
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, dirs):
    """sh [P, M, 3], dirs [P, 3] (unit) -> [P, 3]; the polynomial basis of forward.cu:20-71 / spherical_harmonics.py:117-178."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * sh[:, 0]
    if deg > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5] + C2[2] * (2 * zz - xx - yy) * sh[:, 6] + C2[3] * xz * sh[:, 7] + C2[4] * (xx - yy) * sh[:, 8]
        if deg > 2:
            res = res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10] + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] \
                + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12] + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14] \
                + C3[6] * x * (xx - 3 * yy) * sh[:, 15]
    return torch.clamp_min(res + 0.5, 0.0)


def splat(means3D, scales, rotations, opacities, shs, sh_degree, viewmatrix, projmatrix, campos, tanfovx, tanfovy, W, H, bg, tile=16):
    """All tensors torch CPU float32 (viewmatrix / projmatrix: the op's transposed 4x4).  Returns the [3, H, W] image."""
    V, PV = viewmatrix.reshape(4, 4), projmatrix.reshape(4, 4)
    pv = means3D @ V[:3, :3] + V[3, :3]
    depth = pv[:, 2]
    keep = torch.nonzero(depth > 0.2)[:, 0]
    order = keep[torch.sort(depth[keep])[1]]  # front to back, once for the whole image
    means3D, scales, rotations, opacities, shs, pv, depth = (t[order] for t in (means3D, scales, rotations, opacities, shs, pv, depth))
    ph = torch.cat([means3D, torch.ones_like(means3D[:, :1])], 1) @ PV
    ndc = ph[:, :2] / (ph[:, 3:4] + 1e-7)
    cx = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    cy = ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5
    # covariance: Sigma = R S S^T R^T, projected with J W (EWA)
    q = rotations
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * scales[:, None, :]
    Sigma = M @ M.transpose(1, 2)
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    tz = pv[:, 2]
    tx = torch.minimum(torch.maximum(pv[:, 0] / tz, torch.tensor(-1.3 * tanfovx)), torch.tensor(1.3 * tanfovx)) * tz
    ty = torch.minimum(torch.maximum(pv[:, 1] / tz, torch.tensor(-1.3 * tanfovy)), torch.tensor(1.3 * tanfovy)) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -fx * tx / (tz * tz), zero, fy / tz, -fy * ty / (tz * tz)], -1).reshape(-1, 2, 3)
    Wm = V[:3, :3].T  # world -> camera rotation
    T = J @ Wm
    cov = T @ Sigma @ T.transpose(1, 2)
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    conA, conB, conC = c / det, -b / det, a / det
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))
    radius = torch.ceil(3.0 * torch.sqrt(lam)).detach()
    dirs = torch.nn.functional.normalize(means3D - campos.reshape(1, 3), dim=-1)
    colors = eval_sh(sh_degree, shs, dirs)
    opac = opacities.reshape(-1)
    x0, x1, y0, y1 = (cx - radius).detach(), (cx + radius).detach(), (cy - radius).detach(), (cy + radius).detach()
    rows = []
    for ty0 in range(0, H, tile):
        row = []
        th = min(tile, H - ty0)
        in_rows = (y1 >= ty0) & (y0 <= ty0 + th - 1)
        for tx0 in range(0, W, tile):
            tw = min(tile, W - tx0)
            sel = torch.nonzero(in_rows & (x1 >= tx0) & (x0 <= tx0 + tw - 1))[:, 0]  # still in depth order
            px = torch.arange(tx0, tx0 + tw, dtype=torch.float32)
            py = torch.arange(ty0, ty0 + th, dtype=torch.float32)
            if sel.numel() == 0:
                row.append(bg.reshape(3, 1, 1).expand(3, th, tw))
                continue
            dx = cx[sel][:, None, None] - px[None, None, :]
            dy = cy[sel][:, None, None] - py[None, :, None]
            power = -0.5 * (conA[sel][:, None, None] * dx * dx + conC[sel][:, None, None] * dy * dy) - conB[sel][:, None, None] * dx * dy
            alpha = torch.clamp_max(opac[sel][:, None, None] * torch.exp(power), 0.99)
            alpha = torch.where((power > 0) | (alpha < 1.0 / 255.0), torch.zeros_like(alpha), alpha)
            trans = torch.cumprod(1.0 - alpha, dim=0)
            t_before = torch.cat([torch.ones_like(trans[:1]), trans[:-1]], 0)
            wgt = alpha * t_before
            img = torch.einsum("gyx,gc->cyx", wgt, colors[sel]) + trans[-1][None] * bg.reshape(3, 1, 1)
            row.append(img)
        rows.append(torch.cat(row, 2))
    return torch.cat(rows, 1)


def time_fwd_bwd(P, W, H, sh_degree=3, views=1, seed=0, threads=None):
    """views/s of forward + autograd backward on the same synthetic generator as the GPU workload (bench.py)."""
    import os
    from gaustar_b200 import scene
    torch.set_num_threads(threads or os.cpu_count())
    g = scene.surface_gaussians(P, sh_degree=sh_degree, seed=seed)
    cams = scene.dome_cameras(max(views, 2), W, H)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    leaves = {k: t(getattr(g, k)).requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    bg = torch.tensor([0.0, 1.0, 0.0])
    rng = np.random.default_rng(1)
    t0 = time.time()
    for v in range(views):
        c = cams[v]
        img = splat(leaves["means3D"], leaves["scales"], leaves["rotations"], leaves["opacities"], leaves["shs"], sh_degree, t(c.viewmatrix),
                    t(c.projmatrix), t(c.campos), c.tanfovx, c.tanfovy, W, H, bg)
        target = torch.from_numpy(np.clip(rng.normal(0.5, 0.2, (3, H, W)), 0, 1).astype(np.float32))
        (img - target).abs().mean().backward()
    dt = time.time() - t0
    return views / dt, dt, g.P
