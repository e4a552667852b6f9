"""Developer tool (GPU box): fwd+bwd time per view of the rasterizer at the Gaussian counts and resolutions of
BASELINE.json's configs #2..#5, this implementation vs the unmodified reference (oracle/_ref) on the SAME tensors, one
stream, CUDA events around `views` back-to-back views, both through their pybind modules with the reference's own
signatures (`_C.rasterize_gaussians` / `_C.rasterize_gaussians_backward`: ours = gaustar_b200._C, reference = the stock
binding built under oracle/_ref).

    python tools/config_table.py [--views 8 --out gpurun_out/config_table.json]
"""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import scene
from gaustar_b200.rasterizer import _C as OURS
from oracle import refgpu

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--out", default="")
ap.add_argument("--only", default="")
a = ap.parse_args()

CONFIGS = [  # (name, P, sh_degree, W, H)
    ("#2 200k 1920x1080 SH3", 200000, 3, 1920, 1080),
    ("#3 1M 1352x1014 SH3", 1000000, 3, 1352, 1014),
    ("#4 500k 1352x1014 SH2", 500000, 2, 1352, 1014),
    ("headline 1M 1920x1080 SH3", 1000000, 3, 1920, 1080),
    ("#5 4M 3840x2160 SH3", 4000000, 3, 3840, 2160),
]
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
rows = []
for name, P, deg, W, H in CONFIGS:
    if a.only and a.only not in name:
        continue
    g = scene.surface_gaussians(P, sh_degree=deg)
    cams = scene.dome_cameras(max(a.views, 2), W, H)
    base = dict(means3D=t(g.means3D), opacities=t(g.opacities), scales=t(g.scales), rotations=t(g.rotations), shs=t(g.shs),
                bg=torch.tensor([0., 1., 0.], device="cuda"))
    camkw = [dict(viewmatrix=t(c.viewmatrix).view(4, 4), projmatrix=t(c.projmatrix).view(4, 4), campos=t(c.campos), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy)
             for c in cams]
    dpix = torch.randn(3, H, W, device="cuda") / (W * H)

    e = torch.Tensor([])

    def one(mod, v):
        c = camkw[v % len(camkw)]
        R, color, radii, gb, bb, ib = mod.rasterize_gaussians(base["bg"], base["means3D"], e, base["opacities"], base["scales"], base["rotations"], 1.0,
                                                              e, c["viewmatrix"], c["projmatrix"], c["tan_fovx"], c["tan_fovy"], H, W, base["shs"], deg,
                                                              c["campos"], False, False)
        mod.rasterize_gaussians_backward(base["bg"], base["means3D"], radii, e, base["scales"], base["rotations"], 1.0, e, c["viewmatrix"],
                                         c["projmatrix"], c["tan_fovx"], c["tan_fovy"], dpix, base["shs"], deg, c["campos"], gb, R, bb, ib, False)
        return R

    def timed(mod):
        for v in range(3):
            R = one(mod, v)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for v in range(a.views):
            R = one(mod, v)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.views, int(R)

    ours_ms, R = timed(OURS)
    row = dict(config=name, gaussians=g.P, resolution=[W, H], sh_degree=deg, num_rendered_last_view=R, ours_ms_per_view=round(ours_ms, 4))
    if refgpu.available():
        ref_ms, R_ref = timed(refgpu.stock_module())
        row.update(reference_ms_per_view=round(ref_ms, 4), speedup=round(ref_ms / ours_ms, 3), num_rendered_equal=bool(R == R_ref))
    rows.append(row)
    print(json.dumps(row), flush=True)
    del base, camkw, dpix, g
    torch.cuda.empty_cache()
res = dict(what="fwd+bwd ms per view, one stream, pybind modules with the reference signatures on both sides, same tensors; surface Gaussians + dome cameras (bench.py's generator)",
           views=a.views, device=torch.cuda.get_device_name(0), rows=rows)
if a.out:
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
