"""Developer tool: run a few fwd+bwd views of the headline workload (for ncu launch lists / captures).

    ncu ... python tools/profile_view.py --views 3 [--P 1000000 --W 1920 --H 1080 --precomp]
"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import capi, scene

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1000000)
ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--H", type=int, default=1080)
ap.add_argument("--views", type=int, default=3)
ap.add_argument("--precomp", action="store_true")
ap.add_argument("--random", action="store_true")
a = ap.parse_args()
g = scene.random_gaussians(a.P, 3, seed=1, scale_range=(0.003, 0.05)) if a.random else scene.surface_gaussians(a.P, sh_degree=3)
cams = scene.dome_cameras(max(a.views, 2), a.W, a.H)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
base = dict(means3D=t(g.means3D), opacities=t(g.opacities), scales=t(g.scales), rotations=t(g.rotations), bg=torch.tensor([0., 1., 0.], device="cuda"))
if a.precomp:
    base["colors_precomp"] = torch.rand(g.P, 3, device="cuda")
else:
    base.update(shs=t(g.shs), sh_degree=3)
dpix = torch.randn(3, a.H, a.W, device="cuda") / (a.W * a.H)
for v in range(a.views):
    c = cams[v % len(cams)]
    kw = dict(base, viewmatrix=t(c.viewmatrix), projmatrix=t(c.projmatrix), campos=t(c.campos), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy, W=a.W, H=a.H)
    f = capi.forward(**kw)
    capi.backward(f, dpix, **{k: v2 for k, v2 in kw.items() if k not in ("opacities", "W", "H")})
    torch.cuda.synchronize()
    print("view", v, "R", f["num_rendered"])
