import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import capi, scene, dist as gdist
P, W, H = 1000000, 1920, 1080
g = scene.surface_gaussians(P, sh_degree=3)
cams = scene.dome_cameras(32, W, H)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
base = dict(means3D=t(g.means3D), scales=t(g.scales), rotations=t(g.rotations), bg=torch.tensor([0., 1., 0.], device="cuda"), shs=t(g.shs), sh_degree=3)
opac = t(g.opacities)
kws = [dict(base, viewmatrix=t(c.viewmatrix), projmatrix=t(c.projmatrix), campos=t(c.campos), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy) for c in cams]
dpix = torch.randn(3, H, W, device="cuda") / (W * H)
flat = gdist.FlatGrads(g.P, 16, "cuda")
def stats():
    s = torch.cuda.memory_stats()
    return s["segment.all.allocated"], s["segment.all.freed"], s["num_alloc_retries"], s["reserved_bytes.all.current"] >> 20
for rep in range(3):
    s0 = stats(); torch.cuda.synchronize(); t0 = time.perf_counter()
    Rs = []
    for i in range(32):
        kw = kws[i % len(kws)]
        f = capi.forward(opacities=opac, W=W, H=H, **kw)
        capi.backward(f, dpix, accumulate_into=flat.views, **kw)
        Rs.append(f["num_rendered"])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"rep {rep}: {(t1-t0)/32*1e3:.3f} ms/view; segments alloc/free/retries/reservedMB before {s0} after {stats()}; R min/max {min(Rs)} {max(Rs)}")
