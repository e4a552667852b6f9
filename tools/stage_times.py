"""Developer tool: per-stage kernel times of the headline workload, measured with the library's CUDA-event hook
(one stage per pass; single stream, so the times are not inflated by overlap).

    python tools/stage_times.py [--P 1000000 --W 1920 --H 1080 --views 6 --precomp --random --no-log]
"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import capi, scene

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1000000)
ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--H", type=int, default=1080)
ap.add_argument("--views", type=int, default=6)
ap.add_argument("--precomp", action="store_true")
ap.add_argument("--random", action="store_true")
ap.add_argument("--no-log", action="store_true")
ap.add_argument("--dual", action="store_true", help="two feature passes in one blend (gstar_fwd_args::colors2)")
ap.add_argument("--ch2", type=int, default=3, help="channels of the second pass with --dual (1..4)")
a = ap.parse_args()
if a.no_log:
    capi.set_hit_log(0)
g = scene.random_gaussians(a.P, 3, seed=1, scale_range=(0.003, 0.05)) if a.random else scene.surface_gaussians(a.P, sh_degree=3)
cams = scene.dome_cameras(max(a.views, 2), a.W, a.H)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
base = dict(means3D=t(g.means3D), opacities=t(g.opacities), scales=t(g.scales), rotations=t(g.rotations), bg=torch.tensor([0., 1., 0.], device="cuda"))
if a.precomp:
    base["colors_precomp"] = torch.rand(g.P, 3, device="cuda")
else:
    base.update(shs=t(g.shs), sh_degree=3)
dpix = torch.randn(3, a.H, a.W, device="cuda") / (a.W * a.H)
camkw = [dict(viewmatrix=t(c.viewmatrix), projmatrix=t(c.projmatrix), campos=t(c.campos), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy) for c in cams]


col2 = torch.rand(g.P, a.ch2, device="cuda") * 5 if a.dual else None
bg2 = torch.full((a.ch2,), 10.0, device="cuda")
dpix2 = torch.randn(a.ch2, a.H, a.W, device="cuda") / (a.W * a.H)


def one(v):
    kw = dict(base, W=a.W, H=a.H, **camkw[v % len(camkw)])
    bk = {k: v2 for k, v2 in kw.items() if k not in ("opacities", "W", "H")}
    if a.dual:
        f = capi.forward(colors2=col2, bg2=bg2, **kw)
        capi.backward(f, dpix, dL_dout_color2=dpix2, colors2=col2, bg2=bg2, **bk)
    else:
        f = capi.forward(**kw)
        capi.backward(f, dpix, **bk)
    return f


for v in range(3):
    f = one(v)
torch.cuda.synchronize()
need, cap, used = capi.hit_log_state(f)
print(capi.debug_header(f))
print(f"R={f['num_rendered']} hit-log slots needed={need} capacity={cap} in_use={used}")
tot = 0.0
for si, name in enumerate(capi.STAGES):
    ts = []
    for v in range(a.views):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record()  # torch creates the CUDA event lazily on first record
        capi.profile_stage(si, e0, e1)
        one(v)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    capi.profile_stage(-1)
    tot += float(np.mean(ts))
    print(f"{name:16s} {np.mean(ts):8.1f} us  (min {np.min(ts):.1f} max {np.max(ts):.1f})")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for v in range(a.views):
    one(v)
e1.record()
torch.cuda.synchronize()
print(f"sum of stages {tot:.1f} us; wall per view (1 stream, incl. host) {e0.elapsed_time(e1) * 1e3 / a.views:.1f} us")
