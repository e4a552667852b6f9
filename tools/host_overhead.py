"""Developer tool: where does the time of one fwd+bwd view go on the host vs the GPU?"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import capi, scene, dist as gdist

P, W, H = 1000000, 1920, 1080
g = scene.surface_gaussians(P, sh_degree=3)
cams = scene.dome_cameras(8, W, H)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
base = dict(means3D=t(g.means3D), scales=t(g.scales), rotations=t(g.rotations), bg=torch.tensor([0., 1., 0.], device="cuda"), shs=t(g.shs), sh_degree=3)
opac = t(g.opacities)
kws = [dict(base, viewmatrix=t(c.viewmatrix), projmatrix=t(c.projmatrix), campos=t(c.campos), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy) for c in cams]
dpix = torch.randn(3, H, W, device="cuda") / (W * H)
flat = gdist.FlatGrads(g.P, 16, "cuda")

def run(n, fused, sync_each=False):
    th_f = th_b = 0.0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        kw = kws[i % len(kws)]
        a = time.perf_counter()
        f = capi.forward(opacities=opac, W=W, H=H, **kw)
        b = time.perf_counter()
        capi.backward(f, dpix, accumulate_into=flat.views if fused else None, **kw)
        c = time.perf_counter()
        th_f += b - a; th_b += c - b
        if sync_each: torch.cuda.synchronize()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    return (t1 - t0) / n * 1e3, th_f / n * 1e3, th_b / n * 1e3

for fused in (False, True):
    run(3, fused)
    tot, hf, hb = run(20, fused)
    print(f"fused={fused}: {tot:.3f} ms/view wall | host time in forward() {hf:.3f} ms, in backward() {hb:.3f} ms")
    tot, hf, hb = run(10, fused, sync_each=True)
    print(f"   sync each view: {tot:.3f} ms/view | host fwd {hf:.3f} bwd {hb:.3f}")
print(torch.cuda.memory_summary(abbreviated=True)[:1500])

# ---- two views in flight on two streams (independent views of one multi-view step) ----
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
flats = [gdist.FlatGrads(g.P, 16, "cuda"), gdist.FlatGrads(g.P, 16, "cuda")]
def run2(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        st = streams[i & 1]
        with torch.cuda.stream(st):
            kw = kws[i % len(kws)]
            f = capi.forward(opacities=opac, W=W, H=H, **kw)
            capi.backward(f, dpix, accumulate_into=flats[i & 1].views, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    return (t1 - t0) / n * 1e3
run2(4)
print(f"two streams, fused accumulate into per-stream buffers: {run2(24):.3f} ms/view wall")

s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s_ev.record(); e_ev.record(); torch.cuda.synchronize()
for fused in (False, True):
    for st, nm in ((5, "blend_bwd"), (6, "preprocess_bwd")):
        capi.profile_stage(st, s_ev, e_ev)
        ts = []
        for i in range(6):
            kw = kws[i % len(kws)]
            f = capi.forward(opacities=opac, W=W, H=H, **kw)
            capi.backward(f, dpix, accumulate_into=flat.views if fused else None, **kw)
            torch.cuda.synchronize(); ts.append(s_ev.elapsed_time(e_ev) * 1e3)
        print(f"fused={fused} stage {nm}: {np.median(ts):.1f} us")
capi.profile_stage(-1)
