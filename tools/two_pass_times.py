"""Developer tool: the RGB + depth training-step pattern of gaustar_trainers/refine.py:552-616 (two rasterizations of the same
Gaussians and camera per view: SH colours, then depth as three equal channels with another background, one backward
through both) timed with two full calls versus inside shared_geometry(), through the operator API.

    python tools/two_pass_times.py [--P 1000000 --W 1920 --H 1080 --views 12 --out gpurun_out/two_pass.json]
"""
import argparse, contextlib, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diff_gaussian_rasterization as dgr
from gaustar_b200 import scene

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1000000)
ap.add_argument("--W", type=int, default=1920)
ap.add_argument("--H", type=int, default=1080)
ap.add_argument("--views", type=int, default=12)
ap.add_argument("--passes", type=int, default=2, help="feature passes per view in the training step (config #5: 3 = RGB, depth, normal)")
ap.add_argument("--out", default="")
ap.add_argument("--depth-one-channel", action="store_true", help="forward_passes variant: depth as ONE channel (and the normal as three): all extra "
                "channels ride in the same blend as the first pass (seven channels)")
a = ap.parse_args()

g = scene.surface_gaussians(a.P, sh_degree=3)
cams = scene.dome_cameras(max(a.views, 2), a.W, a.H)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
params = {k: t(getattr(g, k)).requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
camkw = [dict(vm=t(c.viewmatrix).view(4, 4), pm=t(c.projmatrix).view(4, 4), campos=t(c.campos), tx=c.tanfovx, ty=c.tanfovy) for c in cams]
bg1 = torch.tensor([0., 1., 0.], device="cuda")
bg2 = torch.full((3,), 10.0, device="cuda")
target = torch.rand(3, a.H, a.W, device="cuda")
dtarget = torch.rand(a.H, a.W, device="cuda") * 5
center = torch.tensor([0.0, 1.0, 0.0], device="cuda")


def settings(c, bg, deg):
    return dgr.GaussianRasterizationSettings(a.H, a.W, c["tx"], c["ty"], bg, 1.0, c["vm"], c["pm"], deg, c["campos"], False, False)


def step_fused(v):
    """The same step through GaussianRasterizer.forward_passes: one autograd node for all passes."""
    c = camkw[v % len(camkw)]
    p = params
    depth = (p["means3D"] @ c["vm"][:3, 2] + c["vm"][3, 2])[:, None]
    passes = [(depth.contiguous(), bg2[:1])] if a.depth_one_channel else [(depth.expand(-1, 3), bg2)]
    for _k in range(a.passes - 2):
        passes.append((torch.nn.functional.normalize(p["means3D"] - center, dim=-1), bg1))
    img, _, extra = dgr.GaussianRasterizer(settings(c, bg1, 3)).forward_passes(
        means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"], requires_grad=True), opacities=p["opacities"], shs=p["shs"],
        scales=p["scales"], rotations=p["rotations"], extra_passes=passes)
    loss = (img - target).abs().mean() + (extra[0][0] - dtarget).abs().mean()
    for e in extra[1:]:
        loss = loss + (e - target).abs().mean()
    loss.backward()
    return loss


def step(v, shared):
    if shared == "fused":
        return step_fused(v)
    c = camkw[v % len(camkw)]
    p = params
    with (dgr.shared_geometry() if shared else contextlib.nullcontext()):
        img, _ = dgr.GaussianRasterizer(settings(c, bg1, 3))(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"], requires_grad=True),
                                                            opacities=p["opacities"], shs=p["shs"], scales=p["scales"], rotations=p["rotations"])
        depth = (p["means3D"] @ c["vm"][:3, 2] + c["vm"][3, 2])[:, None].expand(-1, 3)
        dimg, _ = dgr.GaussianRasterizer(settings(c, bg2, 0))(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"], requires_grad=True),
                                                             opacities=p["opacities"], colors_precomp=depth, scales=p["scales"],
                                                             rotations=p["rotations"])
        extra = []
        for _k in range(a.passes - 2):  # config #5's third pass: a per-Gaussian direction as colour (normal xyz)
            nrm = torch.nn.functional.normalize(p["means3D"] - center, dim=-1)
            extra.append(dgr.GaussianRasterizer(settings(c, bg1, 0))(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"], requires_grad=True),
                                                                    opacities=p["opacities"], colors_precomp=nrm, scales=p["scales"],
                                                                    rotations=p["rotations"])[0])
    loss = (img - target).abs().mean() + (dimg[0] - dtarget).abs().mean()
    for e in extra:
        loss = loss + (e - target).abs().mean()
    loss.backward()
    return loss


def timed(shared):
    for q in params.values():
        q.grad = None
    for v in range(3):
        step(v, shared)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for v in range(a.views):
        step(v, shared)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.views, {k: q.grad.clone() for k, q in params.items()}


def timed_inference(shared, passes=3):
    """refined_mesh.py:733-774: three forward-only renders of one camera (RGB through SH, then two colors_precomp passes)."""
    p = {k: v.detach() for k, v in params.items()}
    cols = [torch.rand(g.P, 3, device="cuda", generator=torch.Generator("cuda").manual_seed(100 + i)) for i in range(passes - 1)]

    def view(v):
        c = camkw[v % len(camkw)]
        with torch.no_grad(), (dgr.shared_geometry() if shared else contextlib.nullcontext()):
            outs = [dgr.GaussianRasterizer(settings(c, bg1, 3))(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]), opacities=p["opacities"],
                                                               shs=p["shs"], scales=p["scales"], rotations=p["rotations"])[0]]
            for col in cols:
                outs.append(dgr.GaussianRasterizer(settings(c, bg2, 0))(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]),
                                                                       opacities=p["opacities"], colors_precomp=col, scales=p["scales"],
                                                                       rotations=p["rotations"])[0])
        return outs

    for v in range(3):
        view(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for v in range(a.views):
        outs = view(v)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.views, outs


full_ms, g_full = timed(False)
shared_ms, g_shared = timed(True)
full_ms2, _ = timed(False)
fused_ms, g_fused = timed("fused")
rel = {k: float((g_shared[k] - g_full[k]).abs().max() / (g_full[k].abs().max() + 1e-30)) for k in g_full}
res = dict(workload=f"surface P={g.P} {a.W}x{a.H} SH3: {a.passes} feature passes (RGB through SH, then colors_precomp) + one backward through all, per view",
           views=a.views, full_calls_ms_per_view=round(min(full_ms, full_ms2), 4), shared_geometry_ms_per_view=round(shared_ms, 4),
           speedup=round(min(full_ms, full_ms2) / shared_ms, 4), grad_rel_diff_accumulated_over_views=rel, device=torch.cuda.get_device_name(0),
           forward_passes_ms_per_view=round(fused_ms, 4), forward_passes_speedup=round(min(full_ms, full_ms2) / fused_ms, 4),
           forward_passes_grad_rel_diff={k: float((g_fused[k] - g_full[k]).abs().max() / (g_full[k].abs().max() + 1e-30)) for k in g_full})
inf_full, o_full = timed_inference(False)
inf_shared, o_shared = timed_inference(True)
res["inference_3_passes"] = dict(three_full_calls_ms_per_view=round(inf_full, 4), shared_geometry_ms_per_view=round(inf_shared, 4),
                                 speedup=round(inf_full / inf_shared, 4), images_identical=bool(all(torch.equal(x, y) for x, y in zip(o_full, o_shared))))
print(json.dumps(res))
if a.out:
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
