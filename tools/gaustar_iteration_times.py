"""Developer tool (GPU box): time GauSTAR's training iteration -- the reference's own SuGaR model (byte-compiled, oracle/_ref/pyref)
rendering RGB + depth and back-propagating, as refine.py:552-616 does -- with the reference rasterizer and with this repository's,
then with the one-line opt-ins: patch_sugar (fused prologue), SH evaluated in the rasterizer, shared_geometry().

    python tools/gaustar_iteration_times.py [--faces 83334 --W 1352 --H 1014 --iters 40]
"""
import argparse, importlib, json, os, sys, time, types
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_zz_reference_callers_gpu as T  # the import contexts of the two arms
from gaustar_b200 import scene
from gaustar_b200 import sugar as gsugar

ap = argparse.ArgumentParser()
ap.add_argument("--faces", type=int, default=83334); ap.add_argument("--W", type=int, default=1352); ap.add_argument("--H", type=int, default=1014)
ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
verts, faces = scene.capsule_mesh(a.faces, seed=2)
cams = scene.dome_cameras(8, a.W, a.H)
vcol = np.random.default_rng(4).uniform(0, 1, (len(verts), 3))
out = {}
for name, op in T._arms():
    with T._Arm(op):
        sm = importlib.import_module("gaustar_scene.sugar_model"); cm = importlib.import_module("gaustar_scene.cameras")
        import open3d
        gs = []
        for i, c in enumerate(cams):
            w2c = c.viewmatrix.reshape(4, 4).T.astype(np.float64)
            gs.append(cm.GSCamera(colmap_id=i, R=w2c[:3, :3].T.copy(), T=w2c[:3, 3].copy(), FoVx=2 * np.arctan(c.tanfovx), FoVy=2 * np.arctan(c.tanfovy), image=None,
                                  gt_alpha_mask=None, image_name=f"i{i}", uid=i, image_height=c.image_height, image_width=c.image_width))
        wrapper = cm.CamerasWrapper(gs)
        nerf = types.SimpleNamespace(device=torch.device("cuda"), training_cameras=wrapper)
        model = sm.SuGaR(nerfmodel=nerf, points=None, colors=None, initialize=False, sh_levels=3, keep_track_of_knn=False,
                         surface_mesh_to_bind=open3d.TriangleMeshLike(verts, faces, vcol), n_gaussians_per_surface_triangle=6, learn_surface_mesh_opacity=True)
        with torch.no_grad():
            model.all_densities.copy_(torch.logit(torch.full_like(model.all_densities, 0.9)))
        params = [p for p in model.parameters() if p.requires_grad]
        tgt = torch.rand(a.H, a.W, 3, device="cuda")

        def iteration(ci, color_in_rasterizer, ctx):
            with ctx():
                rgb = model.render_image_gaussian_rasterizer(camera_indices=ci, bg_color=[0.0, 1.0, 0.0], sh_deg=2, compute_color_in_rasterizer=color_in_rasterizer)
                depth_pts = wrapper.p3d_cameras[ci].get_world_to_view_transform().transform_points(model.points)[..., 2:].expand(-1, 3)
                depth = model.render_image_gaussian_rasterizer(camera_indices=ci, bg_color=10.0 + torch.zeros(3, device="cuda"), sh_deg=0, point_colors=depth_pts)[..., 0]
            loss = (rgb - tgt).abs().mean() + (depth - 3.0).abs().mean()
            loss.backward()
            for p in params:
                p.grad = None
            return float(loss.item())

        import contextlib
        variants = [("unchanged", False, contextlib.nullcontext, False)]
        if name == "ours":
            variants += [("patch_sugar", False, contextlib.nullcontext, True), ("patch_sugar + SH in rasterizer", True, contextlib.nullcontext, True),
                         ("patch_sugar + SH in rasterizer + shared_geometry", True, op.shared_geometry, True)]
        for vname, cir, ctx, patch in variants:
            if patch:
                gsugar.patch_sugar(model)
            for it in range(5):
                iteration(it % len(cams), cir, ctx)
            torch.cuda.synchronize(); t0 = time.time()
            for it in range(a.iters):
                iteration(it % len(cams), cir, ctx)
            torch.cuda.synchronize()
            out[f"{name}: {vname}"] = round((time.time() - t0) / a.iters * 1e3, 3)
            gsugar.unpatch_sugar(model)
print(json.dumps({"workload": f"SuGaR bound to a {a.faces}-face mesh ({a.faces * 6} Gaussians), {a.W}x{a.H}, RGB + depth render, L1 losses, backward; ms per iteration",
                  "ms_per_iteration": out}, indent=1))
