"""View-sharded run of the reference's training step on N GPUs WITHOUT editing the loop (gaustar_b200.dist.TrainerSharding).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_refine.py

Every rank builds the same SuGaR model (the reference's gaustar_scene/sugar_model.py, byte-compiled into oracle/_ref/pyref by
oracle/build_ref.py; test infrastructure) with the reference's SuGaROptimizer over the parameter groups of sugar_optimizer.py:67-87,
and runs a loop shaped like gaustar_trainers/refine.py:529-548,:794-795 -- `torch.randperm(n_cams)` per pass, ONE view per
iteration through this repository's operator, `optimizer.step(); optimizer.zero_grad(set_to_none=True)` -- that knows nothing of
ranks.  TrainerSharding makes the ranks hold different views at every iteration and sums their gradients before Adam.  Checked and
printed by rank 0 as one JSON line: the views of every iteration are pairwise different across ranks, the replicas' parameters are
BIT-IDENTICAL after the run (same reduced gradients, same Adam state), the loss decreases, bytes reduced per step.
"""
import importlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "shims"), ROOT):
    sys.path.insert(0, p)

from gaustar_b200 import dist as gdist, scene  # noqa: E402
from oracle import pyref  # noqa: E402


def main():
    rank, world, local = gdist.init_from_env()
    torch.cuda.set_device(local)
    iters = int(os.environ.get("ITERS", "48"))
    stub = types.ModuleType("gaussian_splatting.scene.dataset_readers")
    stub.fetchPly = None
    sys.modules["gaussian_splatting.scene.dataset_readers"] = stub
    pyref.install()
    sm = importlib.import_module("gaustar_scene.sugar_model")
    cm = importlib.import_module("gaustar_scene.cameras")
    so = importlib.import_module("gaustar_scene.sugar_optimizer")
    lu = importlib.import_module("gaustar_utils.loss_utils")
    import open3d  # shims/
    verts, faces = scene.capsule_mesh(6000, seed=2)
    cams = scene.dome_cameras(8, 480, 270)
    vcol = np.random.default_rng(4).uniform(0, 1, (len(verts), 3))
    gs_cams = []
    for i, c in enumerate(cams):
        w2c = c.viewmatrix.reshape(4, 4).T.astype(np.float64)
        gs_cams.append(cm.GSCamera(colmap_id=i, R=w2c[:3, :3].T.copy(), T=w2c[:3, 3].copy(), FoVx=2.0 * np.arctan(c.tanfovx), FoVy=2.0 * np.arctan(c.tanfovy),
                                   image=None, gt_alpha_mask=None, image_name=f"img_{i:04d}", uid=i, image_height=c.image_height, image_width=c.image_width))
    wrapper = cm.CamerasWrapper(gs_cams)
    nerf = types.SimpleNamespace(device=torch.device("cuda"), training_cameras=wrapper)

    def make(vertices, colours):
        torch.manual_seed(0)
        return sm.SuGaR(nerfmodel=nerf, points=None, colors=None, initialize=False, sh_levels=3, keep_track_of_knn=False,
                        surface_mesh_to_bind=open3d.TriangleMeshLike(vertices, faces, colours), n_gaussians_per_surface_triangle=6, learn_surface_mesh_opacity=True)

    def render(model, ci):
        return model.render_image_gaussian_rasterizer(camera_indices=ci, bg_color=[0.0, 1.0, 0.0], sh_deg=2, compute_color_in_rasterizer=False,
                                                      compute_covariance_in_rasterizer=True, return_2d_radii=False)

    with torch.no_grad():
        target = make(verts * np.array([1.03, 0.98, 1.02]) + np.array([0.01, -0.015, 0.0]), np.clip(vcol * 0.6 + 0.3, 0, 1))
        gt = [render(target, ci).clone() for ci in range(len(cams))]
        del target
    sugar = make(verts, vcol)
    optimizer = so.SuGaROptimizer(sugar, so.OptimizationParams(iterations=iters, position_lr_max_steps=iters), spatial_lr_scale=wrapper.get_spatial_extent())
    groups = [g["name"] for g in optimizer.optimizer.param_groups]
    seen, losses = [], []
    torch.manual_seed(100 + rank)  # whatever else the trainer draws differs per rank; the shared camera order must not depend on it

    with gdist.TrainerSharding() as ts:
        # ---- from here to the end of the block: the shape of refine.py's loop, nothing rank-aware ----
        iteration = 0
        while iteration < iters:
            shuffled_idx = torch.randperm(len(wrapper.gs_cameras))
            for i in range(0, len(shuffled_idx), 1):
                if iteration >= iters:
                    break
                iteration += 1
                optimizer.update_learning_rate(iteration)
                cmr_i = shuffled_idx[i:i + 1].item()
                seen.append(cmr_i)
                pr = render(sugar, cmr_i).permute(2, 0, 1)[None]
                g_ = gt[cmr_i].permute(2, 0, 1)[None]
                loss = 0.8 * lu.l1_loss(pr, g_) + 0.2 * (1.0 - lu.ssim(pr, g_))
                loss.backward()
                optimizer.step()
                optimizer.zero_grad(set_to_none=True)
                losses.append(float(loss.item()))

    dev = torch.device("cuda", local)
    mine = torch.tensor(seen, device=dev)
    all_seen = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(all_seen, mine)
    all_seen = torch.stack(all_seen).cpu()
    disjoint = all(len(set(all_seen[:, i].tolist())) == world for i in range(all_seen.shape[1]))
    flat = torch.cat([p.detach().reshape(-1).float() for g in optimizer.optimizer.param_groups for p in g["params"]])
    lo, hi = flat.clone(), flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
    lt = torch.tensor(losses, device=dev)
    dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    curve = (lt / world).cpu().numpy()
    if rank == 0:
        print(json.dumps({"tool": "sharded_refine", "world": world, "iterations": iters, "views_per_step": world, "param_groups": groups,
                          "parameters": int(flat.numel()), "allreduce_bytes_per_step": ts.bytes_last, "hook_calls": ts.steps,
                          "views_disjoint_every_iteration": disjoint, "replicas_bit_identical": identical,
                          "loss_first8": float(curve[:8].mean()), "loss_last8": float(curve[-8:].mean()),
                          "views_rank0_first8": all_seen[0, :8].tolist(), "views_rank1_first8": all_seen[min(1, world - 1), :8].tolist()}))
        assert disjoint and identical and curve[-8:].mean() < 0.9 * curve[:8].mean()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
