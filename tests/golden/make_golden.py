"""Generate the golden vectors in tests/golden/ by running the UNMODIFIED reference rasterizer.

Runs on the GPU box only (needs oracle/_ref/libref_dgr.so, built from /root/reference by
oracle/build_ref.py, and a B200):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy *.npz to tests/golden/

Each .npz holds the inputs of one small scene and every output / intermediate buffer of the reference
forward and backward (GeometryState, BinningState, ImageState, the nine gradient tensors).  The CPU
oracle is pinned against these files by tests/test_oracle_golden.py (no GPU needed), and the CUDA
path by tests/test_parity_gpu.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaustar_b200 import scene  # noqa: E402
from oracle import refgpu  # noqa: E402


def cov3d_numpy(scales, rots):
    r, x, y, z = rots[:, 0], rots[:, 1], rots[:, 2], rots[:, 3]
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    S = R * scales[:, None, :]
    Sig = S @ S.transpose(0, 2, 1)
    return np.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], -1).astype(np.float32)


def cases():
    g = scene.surface_gaussians(300, sh_degree=3, seed=2)
    cam = scene.look_at_camera([0.6, 1.2, 1.1], [0, 1.0, 0], 64, 48, fy_over_H=1.0)
    # enlarge the in-plane scales so that the tiny mesh covers many pixels
    g.scales[:, 1:] *= 6.0
    yield "surf_sh3", g, cam, dict(use="sh", deg=3), [0.0, 1.0, 0.0]
    g = scene.random_gaussians(300, sh_degree=2, seed=7, scale_range=(0.02, 0.4))
    cam = scene.look_at_camera([0.3, 1.4, 3.5], [0, 1, 0], 48, 48, fy_over_H=1.1)
    yield "rand_sh2", g, cam, dict(use="sh", deg=2), [0.2, 0.3, 0.4]
    g = scene.random_gaussians(200, sh_degree=0, seed=11, scale_range=(0.01, 0.2))
    cam = scene.look_at_camera([-1.0, 0.8, 2.5], [0, 1, 0], 40, 56, fy_over_H=0.8)
    yield "rand_precomp_cov", g, cam, dict(use="precomp_cov"), [10.0, 10.0, 10.0]
    g = scene.random_gaussians(256, sh_degree=1, seed=13, scale_range=(0.01, 0.15))
    cam = scene.look_at_camera([0.0, 1.0, 2.0], [0, 1, 0], 33, 17, fy_over_H=1.3)
    yield "rand_sh1_deg0", g, cam, dict(use="sh", deg=0), [0.0, 0.0, 0.0]
    # round 2: inputs the callers can reach that round 1's vectors did not cover -- render(..., scaling_modifier)
    # (gaussian_renderer/__init__.py:18,44) and an off-centre principal point (sugar_model.py:1160-1161)
    g = scene.surface_gaussians(300, sh_degree=3, seed=3)
    g.scales[:, 1:] *= 6.0
    cam = scene.look_at_camera([0.5, 1.1, 1.2], [0, 1.0, 0], 64, 48, fy_over_H=1.0)
    yield "surf_sh3_mod05", g, cam, dict(use="sh", deg=3, scale_modifier=0.5), [0.0, 1.0, 0.0]
    g = scene.random_gaussians(300, sh_degree=2, seed=17, scale_range=(0.02, 0.2))
    cam = scene.look_at_camera([0.3, 1.4, 3.5], [0, 1, 0], 48, 48, fy_over_H=1.1)
    yield "rand_sh2_mod2", g, cam, dict(use="sh", deg=2, scale_modifier=2.0), [0.2, 0.3, 0.4]
    g = scene.surface_gaussians(300, sh_degree=3, seed=5)
    g.scales[:, 1:] *= 6.0
    cam = scene.look_at_camera([0.6, 1.2, 1.1], [0, 1.0, 0], 64, 48, fy_over_H=1.0, principal_ndc=(0.23, -0.17))
    yield "surf_sh3_offcentre", g, cam, dict(use="sh", deg=3), [0.0, 1.0, 0.0]


def main(out_dir, only=None):
    os.makedirs(out_dir, exist_ok=True)
    dev = "cuda"
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name, g, cam, mode, bg in cases():
        if only and name not in only:
            continue
        W, H = cam.image_width, cam.image_height
        rng = np.random.default_rng(5)
        inputs = dict(means3D=g.means3D, opacities=g.opacities, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos,
                      bg=np.asarray(bg, np.float32), tan_fovx=np.float32(cam.tanfovx), tan_fovy=np.float32(cam.tanfovy), W=W, H=H,
                      scale_modifier=np.float32(mode.get("scale_modifier", 1.0)), sh_degree=0)
        if mode["use"] == "sh":
            inputs.update(shs=g.shs, scales=g.scales, rotations=g.rotations, sh_degree=mode["deg"])
        else:
            inputs.update(colors_precomp=rng.uniform(0, 1, (g.P, 3)).astype(np.float32), cov3D_precomp=cov3d_numpy(g.scales, g.rotations))
        kw = {}
        for k, v in inputs.items():
            if isinstance(v, np.ndarray) and v.ndim > 0:
                kw[k] = t(v)
            elif k in ("tan_fovx", "tan_fovy", "scale_modifier"):
                kw[k] = float(v)
            else:
                kw[k] = int(v)
        fwd = refgpu.forward(**kw)
        dpix = rng.normal(0, 1, (3, H, W)).astype(np.float32)
        bkw = {k: v for k, v in kw.items() if k not in ("opacities", "W", "H")}
        grads = refgpu.backward(fwd, t(dpix), **bkw)
        out = {"in_" + k: v for k, v in inputs.items()}
        out["in_dL_dpix"] = dpix
        for k, v in fwd.items():
            out["fwd_" + k] = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.int64(v)
        for k, v in grads.items():
            out["bwd_" + k] = v.cpu().numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "P", g.P, f"{W}x{H}", "R", fwd["num_rendered"], "visible", int((fwd["radii"] > 0).sum()), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"), only=sys.argv[2:] or None)
