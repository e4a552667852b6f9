"""Generates tests/golden/config1/sh_rgb.npz -- run HERE (the container that has /root/reference), not on the GPU box.

BASELINE.json config #1 (the reference's CPU-runnable case): 5k-face synthetic mesh -> 30 000 SuGaR-bound Gaussians,
SH evaluation for one 128x128 view on the CPU.  The colours are computed by the REFERENCE's own
gaustar_utils/spherical_harmonics.py::eval_sh exactly as gaustar_scene/sugar_model.py:714-716 does
(clamp_min(eval_sh(deg, sh^T, normalize(points - campos)) + 0.5, 0)), on the scene our generator builds; every 8th
Gaussian is kept.  tests/test_scene_cpu.py rebuilds the scene and checks the oracle's SH->RGB stage against it.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gaustar_b200 import scene  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_sh", "/root/reference/gaustar_utils/spherical_harmonics.py")
ref_sh = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_sh)

g = scene.surface_gaussians(30000, sh_degree=3, seed=0)
cam = scene.dome_cameras(4, 128, 128)[1]
pts = torch.from_numpy(g.means3D)
dirs = torch.nn.functional.normalize(pts - torch.from_numpy(cam.campos).view(1, 3), dim=-1)
sh = torch.from_numpy(g.shs).transpose(-1, -2)  # [P, 3, M]  (sugar_model.py:714)
rgb = torch.clamp_min(ref_sh.eval_sh(3, sh, dirs) + 0.5, 0.0).numpy().astype(np.float32)
idx = np.arange(0, g.P, 8)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1", "sh_rgb.npz")
np.savez_compressed(out, idx=idx.astype(np.int32), rgb=rgb[idx], campos=cam.campos.astype(np.float32), P=np.int32(g.P))
print("wrote", out, rgb[idx].shape, os.path.getsize(out))
