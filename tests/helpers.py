"""Shared helpers for the tests (not product code)."""
import glob
import os

import numpy as np

from gaustar_b200 import scene
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GRAD_KEYS = ("dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(path):
    z = np.load(path)
    inp = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    fwd = {k[4:]: z[k] for k in z.files if k.startswith("fwd_")}
    bwd = {k[4:]: z[k] for k in z.files if k.startswith("bwd_")}
    return inp, fwd, bwd


def oracle_inputs_from_dict(d) -> O.Inputs:
    g = lambda k: d[k] if k in d else None
    return O.Inputs(means3D=d["means3D"], opacities=d["opacities"], viewmatrix=d["viewmatrix"], projmatrix=d["projmatrix"], campos=d["campos"],
                    bg=d["bg"], tan_fovx=float(d["tan_fovx"]), tan_fovy=float(d["tan_fovy"]), W=int(d["W"]), H=int(d["H"]), shs=g("shs"),
                    colors_precomp=g("colors_precomp"), scales=g("scales"), rotations=g("rotations"), cov3D_precomp=g("cov3D_precomp"),
                    scale_modifier=float(d.get("scale_modifier", 1.0)), sh_degree=int(d.get("sh_degree", 0)))


def scene_dict(g: scene.Gaussians, cam: scene.Camera, use_sh=True, sh_degree=None, bg=(0.0, 1.0, 0.0), seed=3):
    d = dict(means3D=g.means3D, opacities=g.opacities, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos,
             bg=np.asarray(bg, np.float32), tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, W=cam.image_width, H=cam.image_height,
             scales=g.scales, rotations=g.rotations, scale_modifier=1.0, sh_degree=0)
    if use_sh:
        d["shs"] = g.shs
        d["sh_degree"] = int(round(g.shs.shape[1] ** 0.5)) - 1 if sh_degree is None else sh_degree
    else:
        d["colors_precomp"] = np.random.default_rng(seed).uniform(0, 1, (g.P, 3)).astype(np.float32)
    return d


def to_torch_kwargs(d, dev="cuda"):
    import torch
    kw = {}
    for k, v in d.items():
        if isinstance(v, np.ndarray) and v.ndim > 0:
            kw[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dev)
        elif k in ("W", "H", "sh_degree"):
            kw[k] = int(v)
        else:
            kw[k] = float(v)
    return kw


def bwd_kwargs(kw):
    return {k: v for k, v in kw.items() if k not in ("opacities", "W", "H")}


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
