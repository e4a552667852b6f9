"""Synthetic on-disk dataset in the layout GauSTAR's loaders read -- the WRITERS only.

Host-side numpy / PIL code, no GPU: it turns the bench's synthetic scene (gaustar_b200/scene.py: capsule mesh, mesh-bound
Gaussians, dome cameras) and per-view images / masks / depths / flows that a caller rendered into the files
`train_seq.py` expects, so that the reference's refine loop can later be pointed at them.  The dependency shims that loop
also needs (pytorch3d, open3d, ...) are NOT here; see DESIGN.md section 9.

Layout (R = sequence root, F = frame index, V = camera index; all paths as the reference builds them):
  R/rgb_cameras.npz                                  ids, intrinsics [N,3,3], extrinsics [N,3,4], dist_coeffs, shape [N,2]=(h,w)
                                                      data_process/ahq2gaustar.py:12-47
  R/FFFF/images/img_VVVV.jpg                         gaustar_scene/cameras.py:33,76-78
  R/FFFF/masks_humanrf/img_VVVV_alpha.png            cameras.py:97-98
  R/FFFF/depth_humanrf/img_VVVV_depth.npz['depth']   cameras.py:102-106, gaustar_tools/warp_mesh.py:279-281
  R/FFFF/flow_bi/VVVV_f.npz['flow'], VVVV_b.npz      warp_mesh.py:269-271 (stored (dy, dx): the loader flips the last axis)
  G/cameras.json                                     gaussian_splatting/utils/camera_utils.py:70-90, read by cameras.py:35-73
  G/point_cloud/iteration_N/point_cloud.ply          gaussian_splatting/scene/gaussian_model.py:177-213 (write), :215-256 (read)
  R/init_mesh.obj                                    plain Wavefront OBJ (v / f, 1-based)
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import scene


def camera_pose(cam: scene.Camera):
    """(R, T) in the 3DGS convention from a scene.Camera: viewmatrix = getWorld2View(R, T).T with Rt[:3,:3] = R.T,
    Rt[:3,3] = T (gaussian_splatting/utils/graphics_utils.py:38-44)."""
    w2v = np.asarray(cam.viewmatrix, np.float64).T
    return w2v[:3, :3].T.copy(), w2v[:3, 3].copy()


def camera_to_json(idx: int, cam: scene.Camera, img_name: str) -> dict:
    """camera_utils.py:70-90: 'position' / 'rotation' are the camera-to-world translation and rotation, fx/fy = fov2focal."""
    R, T = camera_pose(cam)
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = T
    Rt[3, 3] = 1.0
    c2w = np.linalg.inv(Rt)
    return {"id": idx, "img_name": img_name, "width": int(cam.image_width), "height": int(cam.image_height),
            "position": c2w[:3, 3].tolist(), "rotation": [r.tolist() for r in c2w[:3, :3]],
            "fy": float(cam.image_height / (2.0 * cam.tanfovy)), "fx": float(cam.image_width / (2.0 * cam.tanfovx))}


def write_cameras_json(gs_out: str, cams, names=None) -> str:
    os.makedirs(gs_out, exist_ok=True)
    names = names or [f"img_{i:04d}" for i in range(len(cams))]
    path = os.path.join(gs_out, "cameras.json")
    with open(path, "w") as f:
        json.dump([camera_to_json(i, c, n) for i, (c, n) in enumerate(zip(cams, names))], f)
    return path


def write_rgb_cameras_npz(root: str, cams) -> str:
    """ahq2gaustar.py:12-47: OpenCV world-to-camera [R|t] (3x4), pinhole K with the principal point in pixels,
    zero distortion, shape = (height, width)."""
    n = len(cams)
    intr, extr = np.zeros((n, 3, 3)), np.zeros((n, 3, 4))
    shape = np.zeros((n, 2), dtype=np.int32)
    for i, c in enumerate(cams):
        w2v = np.asarray(c.viewmatrix, np.float64).T
        extr[i] = w2v[:3, :4]
        intr[i] = [[c.image_width / (2.0 * c.tanfovx), 0.0, 0.5 * c.image_width], [0.0, c.image_height / (2.0 * c.tanfovy), 0.5 * c.image_height],
                   [0.0, 0.0, 1.0]]
        shape[i] = (c.image_height, c.image_width)
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, "rgb_cameras.npz")
    np.savez_compressed(path, ids=np.arange(n), intrinsics=intr, extrinsics=extr, dist_coeffs=np.zeros((n, 5)), shape=shape)
    return path


def ply_attribute_names(n_sh: int):
    """gaussian_model.py:177-189."""
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(3 * (n_sh - 1))]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def write_point_cloud_ply(gs_out: str, g: scene.Gaussians, iteration: int = 1) -> str:
    """Binary little-endian PLY with the attribute list of GaussianModel.save_ply (gaussian_model.py:191-213).  Stored values
    are PRE-activation, as the model keeps them: opacity = logit, scale = log, rotation unnormalised (here: the unit
    quaternion); features are laid out channel-major: f_rest_{c*K + k} = shs[:, 1 + k, c] (save_ply transposes
    [P, K, 3] -> [P, 3, K] before flattening; load_ply reshapes (P, 3, K), :235)."""
    P, M = g.shs.shape[0], g.shs.shape[1]
    op = np.clip(g.opacities.reshape(P, 1).astype(np.float64), 1e-6, 1.0 - 1e-6)
    cols = [g.means3D.astype(np.float32), np.zeros((P, 3), np.float32), g.shs[:, 0, :].astype(np.float32),
            np.ascontiguousarray(g.shs[:, 1:, :].transpose(0, 2, 1)).reshape(P, 3 * (M - 1)).astype(np.float32),
            np.log(op / (1.0 - op)).astype(np.float32), np.log(g.scales.astype(np.float64)).astype(np.float32), g.rotations.astype(np.float32)]
    table = np.ascontiguousarray(np.concatenate(cols, axis=1), dtype="<f4")
    names = ply_attribute_names(M)
    assert table.shape[1] == len(names)
    d = os.path.join(gs_out, "point_cloud", f"iteration_{iteration}")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "point_cloud.ply")
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {P}\n" + "".join(f"property float {n}\n" for n in names) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(table.tobytes())
    return path


def write_obj(path: str, verts: np.ndarray, faces: np.ndarray) -> str:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        for v in np.asarray(verts, np.float64):
            f.write(f"v {v[0]:.8f} {v[1]:.8f} {v[2]:.8f}\n")
        for t in np.asarray(faces, np.int64) + 1:
            f.write(f"f {t[0]} {t[1]} {t[2]}\n")
    return path


def write_frame(root: str, frame: int, images, alphas, depths, flows_fwd=None, flows_bwd=None, label: str = "_humanrf") -> str:
    """One frame directory.  images: [V,H,W,3] uint8; alphas: [V,H,W] uint8 (0..255); depths: [V,H,W] float32 (metres along the
    camera z axis, what the depth pass of refine.py:600-616 renders); flows_*: [V,H,W,2] float32 as (dx, dy) in pixels -- written
    as (dy, dx) because warp_mesh.py:270-271 flips the last axis when loading."""
    from PIL import Image
    fd = os.path.join(root, f"{frame:04d}")
    for sub in ("images", f"masks{label}", f"depth{label}") + (("flow_bi",) if flows_fwd is not None else ()):
        os.makedirs(os.path.join(fd, sub), exist_ok=True)
    for v in range(len(images)):
        Image.fromarray(np.asarray(images[v], np.uint8)).save(os.path.join(fd, "images", f"img_{v:04d}.jpg"), quality=95, subsampling=0)
        Image.fromarray(np.asarray(alphas[v], np.uint8)).save(os.path.join(fd, f"masks{label}", f"img_{v:04d}_alpha.png"))
        np.savez_compressed(os.path.join(fd, f"depth{label}", f"img_{v:04d}_depth.npz"), depth=np.asarray(depths[v], np.float32))
        if flows_fwd is not None:
            np.savez_compressed(os.path.join(fd, "flow_bi", f"{v:04d}_f.npz"), flow=np.asarray(flows_fwd[v], np.float32)[..., ::-1])
            np.savez_compressed(os.path.join(fd, "flow_bi", f"{v:04d}_b.npz"), flow=np.asarray(flows_bwd[v], np.float32)[..., ::-1])
    return fd
