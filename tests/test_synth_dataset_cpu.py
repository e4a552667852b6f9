"""The synthetic on-disk dataset writers (gaustar_b200/synth_dataset.py) against restatements of the
reference's READERS: every file is read back here the way GauSTAR's loaders read it (file:line cited per reader) and must
reproduce the scene it was written from.  CPU only."""
import json
import math
import os

import numpy as np

from gaustar_b200 import scene, synth_dataset as SD


def _read_gs_cameras(gs_out):
    """gaustar_scene/cameras.py:35-73 (load_gs_cameras): sorted by img_name; 'rotation'/'position' form a camera-to-world
    matrix that is inverted; R = Rt[:3,:3].T, T = Rt[:3,3]; fov = focal2fov(f, pixels) = 2 atan(pixels / (2 f))."""
    with open(os.path.join(gs_out, "cameras.json")) as f:
        entries = sorted(json.load(f), key=lambda x: x["img_name"])
    out = []
    for e in entries:
        c2w = np.zeros((4, 4))
        c2w[:3, :3] = np.array(e["rotation"])
        c2w[:3, 3] = np.array(e["position"])
        c2w[3, 3] = 1
        Rt = np.linalg.inv(c2w)
        out.append(dict(R=Rt[:3, :3].transpose(), T=Rt[:3, 3], width=e["width"], height=e["height"], name=e["img_name"], id=e["id"],
                        fov_x=2 * math.atan(e["width"] / (2 * e["fx"])), fov_y=2 * math.atan(e["height"] / (2 * e["fy"]))))
    return out


def _world2view(R, t):
    """gaussian_splatting/utils/graphics_utils.py:38-44 getWorld2View."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    return Rt


def _read_ply(path, max_sh_degree):
    """gaussian_splatting/scene/gaussian_model.py:215-256 (load_ply) on a minimal binary-little-endian PLY parser (plyfile is
    not in this image): columns are looked up BY NAME, f_rest_* and scale_*/rot_* sorted by their numeric suffix."""
    with open(path, "rb") as f:
        assert f.readline().strip() == b"ply" and f.readline().strip() == b"format binary_little_endian 1.0"
        names, n = [], None
        while True:
            line = f.readline().decode("ascii").strip()
            if line == "end_header":
                break
            tok = line.split()
            if tok[:2] == ["element", "vertex"]:
                n = int(tok[2])
            elif tok[0] == "property":
                assert tok[1] == "float"
                names.append(tok[2])
        tab = np.frombuffer(f.read(), dtype="<f4").reshape(n, len(names))
    col = {k: tab[:, i] for i, k in enumerate(names)}
    xyz = np.stack((col["x"], col["y"], col["z"]), axis=1)
    opacities = col["opacity"][..., None]
    dc = np.stack([col[f"f_dc_{i}"] for i in range(3)], axis=1)[:, :, None]          # (P, 3, 1)
    extra = sorted([k for k in names if k.startswith("f_rest_")], key=lambda x: int(x.split("_")[-1]))
    assert len(extra) == 3 * (max_sh_degree + 1) ** 2 - 3
    rest = np.stack([col[k] for k in extra], axis=1).reshape(n, 3, (max_sh_degree + 1) ** 2 - 1)
    scales = np.stack([col[k] for k in sorted([k for k in names if k.startswith("scale_")], key=lambda x: int(x.split("_")[-1]))], axis=1)
    rots = np.stack([col[k] for k in sorted([k for k in names if k.startswith("rot")], key=lambda x: int(x.split("_")[-1]))], axis=1)
    features = np.concatenate([dc.transpose(0, 2, 1), rest.transpose(0, 2, 1)], axis=1)  # get_features: cat(dc, rest) of the [P,K,3] params
    return xyz, opacities, features, scales, rots


def test_cameras_json_round_trip(tmp_path):
    cams = scene.dome_cameras(7, 320, 200)
    names = [f"img_{i:04d}" for i in range(len(cams))]
    SD.write_cameras_json(str(tmp_path), cams[::-1], names[::-1])  # written unsorted: the loader sorts by name
    got = _read_gs_cameras(str(tmp_path))
    assert [g["name"] for g in got] == names
    for g, c in zip(got, cams):
        np.testing.assert_allclose(_world2view(g["R"], g["T"]).T, c.viewmatrix, atol=1e-6)
        assert g["width"] == c.image_width and g["height"] == c.image_height
        np.testing.assert_allclose(math.tan(g["fov_x"] / 2), c.tanfovx, rtol=1e-6)
        np.testing.assert_allclose(math.tan(g["fov_y"] / 2), c.tanfovy, rtol=1e-6)
        # the camera centre the rasterizer is given (sugar_model.py:1149-1163) is the 'position' the file stores
        np.testing.assert_allclose(np.linalg.inv(_world2view(g["R"], g["T"]))[:3, 3], c.campos, atol=1e-5)


def test_point_cloud_ply_round_trip(tmp_path):
    g = scene.surface_gaussians(600, sh_degree=2, seed=3)
    path = SD.write_point_cloud_ply(str(tmp_path), g, iteration=1)
    assert path.endswith(os.path.join("point_cloud", "iteration_1", "point_cloud.ply"))
    xyz, op, feats, scales, rots = _read_ply(path, max_sh_degree=2)
    np.testing.assert_array_equal(xyz, g.means3D.astype(np.float32))
    np.testing.assert_array_equal(feats, g.shs.astype(np.float32))
    # the activations GaussianModel applies (gaussian_model.py:26-43): sigmoid, exp, normalize
    np.testing.assert_allclose(1.0 / (1.0 + np.exp(-op.astype(np.float64))), g.opacities.reshape(-1, 1), rtol=1e-5)
    np.testing.assert_allclose(np.exp(scales.astype(np.float64)), g.scales, rtol=1e-5)
    np.testing.assert_allclose(rots / np.linalg.norm(rots, axis=1, keepdims=True), g.rotations, atol=1e-6)


def test_rgb_cameras_npz_projects_like_the_rasterizer_camera(tmp_path):
    cams = scene.dome_cameras(5, 400, 240)
    info = dict(np.load(SD.write_rgb_cameras_npz(str(tmp_path), cams)))
    assert info["intrinsics"].shape == (5, 3, 3) and info["extrinsics"].shape == (5, 3, 4) and info["shape"].tolist() == [[240, 400]] * 5
    pts = np.random.default_rng(0).normal([0, 1, 0], 0.3, (50, 3))
    for i, c in enumerate(cams):
        # OpenCV pinhole projection (ahq2gaustar.py:12-47 conventions) ...
        pc = pts @ info["extrinsics"][i][:, :3].T + info["extrinsics"][i][:, 3]
        uv = (pc @ info["intrinsics"][i].T)
        uv = uv[:, :2] / uv[:, 2:3]
        # ... equals the rasterizer's pixel coordinates: ndc2Pix of the projected point, plus the half-pixel centre
        # (auxiliary.h:41-44: pix = ((ndc + 1) * S - 1) / 2 is the index of the pixel whose CENTRE is hit)
        ph = np.c_[pts, np.ones(len(pts))] @ c.projmatrix
        ndc = ph[:, :2] / ph[:, 3:4]
        pix = ((ndc + 1.0) * np.array([c.image_width, c.image_height]) - 1.0) * 0.5
        np.testing.assert_allclose(uv - 0.5, pix, atol=1e-3)


def test_frame_files_read_back_like_the_loaders(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(1)
    V, H, W = 2, 24, 32
    yy, xx = np.mgrid[0:H, 0:W]
    images = np.stack([np.stack([xx * 7 + 10 * v, yy * 9, 255 - xx * 5 - yy * 3], axis=-1) for v in range(V)]).astype(np.uint8)  # smooth: survives JPEG
    alphas = (rng.random((V, H, W)) > 0.5).astype(np.uint8) * 255
    depths = rng.uniform(1.0, 4.0, (V, H, W)).astype(np.float32)
    ff, fb = rng.normal(0, 2, (V, H, W, 2)).astype(np.float32), rng.normal(0, 2, (V, H, W, 2)).astype(np.float32)
    root = str(tmp_path) + "/"
    fd = SD.write_frame(root, 3, images, alphas, depths, ff, fb)
    assert fd.endswith("0003")
    source_path = root + "0003/"
    for v in range(V):
        name = f"img_{v:04d}"
        # cameras.py:33,40,76-78: extension from the first file of images/, image = PIL open
        ext = "." + os.listdir(os.path.join(source_path, "images"))[0].split(".")[-1]
        assert ext == ".jpg"
        img = np.asarray(Image.open(os.path.join(source_path, "images", name + ext)))
        assert img.shape == (H, W, 3) and np.abs(img.astype(int) - images[v].astype(int)).mean() < 3
        # cameras.py:97-98 mask, :102-106 depth
        np.testing.assert_array_equal(np.asarray(Image.open(source_path + f"masks_humanrf/{name}_alpha.png")), alphas[v])
        np.testing.assert_array_equal(np.load(source_path + f"depth_humanrf/{name}_depth.npz")["depth"], depths[v])
        # warp_mesh.py:269-271: the loader flips the last axis and then treats it as (dx, dy)
        np.testing.assert_array_equal(np.load(source_path + f"flow_bi/{v:04d}_f.npz")["flow"][..., ::-1], ff[v])
        np.testing.assert_array_equal(np.load(source_path + f"flow_bi/{v:04d}_b.npz")["flow"][..., ::-1], fb[v])


def test_obj_round_trip(tmp_path):
    verts, faces = scene.capsule_mesh(200, seed=1)[:2]
    path = SD.write_obj(str(tmp_path / "init_mesh.obj"), verts, faces)
    v, f = [], []
    for line in open(path):
        t = line.split()
        if t[0] == "v":
            v.append([float(x) for x in t[1:]])
        elif t[0] == "f":
            f.append([int(x) - 1 for x in t[1:]])
    np.testing.assert_allclose(np.array(v), verts, atol=1e-7)
    np.testing.assert_array_equal(np.array(f), faces)
