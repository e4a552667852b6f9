"""Developer tool (GPU box): compare gaustar_b200 vs the unmodified reference vs the CPU oracle.

    gpurun -- python tools/parity_report.py [--quick]

Prints, per scene, bit-exact mismatch counts for the integer / key quantities and max errors for the
floating-point ones.  The pytest parity tests (tests/test_parity_gpu.py) assert the same things.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import capi, scene  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import refgpu  # noqa: E402


def make_inputs(g, cam, use_sh=True, dev="cuda"):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kw = dict(means3D=t(g.means3D), opacities=t(g.opacities), viewmatrix=t(cam.viewmatrix), projmatrix=t(cam.projmatrix), campos=t(cam.campos),
              bg=torch.tensor([0.0, 1.0, 0.0], device=dev), tan_fovx=cam.tanfovx, tan_fovy=cam.tanfovy, W=cam.image_width, H=cam.image_height,
              scales=t(g.scales), rotations=t(g.rotations))
    if use_sh:
        kw.update(shs=t(g.shs), sh_degree=int(round(g.shs.shape[1] ** 0.5)) - 1)
    else:
        kw.update(colors_precomp=torch.rand(g.P, 3, device=dev, generator=torch.Generator(dev).manual_seed(3)))
    return kw


def bits(x):
    return x.contiguous().view(torch.int32)


def nmis(a, b):
    return int((a != b).sum().item())


def relerr(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max().item()), float(((a - b).abs().max() / (b.abs().max() + 1e-30)).item())


def compare(name, kw, with_oracle=True):
    P = kw["means3D"].shape[0]
    W, H = kw["W"], kw["H"]
    mine = capi.forward(**kw)
    torch.cuda.synchronize()
    mg = capi.unpack_geometry(mine, P)
    mi = capi.image_state(mine, W, H)
    mpl = capi.point_list(mine)
    ref = refgpu.forward(**kw)
    R = ref["num_rendered"]
    print(f"== {name}: P={P} {W}x{H} R_ref={R} R_mine={mine['num_rendered']} visible={int((ref['radii'] > 0).sum())} "
          f"max_tile={int((ref['ranges'][:, 1] - ref['ranges'][:, 0]).max())}")
    vis = ref["radii"] > 0
    print("   radii mismatches       ", nmis(mine["radii"], ref["radii"]))
    print("   tiles_touched mismatch ", nmis(mg["tiles_touched"], ref["tiles_touched"]))
    print("   depth bits mismatch    ", nmis(bits(mg["depths"])[vis], bits(ref["depths"])[vis]))
    print("   means2D bits mismatch  ", nmis(bits(mg["means2D"])[vis], bits(ref["means2D"])[vis]))
    print("   conic_op bits mismatch ", nmis(bits(mg["conic_opacity"])[vis], bits(ref["conic_opacity"])[vis]))
    print("   rgb max abs err        ", relerr(mg["rgb"][vis], ref["rgb"][vis])[0] if "shs" in kw else "n/a (precomp)")
    if mine["num_rendered"] == R:
        print("   point_list mismatches  ", nmis(mpl, ref["point_list"]))
        print("   ranges mismatches      ", nmis(mi["ranges"], ref["ranges"]))
    print("   n_contrib mismatches   ", nmis(mi["n_contrib"], ref["n_contrib"]), "of", W * H)
    print("   final_T max abs err    ", relerr(mi["final_T"], ref["final_T"])[0])
    print("   out_color max abs err  ", relerr(mine["out_color"], ref["out_color"])[0])
    gen = torch.Generator("cuda").manual_seed(1)
    dpix = torch.randn(3, H, W, device="cuda", generator=gen) / (W * H)
    bkw = {k: v for k, v in kw.items() if k not in ("opacities", "W", "H")}
    gm = capi.backward(mine, dpix, **bkw)
    torch.cuda.synchronize()
    gr = refgpu.backward(ref, dpix, **bkw)
    for k in ("dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
        if gm[k].numel() == 0:
            continue
        a, r = relerr(gm[k], gr[k])
        print(f"   {k:14s} max abs err {a:.3e}  (rel to max |ref| {r:.3e})")
    if with_oracle:
        inp = O.Inputs(means3D=kw["means3D"].cpu().numpy(), opacities=kw["opacities"].cpu().numpy(), viewmatrix=kw["viewmatrix"].cpu().numpy(),
                       projmatrix=kw["projmatrix"].cpu().numpy(), campos=kw["campos"].cpu().numpy(), bg=kw["bg"].cpu().numpy(),
                       tan_fovx=kw["tan_fovx"], tan_fovy=kw["tan_fovy"], W=W, H=H,
                       shs=kw["shs"].cpu().numpy() if "shs" in kw else None,
                       colors_precomp=kw["colors_precomp"].cpu().numpy() if "colors_precomp" in kw else None,
                       scales=kw["scales"].cpu().numpy(), rotations=kw["rotations"].cpu().numpy(), sh_degree=kw.get("sh_degree", 0))
        t0 = time.time()
        of = O.forward(inp)
        ob = O.backward(inp, of, dpix.cpu().numpy())
        print(f"   [oracle {time.time() - t0:.1f}s] R={of.num_rendered}")
        tt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        print("   oracle vs ref: radii", nmis(tt(of.radii), ref["radii"]), "tiles", nmis(tt(of.tiles_touched.astype(np.int32)), ref["tiles_touched"]),
              "depth", nmis(bits(tt(of.depths))[vis], bits(ref["depths"])[vis]), "means2D", nmis(bits(tt(of.means2D))[vis], bits(ref["means2D"])[vis]),
              "conic", nmis(bits(tt(of.conic_opacity))[vis], bits(ref["conic_opacity"])[vis]),
              "cov3D", nmis(bits(tt(of.cov3D))[vis], bits(ref["cov3D"])[vis]))
        if of.num_rendered == R:
            print("   oracle vs ref: keys_unsorted", nmis(tt(of.keys_unsorted.view(np.int64)), ref["keys_unsorted"]),
                  "keys_sorted", nmis(tt(of.keys_sorted.view(np.int64)), ref["keys_sorted"]),
                  "point_list", nmis(tt(of.point_list.astype(np.int32)), ref["point_list"]),
                  "ranges", nmis(tt(of.ranges.astype(np.int32)), ref["ranges"]))
        print("   oracle vs ref: n_contrib", nmis(tt(of.n_contrib.astype(np.int32)), ref["n_contrib"]),
              "out_color err", relerr(tt(of.out_color), ref["out_color"])[0])
        for k in ("dL_dmeans2D", "dL_dconic", "dL_dopacity", "dL_dcolors", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
            v = getattr(ob, k)
            if v.size == 0:
                continue
            a, r = relerr(gm[k], tt(v).view_as(gm[k]))
            a2, r2 = relerr(gr[k], tt(v).view_as(gr[k]))
            print(f"   {k:14s} mine-vs-oracle rel {r:.3e} | ref-vs-oracle rel {r2:.3e}")
    return mine, ref


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def speed(P, W, H, use_sh=True, random=False):
    g = scene.random_gaussians(P, 3, seed=1, scale_range=(0.003, 0.06)) if random else scene.surface_gaussians(P, sh_degree=3)
    cams = scene.dome_cameras(8, W, H)
    kws = [make_inputs(g, c, use_sh) for c in cams]
    dpix = torch.randn(3, H, W, device="cuda") / (W * H)
    state = {"i": 0}

    def mine():
        kw = kws[state["i"] % len(kws)]; state["i"] += 1
        f = capi.forward(**kw)
        capi.backward(f, dpix, **{k: v for k, v in kw.items() if k not in ("opacities", "W", "H")})

    def ref():
        kw = kws[state["i"] % len(kws)]; state["i"] += 1
        f = refgpu.forward(intermediates=False, **kw)
        refgpu.backward(f, dpix, **{k: v for k, v in kw.items() if k not in ("opacities", "W", "H")})

    tm = timeit(mine)
    tr = timeit(ref, n=8, warm=2)
    print(f"== speed {'random-cloud' if random else 'surface'} P={g.P} {W}x{H} sh={use_sh}: mine {tm:.3f} ms/view ({1000 / tm:.0f} views/s) | ref(shim, sync'd) {tr:.3f} ms/view "
          f"({1000 / tr:.0f} views/s) | x{tr / tm:.2f}")
    # per-stage times of mine
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record(); e.record(); torch.cuda.synchronize()
    for st, nm in enumerate(capi.STAGES):
        capi.profile_stage(st, s, e)
        ts = []
        for _ in range(5):
            mine(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        print(f"   stage {nm:15s} {np.median(ts) * 1000:8.1f} us")
    capi.profile_stage(-1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-speed", action="store_true")
    a = ap.parse_args()
    torch.manual_seed(0)
    g = scene.surface_gaussians(30000, sh_degree=3)
    cam = scene.dome_cameras(4, 480, 270)[1]
    compare("surface-30k sh3", make_inputs(g, cam, True))
    compare("surface-30k precomp", make_inputs(g, cam, False), with_oracle=False)
    g2 = scene.random_gaussians(20000, sh_degree=2, seed=5)
    cam2 = scene.look_at_camera([0.5, 1.3, 4.0], [0, 1, 0], 640, 360, fy_over_H=1.2)
    compare("random-20k big splats sh2", make_inputs(g2, cam2, True))
    cam3 = scene.look_at_camera([0.2, 1.0, 0.9], [0, 1, 0], 333, 201, fy_over_H=0.9)
    compare("random-20k close-up odd size", make_inputs(g2, cam3, True), with_oracle=False)
    if not a.quick:
        g3 = scene.surface_gaussians(1000000, sh_degree=3)
        cam4 = scene.dome_cameras(8, 1920, 1080)[3]
        compare("surface-1M 1080p", make_inputs(g3, cam4, True), with_oracle=False)
    if not a.no_speed:
        speed(1000000, 1920, 1080, True)
        speed(1000000, 1920, 1080, False)
        speed(200000, 1920, 1080, True)
        speed(500000, 1920, 1080, True, random=True)
