"""Shared-geometry re-blend (gstar_raster_reblend) -- GPU tests, run last.

GauSTAR rasterizes the same Gaussians from the same camera twice per step: RGB, then depth as three equal colour
channels with another background (gaustar_trainers/refine.py:552-564 and :607-616).  The second call may start from the
first call's sorted record stream.  Bar: the re-blend's image / final_T / n_contrib are BIT-IDENTICAL to a full forward
on the same inputs (same records, same order, same blend kernel); its gradients agree with the full call's within the
run-to-run spread of fp32 atomics (the same bounds test_parity_gpu.py uses against the live reference).
"""
import contextlib

import numpy as np
import pytest
import torch

from gaustar_b200 import capi

import helpers as Hh
from oracle import oracle as O
from test_parity_gpu import (GRAD_TOL, LIVE_REF_TOL, SCENES, backward_path, check_forward_against, check_grads,  # noqa: F401
                             run_mine)  # backward_path: autouse fixture

pytestmark = pytest.mark.gpu


def depth_colors(kw):
    """View-space depth of every Gaussian as three equal channels (refine.py:600-605)."""
    vm = kw["viewmatrix"].view(4, 4)
    z = kw["means3D"] @ vm[:3, 2] + vm[3, 2]
    return z[:, None].expand(-1, 3).contiguous()


def second_pass_kwargs(kw, col, bg):
    kw2 = {k: v for k, v in kw.items() if k != "shs"}
    kw2.update(colors_precomp=col, bg=bg, sh_degree=0)
    return kw2


@pytest.mark.parametrize("name", ["surface_sh3", "surface_precomp", "random_big_sh2", "random_closeup_odd"])
def test_reblend_equals_full_forward_through_the_c_abi(name, backward_path):
    d = SCENES[name]()
    kw, first = run_mine(d)
    W, H = kw["W"], kw["H"]
    col = depth_colors(kw)
    bg2 = torch.full((3,), 7.0, device="cuda")
    rb = capi.reblend(first, col, bg2, W, H)
    kw2 = second_pass_kwargs(kw, col, bg2)
    full = capi.forward(**kw2)
    torch.cuda.synchronize()
    assert rb["num_rendered"] == full["num_rendered"] == first["num_rendered"]
    assert torch.equal(rb["out_color"], full["out_color"])
    sr, sf = capi.image_state(rb, W, H), capi.image_state(full, W, H)
    for k in ("final_T", "n_contrib", "ranges"):
        assert torch.equal(sr[k], sf[k]), k
    assert capi.hit_log_state(rb)[2] == (backward_path == "hitlog")
    # the first pass is untouched by the second
    again = capi.forward(**kw)
    assert torch.equal(again["out_color"], first["out_color"])
    # backward of the re-blend: the source call's geometry buffer + its own binning / image buffers
    dpix = torch.randn(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(11))
    g_rb = capi.backward(rb, dpix, **Hh.bwd_kwargs(kw2))
    g_full = capi.backward(full, dpix, **Hh.bwd_kwargs(kw2))
    torch.cuda.synchronize()
    check_grads(g_rb, {k: g_full[k].cpu().numpy() for k in Hh.GRAD_KEYS}, per_key=LIVE_REF_TOL)
    # a re-blend can be re-blended (three passes in refined_mesh.py:733-774)
    col3 = torch.rand_like(col)
    rb3 = capi.reblend(rb, col3, kw["bg"], W, H, forward_only=True)
    full3 = capi.forward(forward_only=True, **second_pass_kwargs(kw, col3, kw["bg"]))
    torch.cuda.synchronize()
    assert torch.equal(rb3["out_color"], full3["out_color"])
    assert not capi.hit_log_state(rb3)[2]


def test_reblend_against_the_cpu_oracle():
    """Independent check: the re-blended second pass (first pass: SH colours) against the CPU oracle run on
    (same geometry, colors_precomp, other background) -- intermediates, sorted list, image and gradients."""
    d = SCENES["surface_sh3"]()
    kw, first = run_mine(d)
    W, H, P = kw["W"], kw["H"], kw["means3D"].shape[0]
    rng = np.random.default_rng(9)
    col = rng.uniform(0, 1, (P, 3)).astype(np.float32)
    d2 = {k: v for k, v in d.items() if k not in ("shs", "sh_degree")}
    d2.update(colors_precomp=col, bg=np.array([10.0, 10.0, 10.0], np.float32), sh_degree=0)  # refine.py:609: the depth pass's background
    kw2 = Hh.to_torch_kwargs(d2)
    rb = capi.reblend(first, kw2["colors_precomp"], kw2["bg"], W, H)
    torch.cuda.synchronize()
    inp = Hh.oracle_inputs_from_dict(d2)
    of = O.forward(inp)
    _, st = check_forward_against(rb, kw2, dict(of.__dict__), exact_image=False, n_contrib_slack=max(2, W * H // 20000))
    dpix = rng.normal(0, 1, (3, H, W)).astype(np.float32)
    mine = capi.backward(rb, torch.from_numpy(dpix).cuda(), **Hh.bwd_kwargs(kw2))
    torch.cuda.synchronize()
    of.n_contrib = st["n_contrib"].astype(np.uint32)
    of.final_T = st["final_T"].copy()
    check_grads(mine, O.backward(inp, of, dpix).__dict__)


def test_reblend_rejects_unknown_source_and_size_mismatch():
    d = SCENES["sh_deg1_of_3"]()
    kw, first = run_mine(d)
    col = depth_colors(kw)
    stale = dict(first)
    # same bytes at an address no call of this thread ever used (256 bytes into a fresh allocation: torch's allocations
    # start on 512-byte boundaries, so this cannot be a recycled image buffer either)
    stale["image"] = torch.cat([torch.zeros(256, dtype=torch.uint8, device="cuda"), first["image"]])[256:]
    with pytest.raises(capi.GstarError):
        capi.reblend(stale, col, kw["bg"], kw["W"], kw["H"])
    with pytest.raises(capi.GstarError):
        capi.reblend(first, col, kw["bg"], kw["W"] + 16, kw["H"])
    with pytest.raises(capi.GstarError):
        capi.reblend(first, col[:-1], kw["bg"], kw["W"], kw["H"])
    ok = capi.reblend(first, col, kw["bg"], kw["W"], kw["H"])  # and the source is still usable afterwards
    torch.cuda.synchronize()
    assert torch.isfinite(ok["out_color"]).all()


def test_reblend_through_another_camera_is_refused_on_the_device():
    """The optional camera guard: equal matrices in other memory pass; another camera gives a NaN image and overflow word 2
    (never a plausible picture of the wrong view) -- through the C ABI and for a shared_geometry() block that was
    wrongly put around two cameras."""
    import diff_gaussian_rasterization as dgr
    from gaustar_b200 import scene
    g = scene.surface_gaussians(12000, 3, seed=2)
    cams = scene.dome_cameras(6, 320, 200)
    kwA, first = run_mine(Hh.scene_dict(g, cams[1]))
    kwB = Hh.to_torch_kwargs(Hh.scene_dict(g, cams[3]))
    W, H = kwA["W"], kwA["H"]
    col = depth_colors(kwA)
    plain = capi.reblend(first, col, kwA["bg"], W, H)
    same = capi.reblend(first, col, kwA["bg"], W, H, viewmatrix=kwA["viewmatrix"].clone(), projmatrix=kwA["projmatrix"].clone())
    other = capi.reblend(first, col, kwA["bg"], W, H, viewmatrix=kwB["viewmatrix"], projmatrix=kwB["projmatrix"])
    torch.cuda.synchronize()
    assert torch.equal(same["out_color"], plain["out_color"]) and capi.debug_header(same)["overflow"] == 0
    assert torch.isnan(other["out_color"]).all() and capi.debug_header(other)["overflow"] == 2
    chained = capi.reblend(other, col, kwA["bg"], W, H)  # a refused call stays refused down the chain
    torch.cuda.synchronize()
    assert torch.isnan(chained["out_color"]).all()
    # backward of a refused call: no work, zero gradients, nothing out of bounds
    gr = capi.backward(other, torch.ones(3, H, W, device="cuda"), **Hh.bwd_kwargs(second_pass_kwargs(kwA, col, kwA["bg"])))
    torch.cuda.synchronize()
    assert float(gr["dL_dmeans3D"].abs().max()) == 0.0 and float(gr["dL_dcolors"].abs().max()) == 0.0

    def call(kw):
        return dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 0))(means3D=kw["means3D"], means2D=torch.zeros_like(kw["means3D"]),
                                                                      opacities=kw["opacities"], colors_precomp=col, scales=kw["scales"],
                                                                      rotations=kw["rotations"])[0]
    with torch.no_grad():
        refA, refB = call(kwA), call(kwB)
        with dgr.shared_geometry():
            a1, a2 = call(kwA), call(kwA)  # full forward, then a legitimate re-blend
            b = call(kwB)                  # the block wrongly spans a second camera
    torch.cuda.synchronize()
    assert torch.equal(a1, refA) and torch.equal(a2, refA) and not torch.equal(refA, refB)
    assert torch.isnan(b).all()


def test_reblend_of_a_view_without_instances_is_the_background():
    dev = "cuda"
    P, W, H = 64, 56, 40
    m = torch.randn(P, 3, device=dev)
    m[:, 2] = -5.0  # behind the camera
    rot = torch.zeros(P, 4, device=dev); rot[:, 0] = 1
    eye = torch.eye(4, device=dev).reshape(-1)
    first = capi.forward(m, torch.full((P, 1), 0.5, device=dev), eye, eye, torch.zeros(3, device=dev), torch.zeros(3, device=dev), 0.5, 0.5, W, H,
                         colors_precomp=torch.rand(P, 3, device=dev), scales=torch.full((P, 3), 0.1, device=dev), rotations=rot)
    assert first["num_rendered"] == 0
    bg2 = torch.tensor([0.25, 0.5, 0.75], device=dev)
    rb = capi.reblend(first, torch.rand(P, 3, device=dev), bg2, W, H)
    torch.cuda.synchronize()
    assert rb["num_rendered"] == 0
    assert torch.equal(rb["out_color"], bg2.view(3, 1, 1).expand(3, H, W))


def _settings(dgr, kw, bg, sh_degree):
    # fresh camera tensors per call, like sugar_model.py:1149-1163 (numpy inverse + upload every time)
    return dgr.GaussianRasterizationSettings(image_height=kw["H"], image_width=kw["W"], tanfovx=kw["tan_fovx"], tanfovy=kw["tan_fovy"], bg=bg,
                                             scale_modifier=1.0, viewmatrix=kw["viewmatrix"].view(4, 4).clone(),
                                             projmatrix=kw["projmatrix"].view(4, 4).clone(), sh_degree=sh_degree, campos=kw["campos"].view(1, 3),
                                             prefiltered=False, debug=False)


def test_operator_shared_geometry_block_matches_two_full_calls():
    """The training-step pattern of refine.py: RGB through SH, then depth through colors_precomp with another background,
    every per-Gaussian tensor REBUILT between the calls (equal values, new tensors).  Inside shared_geometry() the second
    call re-blends; images and radii are identical and the parameter gradients of the step agree."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    names = ("means3D", "opacities", "scales", "rotations")
    params = {k: kw[k].clone().requires_grad_(True) for k in names}
    shs = kw["shs"].clone().requires_grad_(True)
    vm = kw["viewmatrix"].view(4, 4)
    gen = torch.Generator("cuda").manual_seed(5)
    w1 = torch.randn(3, kw["H"], kw["W"], device="cuda", generator=gen)
    w2 = torch.randn(3, kw["H"], kw["W"], device="cuda", generator=gen)

    def step(shared):
        for p in list(params.values()) + [shs]:
            p.grad = None
        with (dgr.shared_geometry(check=True) if shared else contextlib.nullcontext()):
            a = {k: v * 1.0 for k, v in params.items()}
            img, radii = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 3))(
                means3D=a["means3D"], means2D=torch.zeros_like(a["means3D"], requires_grad=True), opacities=a["opacities"], shs=shs,
                scales=a["scales"], rotations=a["rotations"])
            b = {k: v * 1.0 for k, v in params.items()}
            depth = (b["means3D"] @ vm[:3, 2] + vm[3, 2])[:, None].expand(-1, 3)
            dimg, radii2 = dgr.GaussianRasterizer(_settings(dgr, kw, torch.full((3,), 9.0, device="cuda"), 0))(
                means3D=b["means3D"], means2D=torch.zeros_like(b["means3D"], requires_grad=True), opacities=b["opacities"],
                colors_precomp=depth, scales=b["scales"], rotations=b["rotations"])
        fn = type(dimg.grad_fn).__name__
        ((img * w1).sum() + (dimg * w2).sum()).backward()
        torch.cuda.synchronize()
        grads = {k: p.grad.clone() for k, p in params.items()}
        grads["shs"] = shs.grad.clone()
        return img.detach(), dimg.detach(), radii, radii2, grads, fn

    img0, dimg0, r0, r20, g0, fn0 = step(False)
    img1, dimg1, r1, r21, g1, fn1 = step(True)
    assert "Reblend" not in fn0 and "Reblend" in fn1, (fn0, fn1)
    assert torch.equal(img0, img1) and torch.equal(dimg0, dimg1)
    assert torch.equal(r0, r1) and torch.equal(r20, r21) and torch.equal(r1, r21)
    tol = {"scales": 3e-3, "rotations": 3e-3}
    for k in g0:
        assert torch.isfinite(g1[k]).all(), k
        assert Hh.rel_err(g1[k].cpu(), g0[k].cpu()) < max(GRAD_TOL, tol.get(k, 0.0)), (k, Hh.rel_err(g1[k].cpu(), g0[k].cpu()))
    # outside the block nothing is remembered: a colors_precomp call is a full forward again
    b = {k: v.detach() for k, v in params.items()}
    out, _ = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 0))(means3D=b["means3D"].requires_grad_(True), means2D=torch.zeros_like(b["means3D"]),
                                                                     opacities=b["opacities"], colors_precomp=torch.rand_like(b["means3D"]),
                                                                     scales=b["scales"], rotations=b["rotations"])
    assert "Reblend" not in type(out.grad_fn).__name__


def test_shared_geometry_check_catches_a_moved_gaussian():
    import diff_gaussian_rasterization as dgr
    d = SCENES["sh_deg1_of_3"]()
    kw = Hh.to_torch_kwargs(d)
    col = torch.rand_like(kw["means3D"])
    call = lambda m: dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 0))(means3D=m, means2D=torch.zeros_like(m), opacities=kw["opacities"],
                                                                             colors_precomp=col, scales=kw["scales"], rotations=kw["rotations"])
    with dgr.shared_geometry(check=True):
        call(kw["means3D"])
        moved = kw["means3D"].clone()
        moved[0, 0] += 1.0
        with pytest.raises(RuntimeError, match="means3D"):
            call(moved)


def test_identity_keyed_cache_reuses_only_unmodified_tensors():
    """set_geometry_cache(True): same tensor memory and version -> re-blend; an in-place update (an optimizer step) or any
    other tensor -> full forward.  Results equal the cache-off calls either way, also under no_grad (inference)."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_precomp"]()
    kw = Hh.to_torch_kwargs(d)
    m = kw["means3D"].clone().requires_grad_(True)
    rs = _settings(dgr, kw, kw["bg"], 0)
    R = dgr.GaussianRasterizer(rs)
    colA, colB = kw["colors_precomp"], depth_colors(kw)
    call = lambda col: R(means3D=m, means2D=torch.zeros_like(m), opacities=kw["opacities"], colors_precomp=col, scales=kw["scales"],
                         rotations=kw["rotations"])
    refA, _ = call(colA)
    refB, _ = call(colB)
    assert "Reblend" not in type(refB.grad_fn).__name__
    old = dgr.set_geometry_cache(True)
    try:
        a, _ = call(colA)
        b, _ = call(colB)
        assert "Reblend" not in type(a.grad_fn).__name__ and "Reblend" in type(b.grad_fn).__name__
        assert torch.equal(a, refA) and torch.equal(b, refB)
        with torch.no_grad():
            bi, radii_i = call(colB)
        assert torch.equal(bi, refB)
        with torch.no_grad():
            m.add_(0.01)  # what an optimizer step does: same memory, new version
        c, _ = call(colB)
        assert "Reblend" not in type(c.grad_fn).__name__
        assert not torch.equal(c, refB)
        c2, _ = call(colA)
        assert "Reblend" in type(c2.grad_fn).__name__
        (c2.sum() + c.sum()).backward()
        torch.cuda.synchronize()
        assert torch.isfinite(m.grad).all() and float(m.grad.abs().max()) > 0
    finally:
        dgr.set_geometry_cache(old)
    d2, _ = call(colA)
    assert "Reblend" not in type(d2.grad_fn).__name__


def test_forward_passes_one_node_matches_separate_calls():
    """GaussianRasterizer.forward_passes (config #5's RGB + depth + normal passes as ONE autograd node: preprocess/sort once,
    the per-Gaussian backward stage once for all passes) against three ordinary calls: identical images, and the step's
    gradients -- including the ones that reach means3D through the extra passes' colours, and means2D summed over the
    passes -- agree within the atomics' spread."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    P, H, W = kw["means3D"].shape[0], kw["H"], kw["W"]
    leaves = {k: kw[k].clone().requires_grad_(True) for k in ("means3D", "opacities", "scales", "rotations", "shs")}
    vm = kw["viewmatrix"].view(4, 4)
    gen = torch.Generator("cuda").manual_seed(21)
    ws = [torch.randn(3, H, W, device="cuda", generator=gen) for _ in range(3)]
    bg2, bg3 = torch.full((3,), 10.0, device="cuda"), torch.tensor([0.5, 0.5, 1.0], device="cuda")
    center = torch.tensor([0.0, 1.0, 0.0], device="cuda")
    free_col = torch.rand(P, 3, device="cuda", generator=gen)  # an extra pass whose colours need no gradient

    def run(fused, use=(True, True, True)):
        for t in leaves.values():
            t.grad = None
        m, o, sc, rot, shs = (leaves[k] for k in ("means3D", "opacities", "scales", "rotations", "shs"))
        m2 = torch.zeros(P, 3, device="cuda", requires_grad=True)
        depth = (m @ vm[:3, 2] + vm[3, 2])[:, None].expand(-1, 3)
        nrm = torch.nn.functional.normalize(m - center, dim=-1)
        r1 = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 3))
        if fused:
            img, radii, extra = r1.forward_passes(means3D=m, means2D=m2, opacities=o, shs=shs, scales=sc, rotations=rot,
                                                  extra_passes=[(depth, bg2), (nrm, bg3), (free_col, bg3)])
        else:
            img, radii = r1(means3D=m, means2D=m2, opacities=o, shs=shs, scales=sc, rotations=rot)
            extra = [dgr.GaussianRasterizer(_settings(dgr, kw, bg, 0))(means3D=m, means2D=m2, opacities=o, colors_precomp=c, scales=sc,
                                                                      rotations=rot)[0] for c, bg in ((depth, bg2), (nrm, bg3), (free_col, bg3))]
        loss = sum((x * w).sum() for x, w, u in zip([img] + extra[:2], ws, use) if u)
        loss.backward()
        torch.cuda.synchronize()
        grads = {k: (torch.zeros_like(t) if t.grad is None else t.grad.clone()) for k, t in leaves.items()}
        grads["means2D"] = m2.grad.clone()
        return [img.detach()] + [e.detach() for e in extra], radii, grads

    for use in ((True, True, True), (True, False, False), (False, True, False)):
        imgs0, radii0, g0 = run(False, use)
        imgs1, radii1, g1 = run(True, use)
        assert torch.equal(radii0, radii1)
        for a, b in zip(imgs0, imgs1):
            assert torch.equal(a, b)
        tol = {"scales": 3e-3, "rotations": 3e-3}
        for k in g0:
            assert torch.isfinite(g1[k]).all(), k
            assert Hh.rel_err(g1[k].cpu(), g0[k].cpu()) < max(GRAD_TOL, tol.get(k, 0.0)), (use, k, Hh.rel_err(g1[k].cpu(), g0[k].cpu()))
    # inference and the empty scene
    with torch.no_grad():
        img, radii, extra = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 3)).forward_passes(
            means3D=kw["means3D"], means2D=torch.zeros(P, 3, device="cuda"), opacities=kw["opacities"], shs=kw["shs"], scales=kw["scales"],
            rotations=kw["rotations"], extra_passes=[(free_col, bg3)])
    assert torch.equal(img, imgs0[0]) and torch.equal(extra[0], imgs0[3])
    z3, z4, z1 = torch.zeros(0, 3, device="cuda"), torch.zeros(0, 4, device="cuda"), torch.zeros(0, 1, device="cuda")
    img, radii, extra = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 0)).forward_passes(z3, z3, z1, colors_precomp=z3, scales=z3, rotations=z4,
                                                                                                 extra_passes=[(z3, bg2)])
    assert float(img.abs().max()) == 0.0 and float(extra[0].abs().max()) == 0.0 and radii.numel() == 0


def test_forward_passes_after_many_forwards_still_finds_its_source():
    """Regression (round 1, profiles/r1last_pytest_failure.txt): the host-side table of recent calls holds 16 entries; with FIFO
    replacement a multi-pass call issued after more than 16 other forwards could evict its own source while re-blending.
    More than 16 forwards, then one forward_passes call with three extra passes, then more forwards and another one."""
    import diff_gaussian_rasterization as dgr
    d = SCENES["surface_sh3"]()
    kw = Hh.to_torch_kwargs(d)
    P = kw["means3D"].shape[0]
    gen = torch.Generator("cuda").manual_seed(5)
    cols = [torch.rand(P, 3, device="cuda", generator=gen) for _ in range(3)]
    bg2 = torch.full((3,), 2.0, device="cuda")
    r = dgr.GaussianRasterizer(_settings(dgr, kw, kw["bg"], 3))
    args = dict(means3D=kw["means3D"], means2D=torch.zeros(P, 3, device="cuda"), opacities=kw["opacities"], shs=kw["shs"], scales=kw["scales"],
                rotations=kw["rotations"])
    keep = []
    for rounds in (19, 23):
        with torch.no_grad():
            for _ in range(rounds):
                keep.append(r(**args)[0])  # (kept alive: every call owns fresh buffers, the table sees distinct addresses)
            img, radii, extra = r.forward_passes(**args, extra_passes=[(c, bg2) for c in cols])
            ref = [dgr.GaussianRasterizer(_settings(dgr, kw, bg2, 0))(means3D=kw["means3D"], means2D=args["means2D"], opacities=kw["opacities"],
                                                                      colors_precomp=c, scales=kw["scales"], rotations=kw["rotations"])[0] for c in cols]
        torch.cuda.synchronize()
        assert torch.equal(img, keep[0])
        for e, x in zip(extra, ref):
            assert torch.equal(e, x)
        del keep[4:]


def test_camera_overwritten_in_place_between_the_calls_is_refused():
    """The device-side camera guard compares the re-blend's camera with a SNAPSHOT of the source call's matrices: overwriting the
    same tensors in place (a loop over cameras that reuses its buffers) must not pass as "same camera" -- the re-blend is refused
    (NaN image), and after release_shared_geometry() / a new block the next call is a full forward again."""
    import diff_gaussian_rasterization as dgr
    from gaustar_b200 import scene
    g = scene.surface_gaussians(8000, 3, seed=3)
    cams = scene.dome_cameras(6, 256, 144)
    kwA, kwB = Hh.to_torch_kwargs(Hh.scene_dict(g, cams[1])), Hh.to_torch_kwargs(Hh.scene_dict(g, cams[4]))
    vm, pm = kwA["viewmatrix"].view(4, 4).clone(), kwA["projmatrix"].view(4, 4).clone()
    P = kwA["means3D"].shape[0]
    col = torch.rand(P, 3, device="cuda")
    args = dict(means3D=kwA["means3D"], means2D=torch.zeros(P, 3, device="cuda"), opacities=kwA["opacities"], scales=kwA["scales"], rotations=kwA["rotations"])

    def rs(bg, deg):
        return dgr.GaussianRasterizationSettings(kwA["H"], kwA["W"], kwA["tan_fovx"], kwA["tan_fovy"], bg, 1.0, vm, pm, deg, kwA["campos"], False, False)

    with torch.no_grad():
        with dgr.shared_geometry():
            a, _ = dgr.GaussianRasterizer(rs(kwA["bg"], 3))(shs=kwA["shs"], **args)
            vm.copy_(kwB["viewmatrix"].view(4, 4)); pm.copy_(kwB["projmatrix"].view(4, 4))  # the SAME tensors now hold another camera
            b, _ = dgr.GaussianRasterizer(rs(kwA["bg"], 0))(colors_precomp=col, **args)
        torch.cuda.synchronize()
        assert torch.isfinite(a).all() and torch.isnan(b).all()
        dgr.release_shared_geometry()
        c, _ = dgr.GaussianRasterizer(rs(kwA["bg"], 0))(colors_precomp=col, **args)  # a full forward through the new camera
        vm, pm = kwB["viewmatrix"].view(4, 4).clone(), kwB["projmatrix"].view(4, 4).clone()
        ref, _ = dgr.GaussianRasterizer(rs(kwA["bg"], 0))(colors_precomp=col, **args)  # the same camera from other memory, outside any block
        torch.cuda.synchronize()
    assert torch.isfinite(c).all() and torch.equal(c, ref)
