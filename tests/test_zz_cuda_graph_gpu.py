"""CUDA-graph capture of the rasterizer ("sync-free, stream-correct, CUDA-graph-capturable fwd+bwd").

The reference cannot be captured: it blocks on a device->host copy of num_rendered in the middle of its forward
(rasterizer_impl.cu:281).  Here the forward detects that its stream is being captured and stays on the device (the binning
capacity comes from earlier un-captured calls; the return value is that capacity).  Bar: replays reproduce the eager
calls bit for bit in the forward and within the atomics' spread in the backward, also after the camera tensors were
overwritten in place (the same graph renders another view).
"""
import numpy as np
import pytest
import torch

from gaustar_b200 import capi, scene

import helpers as Hh
from test_parity_gpu import LIVE_REF_TOL, backward_path, check_grads, run_mine  # noqa: F401 (backward_path: autouse fixture)

pytestmark = pytest.mark.gpu


def _np_grads(g):
    return {k: g[k].cpu().numpy() for k in Hh.GRAD_KEYS}


def test_forward_and_backward_replay_from_one_graph_for_two_cameras():
    g = scene.surface_gaussians(20000, 3, seed=1)
    cams = scene.dome_cameras(6, 400, 225)
    dA, dB = Hh.scene_dict(g, cams[2]), Hh.scene_dict(g, cams[4])
    kwA, eagerA = run_mine(dA)
    kwB, eagerB = run_mine(dB)
    assert kwA["tan_fovx"] == kwB["tan_fovx"] and kwA["tan_fovy"] == kwB["tan_fovy"]  # scalars are baked into the graph
    H, W = kwA["H"], kwA["W"]
    dpix = torch.randn(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    gA = _np_grads(capi.backward(eagerA, dpix, **Hh.bwd_kwargs(kwA)))
    gB = _np_grads(capi.backward(eagerB, dpix, **Hh.bwd_kwargs(kwB)))
    torch.cuda.synchronize()

    static = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in kwA.items()}
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        f = capi.forward(**static)
        gr = capi.backward(f, dpix, **Hh.bwd_kwargs(static))
    assert f["num_rendered"] >= max(eagerA["num_rendered"], eagerB["num_rendered"])  # the provisioned capacity, not R

    for kw, eager, gref in ((kwA, eagerA, gA), (kwB, eagerB, gB), (kwA, eagerA, gA)):
        for k in ("viewmatrix", "projmatrix", "campos"):
            static[k].copy_(kw[k])
        graph.replay()
        torch.cuda.synchronize()
        hdr = capi.debug_header(f)
        assert hdr["overflow"] == 0 and hdr["num_rendered"] == eager["num_rendered"]
        assert torch.equal(f["out_color"], eager["out_color"])
        assert torch.equal(f["radii"], eager["radii"])
        st, se = capi.image_state(f, W, H), capi.image_state(eager, W, H)
        assert torch.equal(st["n_contrib"], se["n_contrib"]) and torch.equal(st["final_T"], se["final_T"])
        check_grads(gr, gref, per_key=LIVE_REF_TOL)


def test_operator_inference_and_reblend_replay_from_a_graph():
    """The public operator under no_grad, RGB pass + a re-blended colors_precomp pass inside shared_geometry(), captured
    once and replayed: same images as the eager calls."""
    import diff_gaussian_rasterization as dgr
    g = scene.surface_gaussians(12000, 3, seed=4)
    cam = scene.dome_cameras(6, 320, 200)[1]
    kw = Hh.to_torch_kwargs(Hh.scene_dict(g, cam))
    col = torch.rand(kw["means3D"].shape[0], 3, device="cuda")
    bg2 = torch.full((3,), 5.0, device="cuda")

    def settings(bg, deg):
        return dgr.GaussianRasterizationSettings(kw["H"], kw["W"], kw["tan_fovx"], kw["tan_fovy"], bg, 1.0, kw["viewmatrix"].view(4, 4),
                                                 kw["projmatrix"].view(4, 4), deg, kw["campos"], False, False)

    def render():
        with torch.no_grad(), dgr.shared_geometry():
            m2 = torch.zeros_like(kw["means3D"])
            a, _ = dgr.GaussianRasterizer(settings(kw["bg"], 3))(means3D=kw["means3D"], means2D=m2, opacities=kw["opacities"], shs=kw["shs"],
                                                                scales=kw["scales"], rotations=kw["rotations"])
            b, _ = dgr.GaussianRasterizer(settings(bg2, 0))(means3D=kw["means3D"], means2D=m2, opacities=kw["opacities"], colors_precomp=col,
                                                           scales=kw["scales"], rotations=kw["rotations"])
        return a, b

    ea, eb = render()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ga, gb = render()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(ga, ea) and torch.equal(gb, eb)
    col.copy_(1.0 - col)  # new colours in place: the replay re-blends them
    graph.replay()
    torch.cuda.synchronize()
    _, eb2 = render()
    torch.cuda.synchronize()
    assert torch.equal(gb, eb2) and not torch.equal(eb2, eb)


def test_captured_replay_that_outgrows_its_capacity_poisons_the_image():
    """A replayed graph cannot grow its binning buffer: a view that needs more instances than the captured capacity sets the
    header's overflow flag and blends nothing -- the image must then be NaN, not uninitialised memory.  (The capacity a capture
    takes is per host thread; a fresh thread starts from the minimum, so the case is reachable here.)"""
    import threading
    result = {}

    def body():
        try:
            torch.cuda.set_device(0)
            g = scene.surface_gaussians(250000, 1, seed=2)
            cam = scene.dome_cameras(4, 960, 540)[1]
            kw = Hh.to_torch_kwargs(Hh.scene_dict(g, cam))
            real = kw["means3D"].clone()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                kw["means3D"].add_(1000.0)  # nothing on screen: the thread's capacity estimate stays at its minimum
                f0 = capi.forward(**kw)
                assert f0["num_rendered"] == 0
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=s):
                    f = capi.forward(**kw)
                kw["means3D"].copy_(real)
                graph.replay()
                torch.cuda.synchronize()
                hdr = capi.debug_header(f)
                result.update(overflow=hdr["overflow"], R=hdr["num_rendered"], cap=hdr["capacity"], nan=bool(torch.isnan(f["out_color"]).all()))
                kw["means3D"].add_(1000.0)  # and a replay that fits again is a picture again (the background)
                graph.replay()
                torch.cuda.synchronize()
                result["bg_again"] = bool(torch.equal(f["out_color"], f0["out_color"]))
        except Exception as e:  # noqa: BLE001 (reported through the assert below)
            result["error"] = repr(e)

    t = threading.Thread(target=body)
    t.start()
    t.join()
    assert "error" not in result, result
    assert result["R"] > result["cap"] and result["overflow"] == 1 and result["nan"] and result["bg_again"], result
