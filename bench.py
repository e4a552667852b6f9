#!/usr/bin/env python
"""bench.py -- forward+backward views/sec of the surface-Gaussian rasterizer (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic views: every rank renders
VIEWS_PER_GPU camera views of the SAME replicated 1M mesh-bound Gaussians at 1920x1080 (forward +
backward, SH degree 3 evaluated in-kernel), accumulates the per-Gaussian parameter gradients of its
views into one flat buffer and (N > 1) joins the ranks with ONE NCCL sum-allreduce of that buffer.
Per-GPU work is fixed as N grows ("weak").

  value : views/s, inputs resident in HBM, calls made straight through the C ABI
          (gaustar_b200/capi.py -> libgstar_raster.so); CUDA events, max over ranks.
  e2e   : the same metric through the public operator API (diff_gaussian_rasterization.GaussianRasterizer
          + autograd) with, per view, the camera and the 8-bit target image copied from pinned host
          memory (H2D) and, per step, the loss read back (D2H) -- all inside the timed region.
  roofline     : the dominant kernel (blend_bwd) timed live with CUDA events on the launching stream
                 through the library's profile hook; achieved = algorithmic bytes / time.
  cpu_baseline : the CPU oracle (oracle/, a C restatement of the reference; kind "port") on the host
                 cores, bounded sample, rank 0 at N=1 only.  The reference rasterizer has no CPU path.

--impl reference times the UNMODIFIED reference rasterizer (oracle/_ref/ref_dgr_C.so, the reference's
own pybind module built for sm_100a by oracle/build_ref.py) on the same GPU with the same protocol.
If that module or a GPU is unavailable it falls back to timing the CPU oracle port.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gaustar_b200 import scene  # noqa: E402
from gaustar_b200 import dist as gdist  # noqa: E402

METRIC = "fwd+bwd views/sec at 1M surface Gaussians, 1080p; HBM GB/s vs roofline"
UNIT = "views/s"
P_TARGET, W, H, SH_DEG = 1_000_000, 1920, 1080, 3
VIEWS_PER_GPU = 20  # BASELINE.json config #3: 160 views / 8 GPUs per step
CAM_POOL = 32
NUM_SMS_FALLBACK = 148


def kernel_sources_sha():
    """Identity of the kernels a capture belongs to: sha256 over gaustar_b200/csrc (sorted file names and contents)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gaustar_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if not os.path.isfile(os.path.join(d, f)) or f.startswith("."):
            continue
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_capture():
    """dram__bytes_read+write and smsp__inst_executed per launch of every kernel, from profiles/ncu_capture.json -- written by
    tools/ncu_capture.py from an `ncu --set full` capture of this workload together with the sha of the kernel sources it was
    taken from.  Returned only when that sha equals the sources' sha now: a capture of other code is not a measurement of this
    code (round 1 printed a constant that predated two kernel rewrites)."""
    path = os.path.join(ROOT, "profiles", "ncu_capture.json")
    try:
        cap = json.load(open(path))
    except Exception:
        return None, "no profiles/ncu_capture.json"
    sha = kernel_sources_sha()
    if cap.get("sources_sha") != sha:
        return None, f"profiles/ncu_capture.json was taken from kernel sources {cap.get('sources_sha')}, these are {sha}: traffic withheld"
    return cap, f"profiles/ncu_capture.json ({cap.get('capture')}, sources {sha})"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(device, rank, world):
    g = scene.surface_gaussians(P_TARGET, sh_degree=SH_DEG, seed=0)
    cams = scene.dome_cameras(CAM_POOL, W, H)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    params = dict(means3D=t(g.means3D), scales=t(g.scales), rotations=t(g.rotations), opacities=t(g.opacities), shs=t(g.shs))
    bg = torch.tensor([0.0, 1.0, 0.0], device=device)
    # 8-bit targets (what a dataset on disk holds): a small pool in pinned host memory
    rng = np.random.default_rng(1 + rank)
    targets = [torch.from_numpy(np.clip(rng.normal(0.5, 0.2, (H, W, 3)) * 255, 0, 255).astype(np.uint8)).pin_memory() for _ in range(4)]
    cam_host = [dict(viewmatrix=torch.from_numpy(c.viewmatrix).pin_memory(), projmatrix=torch.from_numpy(c.projmatrix).pin_memory(),
                     campos=torch.from_numpy(c.campos).pin_memory(), tanfovx=c.tanfovx, tanfovy=c.tanfovy) for c in cams]
    cam_dev = [dict(viewmatrix=c["viewmatrix"].to(device), projmatrix=c["projmatrix"].to(device), campos=c["campos"].to(device),
                    tanfovx=c["tanfovx"], tanfovy=c["tanfovy"]) for c in cam_host]
    return g, params, bg, targets, cam_host, cam_dev


# ------------------------------------------------------------------------------------------------
# the two implementations behind one tiny interface
# ------------------------------------------------------------------------------------------------
class OursCABI:
    """Device-resident arm: straight through the C ABI."""
    name = "gaustar_b200 (C ABI)"
    launches_per_view = 9  # preprocess_fwd, tile_scan, emit, tile_sort, blend_fwd, blend_fwd_fat (no marked tile here), blend_bwd_gather, blend_bwd (no-op), preprocess_bwd

    def __init__(self):
        from gaustar_b200 import capi
        self.capi = capi
        capi.lib()

    fused_accumulate = True  # K8 adds each view's parameter gradients straight into the flat allreduce buffer

    atomic = False  # several streams add into ONE flat buffer (accumulate_param_grads = 2)

    def fwd_bwd(self, params, cam, bg, grad_fn, flat=None):
        kw = dict(means3D=params["means3D"], viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], campos=cam["campos"], bg=bg,
                  tan_fovx=cam["tanfovx"], tan_fovy=cam["tanfovy"], shs=params["shs"], scales=params["scales"], rotations=params["rotations"],
                  sh_degree=SH_DEG)
        f = self.capi.forward(opacities=params["opacities"], W=W, H=H, **kw)
        g = self.capi.backward(f, grad_fn(f["out_color"]), accumulate_into=flat.views if flat is not None else None, lean=True,
                               atomic_accumulate=self.atomic, **kw)
        return f["num_rendered"], int(0), g


class RefStock:
    """The unmodified reference through its own pybind module (stock binding, reference kernels)."""
    name = "reference diff-gaussian-rasterization (oracle/_ref/ref_dgr_C.so, sm_100a)"
    launches_per_view = 0

    def __init__(self):
        from oracle import refgpu
        self.C = refgpu.stock_module()
        self.empty = torch.Tensor([])

    fused_accumulate = False

    def fwd_bwd(self, params, cam, bg, grad_fn, flat=None):
        C, e = self.C, self.empty
        R, color, radii, gb, bb, ib = C.rasterize_gaussians(bg, params["means3D"], e, params["opacities"], params["scales"], params["rotations"], 1.0, e,
                                                            cam["viewmatrix"], cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], H, W, params["shs"],
                                                            SH_DEG, cam["campos"], False, False)
        out = C.rasterize_gaussians_backward(bg, params["means3D"], radii, e, params["scales"], params["rotations"], 1.0, e, cam["viewmatrix"],
                                             cam["projmatrix"], cam["tanfovx"], cam["tanfovy"], grad_fn(color), params["shs"], SH_DEG, cam["campos"],
                                             gb, R, bb, ib, False)
        g = dict(dL_dmeans2D=out[0], dL_dcolors=out[1], dL_dopacity=out[2], dL_dmeans3D=out[3], dL_dcov3D=out[4], dL_dsh=out[5], dL_dscales=out[6],
                 dL_drotations=out[7])
        return R, 0, g


def make_autograd_rasterizer(impl):
    """Public-API arm.  ours: the shipped drop-in module.  reference: the reference's own wrapper logic
    (DGR/diff_gaussian_rasterization/__init__.py:44-155) around its stock pybind module."""
    if impl == "ours":
        import diff_gaussian_rasterization as d
        return d.GaussianRasterizationSettings, d.GaussianRasterizer
    from oracle import refgpu
    C = refgpu.stock_module()
    import diff_gaussian_rasterization as d  # only for the settings NamedTuple (plain data)

    class _RefFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
            R, color, radii, gb, bb, ib = C.rasterize_gaussians(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                                                cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                                                                rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug)
            ctx.rs, ctx.R = rs, R
            ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, gb, bb, ib)
            return color, radii

        @staticmethod
        def backward(ctx, g, _):
            rs = ctx.rs
            colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, gb, bb, ib = ctx.saved_tensors
            o = C.rasterize_gaussians_backward(rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                                               rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, g, sh, rs.sh_degree, rs.campos, gb, ctx.R, bb, ib,
                                               rs.debug)
            return o[3], o[0], o[5], o[1], o[2], o[6], o[7], o[4], None

    class RefRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.raster_settings = raster_settings

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
            e = torch.Tensor([])
            return _RefFn.apply(means3D, means2D, shs if shs is not None else e, colors_precomp if colors_precomp is not None else e, opacities,
                                scales if scales is not None else e, rotations if rotations is not None else e,
                                cov3D_precomp if cov3D_precomp is not None else e, self.raster_settings)

    return d.GaussianRasterizationSettings, RefRasterizer


# ------------------------------------------------------------------------------------------------
class _L1Mean(torch.autograd.Function):
    """mean |img - target| with four elementwise/reduction kernels instead of the seven of
    torch.nn.functional.l1_loss + autograd (same value and gradient); used by BOTH arms of the e2e measurement."""

    @staticmethod
    def forward(ctx, img, target):
        d = img - target
        ctx.save_for_backward(d)
        return torch.linalg.vector_norm(d, ord=1) / d.numel()

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return torch.sign(d, out=d).mul_(g / d.numel()), None


def l1_grad_fn(target_f):
    """dL/dcolor of L = mean |render - target| (SURVEY 8d upstream gradient)."""
    inv = 1.0 / (3 * H * W)
    return lambda img: torch.sign(img - target_f) * inv


def cameras_of_step(step, rank, world):
    """Camera indices (into the pool of CAM_POOL) of the views rank `rank` renders in step `step` (see the comment in run_gpu)."""
    views = gdist.views_for_rank(VIEWS_PER_GPU * world, rank, world)
    if os.environ.get("GSTAR_BENCH_CAMERA_MAP") == "round1":
        return [(step * VIEWS_PER_GPU * world + v) % CAM_POOL for v in views]
    return [(step * VIEWS_PER_GPU + v // world + (v % world) * (CAM_POOL // world)) % CAM_POOL for v in views]


def run_gpu(args, impl_name, rank, world, local):
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    g, params, bg, targets, cam_host, cam_dev = build_workload(device, rank, world)
    P, M = g.P, g.shs.shape[1]
    impl = OursCABI() if impl_name == "ours" else RefStock()
    target_f = [(t.to(device).permute(2, 0, 1).float() / 255.0).contiguous() for t in targets[:2]]
    flat = gdist.FlatGrads(P, M, device)
    # The step's views are partitioned round robin (rank r takes the view indices v = r, r + N, ...: SURVEY 8e); view v looks through
    # camera (step * 20 + v // N + (v % N) * 32 / N) mod 32, so that every rank walks 20 CONSECUTIVE cameras of the pool, a window that
    # moves with the step -- over a step every camera is still rendered the same number of times.  (Camera = v mod 32, round 1's
    # choice, pinned rank r to the four cameras r, r + 8, r + 16, r + 24 for the whole run: the rank with the heaviest four was the
    # straggler of every step; GSTAR_BENCH_CAMERA_MAP=round1 selects it for an A/B -- 13.87 vs 13.76 ms per step on the same 8-GPU
    # box.)  One GPU: the same cameras as before.
    my_views = lambda step: cameras_of_step(step, rank, world)

    # ---------------- device-resident arm (value) ----------------
    prof = None
    if impl_name == "ours":
        from gaustar_b200 import capi
        n_ev = args.steps * VIEWS_PER_GPU
        prof = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_ev)]
        for a, b in prof:
            a.record(); b.record()
        torch.cuda.synchronize()
    R_sum, V_count = 0, 0
    # Our arm keeps TWO views of a step in flight on two CUDA streams (the views of a multi-view step are independent;
    # every kernel of the library is launched on the caller's current stream).  Both streams add into ONE flat gradient buffer
    # (accumulate_param_grads = 2: reductions at L2), which is all-reduced once per step; with --per-stream-buffers each stream
    # has a buffer of its own (plain read-modify-write) and the two are summed before the allreduce, as in round 1.  The reference
    # launches on the legacy default stream and blocks on a D2H copy inside every forward, so it runs its views one after the other.
    n_streams = max(1, args.streams) if impl.fused_accumulate else 1
    side = [torch.cuda.Stream(device) for _ in range(n_streams)] if n_streams > 1 else []
    shared_flat = impl.fused_accumulate and n_streams > 1 and not args.per_stream_buffers
    if impl_name == "ours":
        impl.atomic = shared_flat
    flats = [flat] + ([] if shared_flat else [gdist.FlatGrads(P, M, device) for _ in range(n_streams - 1)])
    fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in side]

    def one_view(step, i, v, timed, fl):
        nonlocal R_sum, V_count
        if prof is not None and timed:
            n = (step - args.warmup) * VIEWS_PER_GPU + i
            a, b = prof[n]
            capi.profile_stage(n % len(capi.STAGES), a, b)  # every view times one stage, round robin over the seven
        if impl.fused_accumulate:
            R, _, grads = impl.fwd_bwd(params, cam_dev[v], bg, l1_grad_fn(target_f[i & 1]), fl)
        else:
            R, _, grads = impl.fwd_bwd(params, cam_dev[v], bg, l1_grad_fn(target_f[i & 1]))
            fl.accumulate(grads)
        if timed:
            R_sum += int(R); V_count += 1

    def step_value(step, timed):
        for fl in flats:
            fl.zero_()
        views = my_views(step)
        if not side:
            for i, v in enumerate(views):
                one_view(step, i, v, timed, flat)
        else:
            main = torch.cuda.current_stream(device)
            fork.record(main)
            for st in side:
                st.wait_event(fork)
            for i, v in enumerate(views):
                with torch.cuda.stream(side[i % n_streams]):
                    one_view(step, i, v, timed, flats[i % len(flats)])
            for st, ev in zip(side, joins):
                ev.record(st)
                main.wait_event(ev)
            for fl in flats[1:]:
                flat.flat.add_(fl.flat)
        flat.allreduce()

    for s in range(args.warmup):
        step_value(s, False)
    gdist.barrier(); torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step_value(args.warmup + s, True)
    e1.record()
    gdist.barrier(); torch.cuda.synchronize()
    ms_value = gdist.max_over_ranks(e0.elapsed_time(e1), device)
    clk = clocks.stop() if rank == 0 else None
    if prof is not None:
        capi.profile_stage(-1)
    views_total = VIEWS_PER_GPU * world * args.steps
    value = views_total / (ms_value / 1e3)

    roofline = None
    if prof is not None:
        R_avg = R_sum / max(V_count, 1)
        T_tiles = ((W + 15) // 16) * ((H + 15) // 16)
        npix = W * H
        # algorithmic bytes per launch (SURVEY 8d / BASELINE.md 2.4; V ~= P for this scene: nothing is frustum-culled)
        alg = {"preprocess_fwd": P * (44 + 12 * M) + 8 * P + P * (52 + 12), "tile_scan": 8 * T_tiles, "emit": 20 * P + 12 * R_avg,
               "tile_sort": 24 * R_avg + 8 * R_avg + 8 * T_tiles, "blend_fwd": 40 * R_avg + 20 * npix + 8 * T_tiles,
               "blend_bwd": 40 * R_avg + 20 * npix + 36 * P, "preprocess_bwd": P * (108 + 12 * M) + P * (40 + 12 * M)}
        kernel_of = {"preprocess_fwd": "k_preprocess_fwd", "tile_scan": "k_tile_scan", "emit": "k_emit", "tile_sort": "k_tile_sort",
                     "blend_fwd": "k_blend_fwd", "blend_bwd": "k_blend_bwd_gather", "preprocess_bwd": "k_preprocess_bwd"}
        peak, how = measured_peak()
        per = {}
        for si, name in enumerate(capi.STAGES):
            ts = [a.elapsed_time(b) for n, (a, b) in enumerate(prof) if n % len(capi.STAGES) == si]
            if ts:
                ms = float(np.mean(ts))
                per[name] = {"ms": round(ms, 4), "algorithmic_bytes": int(alg[name]), "gbs": round(alg[name] / (ms * 1e-3) / 1e9, 1),
                             "frac": round(alg[name] / (ms * 1e-3) / 1e9 / peak, 4), "samples": len(ts)}
        # issue-slot ceiling next to the HBM one: warp instructions of the kernel (ncu capture of the same sources)
        # over the issue slots its live duration offered -- SMs x 4 schedulers x SM clock under load
        cap, cap_note = ncu_capture()
        sm_mhz = (clk or {}).get("sm_mhz") or (clk or {}).get("sm_max_mhz") or 1965.0
        n_sms = torch.cuda.get_device_properties(device).multi_processor_count or NUM_SMS_FALLBACK
        cap_kernels = {kn.split("::")[-1]: kv for kn, kv in (cap or {}).get("kernels", {}).items()}  # (ncu prints the namespace: gstar::k_...)
        for name, st_ in per.items():
            k = cap_kernels.get(kernel_of[name])
            st_["traffic"] = k["dram_bytes"] if k else None
            st_["warp_instructions"] = k["warp_inst"] if k else None
            st_["issue_frac"] = round(k["warp_inst"] / (st_["ms"] * 1e-3 * n_sms * 4 * sm_mhz * 1e6), 4) if k else None
        dom = max(per, key=lambda k: per[k]["ms"])
        roofline = {"kernel": kernel_of[dom], "stage": dom, "bound": "hbm", "achieved": per[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": per[dom]["frac"], "traffic": per[dom]["traffic"], "traffic_source": cap_note, "issue_frac": per[dom]["issue_frac"],
                    "peak_source": how, "algorithmic_bytes_per_launch": per[dom]["algorithmic_bytes"],
                    "avg_launch_ms": per[dom]["ms"], "avg_num_rendered": int(R_avg),
                    "note": "dominant (longest) kernel of the step, timed live with CUDA events on its launching stream while the other "
                            "stream's view runs concurrently; blend kernels are SIMT-issue-bound by construction (DESIGN.md section 4)",
                    "stages": per}

    # ---------------- public-API arm (e2e): host buffers, copies inside the timed region ----------------
    Settings, Rasterizer = make_autograd_rasterizer(impl_name)
    name_of = {"means3D": "dL_dmeans3D", "scales": "dL_dscales", "rotations": "dL_drotations", "opacities": "dL_dopacity", "shs": "dL_dsh"}
    base_leaves = {k: params[k].clone() for k in name_of}
    copy_stream = torch.cuda.Stream(device)
    h2d_per_view = H * W * 3 + (16 + 16 + 3) * 4

    def measure_e2e(ns, fusion):
        """views/s through the operator API with `ns` views in flight on `ns` CUDA streams; fusion: gradient accumulation inside the
        backward kernel into the leaves' .grad (ours only; needs LEAF inputs -- GauSTAR's own call sites pass non-leaf tensors)."""
        one_buffer = bool(fusion) and shared_flat and ns > 1  # every stream's leaves share ONE .grad buffer; the kernel adds atomically
        if impl_name == "ours":
            import gaustar_b200
            gaustar_b200.set_grad_accumulation_fusion(bool(fusion), atomic=one_buffer)
        streams = [torch.cuda.Stream(device) for _ in range(ns)] if ns > 1 else [torch.cuda.current_stream(device)]
        nbuf = 1 if one_buffer else ns
        fl_set = flats[:nbuf] if len(flats) >= nbuf else flats + [gdist.FlatGrads(P, M, device) for _ in range(nbuf - len(flats))]
        # one set of autograd leaves per compute stream (same storage, separate .grad buffers = the per-stream flat buffers),
        # so that each stream's AccumulateGrad nodes run on that stream and never touch the other stream's buffer
        leaf_sets, m2d_sets = [], []
        for si, st in enumerate(streams):
            with torch.cuda.stream(st):
                ls = {k: base_leaves[k].detach().requires_grad_(True) for k in name_of}
                for k, p_ in ls.items():
                    p_.grad = fl_set[si % nbuf].views[name_of[k]]  # autograd accumulates in place into the flat allreduce buffer
                leaf_sets.append(ls)
                m2d_sets.append(torch.zeros(P, 3, device=device, requires_grad=True))
        torch.cuda.synchronize()
        slots = [dict(tgt=torch.empty(H, W, 3, dtype=torch.uint8, device=device), tgt_f=torch.empty(3, H, W, device=device), vm=torch.empty(4, 4, device=device),
                      pm=torch.empty(4, 4, device=device), cp=torch.empty(3, device=device), ev=torch.cuda.Event(), free=torch.cuda.Event())
                 for _ in range(ns + 1)]
        NSLOT = len(slots)
        fork_e, joins_e = torch.cuda.Event(), [torch.cuda.Event() for _ in streams]

        def prefetch(slot, v, i):
            """H2D of the next view's 8-bit target + camera on the copy stream, and its uint8 -> float CHW conversion there too."""
            sl = slots[slot]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(sl["free"])
                sl["tgt"].copy_(targets[i % len(targets)], non_blocking=True)
                sl["vm"].copy_(cam_host[v]["viewmatrix"], non_blocking=True)
                sl["pm"].copy_(cam_host[v]["projmatrix"], non_blocking=True)
                sl["cp"].copy_(cam_host[v]["campos"], non_blocking=True)
                torch.mul(sl["tgt"].permute(2, 0, 1), 1.0 / 255.0, out=sl["tgt_f"])  # uint8 HWC -> float CHW in one kernel
                sl["ev"].record(copy_stream)

        loss_acc = [torch.zeros((), device=device) for _ in streams]

        def step_e2e(step):
            main = torch.cuda.current_stream(device)
            for fl in fl_set:
                fl.zero_()
            for la in loss_acc:
                la.zero_()
            views = my_views(step)
            prefetch(0, views[0], 0)
            if ns > 1:
                fork_e.record(main)
                for st in streams:
                    st.wait_event(fork_e)
            for i, v in enumerate(views):
                sl = slots[i % NSLOT]
                if i + 1 < len(views):
                    prefetch((i + 1) % NSLOT, views[i + 1], i + 1)
                st, ls = streams[i % ns], leaf_sets[i % ns]
                with torch.cuda.stream(st):
                    st.wait_event(sl["ev"])
                    rs = Settings(image_height=H, image_width=W, tanfovx=cam_host[v]["tanfovx"], tanfovy=cam_host[v]["tanfovy"], bg=bg, scale_modifier=1.0,
                                  viewmatrix=sl["vm"], projmatrix=sl["pm"], sh_degree=SH_DEG, campos=sl["cp"], prefiltered=False, debug=False)
                    img, _radii = Rasterizer(rs)(means3D=ls["means3D"], means2D=m2d_sets[i % ns], opacities=ls["opacities"], shs=ls["shs"],
                                                 scales=ls["scales"], rotations=ls["rotations"])
                    loss = _L1Mean.apply(img, sl["tgt_f"])
                    loss.backward()
                    loss_acc[i % ns] += loss.detach()
                    sl["free"].record(st)
            if ns > 1:
                for st, ev in zip(streams, joins_e):
                    ev.record(st)
                    main.wait_event(ev)
                for fl in fl_set[1:]:
                    fl_set[0].flat.add_(fl.flat)
            fl_set[0].allreduce()
            return float(sum(loss_acc).item())  # D2H read of the step's result

        for e in slots:
            e["free"].record(torch.cuda.current_stream(device))
        e2e_steps = max(1, args.steps)
        for s_ in range(min(args.warmup, 3)):
            step_e2e(s_)
        gdist.barrier(); torch.cuda.synchronize()
        e0.record()
        for s_ in range(e2e_steps):
            step_e2e(args.warmup + s_)
        e1.record()
        gdist.barrier(); torch.cuda.synchronize()
        ms = gdist.max_over_ranks(e0.elapsed_time(e1), device)
        del leaf_sets, m2d_sets, slots
        return VIEWS_PER_GPU * world * e2e_steps / (ms / 1e3)

    ours = impl_name == "ours"
    ns_main = n_streams if ours else 1
    e2e_val = measure_e2e(ns_main, fusion=ours)
    e2e = {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d_per_view * VIEWS_PER_GPU, "d2h_bytes_per_step": 4,
           "api": "diff_gaussian_rasterization.GaussianRasterizer + autograd; target image (uint8) and camera copied from pinned host memory per view "
                  "(prefetched on a copy stream), loss scalar read back per step; compute streams: " + str(ns_main)
                  + ("; gradient-accumulation fusion into the leaves' .grad (set_grad_accumulation_fusion)" if ours else "")}
    if ours and not args.quick:
        # the same measurement under the conditions an UNMODIFIED GauSTAR call site has: its rasterizer inputs are non-leaf tensors,
        # so the fusion cannot apply, and it renders one view at a time on one stream -- the reference arm's own conditions
        e2e["value_unfused"] = round(measure_e2e(ns_main, fusion=False), 2)
        e2e["value_like_for_like"] = round(measure_e2e(1, fusion=False), 2)
        e2e["like_for_like"] = "1 view in flight on 1 stream, ordinary AccumulateGrad: the conditions of the reference arm and of unmodified GauSTAR call sites"
        import gaustar_b200
        gaustar_b200.set_grad_accumulation_fusion(True)

    # ---------------- GauSTAR's own training step (refine.py:552-616): RGB, then depth, through colors_precomp ----------------
    refine = None
    if not args.quick:
        refine = measure_refine_step(args, impl_name, Settings, Rasterizer, params, bg, cam_host, targets, device, P)
    return dict(value=value, ms_per_step=ms_value / args.steps, roofline=roofline, e2e=e2e, clocks=clk, P=P, M=M, device_name=torch.cuda.get_device_name(device),
                launches=impl.launches_per_view * VIEWS_PER_GPU * args.steps, impl_desc=impl.name, allreduce_bytes=flat.nbytes, refine_step=refine)


def measure_refine_step(args, impl_name, Settings, Rasterizer, params, bg, cam_host, targets, device, P, iters=60):
    """iterations/s of the step gaustar_trainers/refine.py runs (:529-841, one view per iteration, single stream): the camera goes
    host -> device, the Gaussians are rendered TWICE from it -- RGB with colours precomputed outside the rasterizer
    (compute_color_in_rasterizer=False: colors_precomp, no SH in the op; :552-564) on bg [0,1,0], then view depth as three equal
    colour channels on bg [10,10,10] (:602-616) -- an L1 loss on both, one backward, and the host reads the loss (the trainer
    logs it).  Rasterizer inputs are NON-LEAF tensors as in SuGaR (points/scaling/quaternions are computed per call)."""
    leaves = {k: params[k].clone().requires_grad_(True) for k in ("means3D", "scales", "rotations", "opacities")}
    rgb_leaf = torch.rand(P, 3, device=device).requires_grad_(True)
    bg_depth = torch.full((3,), 10.0, device=device)
    tgt = (targets[0].to(device).permute(2, 0, 1).float() / 255.0).contiguous()
    tgt_depth = torch.full((3, H, W), 3.0, device=device)
    vm_d, pm_d, cp_d = torch.empty(4, 4, device=device), torch.empty(4, 4, device=device), torch.empty(3, device=device)
    import contextlib
    variants = [("unchanged", contextlib.nullcontext)]
    if impl_name == "ours":
        import diff_gaussian_rasterization as dgr
        variants.append(("shared_geometry", dgr.shared_geometry))  # the one-line edit: `with shared_geometry():` around the two calls
        variants.append(("forward_passes", contextlib.nullcontext))  # the two calls replaced by ONE forward_passes() call: both passes in one blend
    out = {}
    for vname, ctx in variants:
        def one_iter(it):
            v = it % len(cam_host)
            vm_d.copy_(cam_host[v]["viewmatrix"], non_blocking=True)
            pm_d.copy_(cam_host[v]["projmatrix"], non_blocking=True)
            cp_d.copy_(cam_host[v]["campos"], non_blocking=True)
            m, sc, rot, op = (leaves[k] * 1.0 for k in ("means3D", "scales", "rotations", "opacities"))  # non-leaf, like SuGaR's properties
            col = rgb_leaf * 1.0
            m2d = torch.zeros(P, 3, device=device, requires_grad=True)
            mk = lambda b: Settings(image_height=H, image_width=W, tanfovx=cam_host[v]["tanfovx"], tanfovy=cam_host[v]["tanfovy"], bg=b, scale_modifier=1.0,
                                    viewmatrix=vm_d, projmatrix=pm_d, sh_degree=0, campos=cp_d, prefiltered=False, debug=False)
            if vname == "forward_passes":
                depth = (m @ vm_d[:3, 2] + vm_d[3, 2])[:, None].expand(-1, 3).contiguous()
                img, _, (dimg,) = Rasterizer(mk(bg)).forward_passes(means3D=m, means2D=m2d, opacities=op, colors_precomp=col, scales=sc, rotations=rot,
                                                                    extra_passes=[(depth, bg_depth)])
            else:
              with ctx():
                img, _ = Rasterizer(mk(bg))(means3D=m, means2D=m2d, opacities=op, colors_precomp=col, scales=sc, rotations=rot)
                depth = (m @ vm_d[:3, 2] + vm_d[3, 2])[:, None].expand(-1, 3)
                dimg, _ = Rasterizer(mk(bg_depth))(means3D=m, means2D=m2d, opacities=op, colors_precomp=depth, scales=sc, rotations=rot)
            loss = (img - tgt).abs().mean() + (dimg - tgt_depth).abs().mean()
            loss.backward()
            for t_ in list(leaves.values()) + [rgb_leaf]:
                t_.grad = None
            return float(loss.item())

        for it in range(5):
            one_iter(it)
        torch.cuda.synchronize()
        t0 = time.time()
        for it in range(iters):
            one_iter(5 + it)
        torch.cuda.synchronize()
        out[vname] = round(iters / (time.time() - t0), 2)
    return {"workload": f"GauSTAR refine step: 1 view/iteration, colors_precomp RGB pass + depth pass (2 fwd + 2 bwd) at {P} Gaussians {W}x{H}, "
                        "non-leaf inputs, single stream, loss read back every iteration (refine.py:552-616)",
            "unit": "iterations/s", "value": out["unchanged"], "value_with_shared_geometry": out.get("shared_geometry"),
            "value_with_forward_passes": out.get("forward_passes"), "iterations": iters,
            "timing": "host wall clock around the loop (the step includes host work by design)"}


def cpu_oracle_views_per_sec(n_views=1, small=False):
    """The CPU oracle (C restatement of the reference, OpenMP) on a bounded sample of the workload."""
    from oracle import oracle as O
    O.build()
    Pn, w, h = (30000, 480, 270) if small else (P_TARGET, W, H)
    g = scene.surface_gaussians(Pn, sh_degree=SH_DEG, seed=0)
    cams = scene.dome_cameras(CAM_POOL, w, h)
    rng = np.random.default_rng(1)
    t0 = time.time()
    for v in range(n_views):
        c = cams[v]
        inp = O.Inputs(means3D=g.means3D, opacities=g.opacities, viewmatrix=c.viewmatrix, projmatrix=c.projmatrix, campos=c.campos,
                       bg=np.array([0, 1, 0], np.float32), tan_fovx=c.tanfovx, tan_fovy=c.tanfovy, W=w, H=h, shs=g.shs, scales=g.scales,
                       rotations=g.rotations, sh_degree=SH_DEG)
        f = O.forward(inp)
        target = np.clip(rng.normal(0.5, 0.2, (3, h, w)), 0, 1).astype(np.float32)
        dpix = (np.sign(f.out_color - target) / (3 * h * w)).astype(np.float32)
        O.backward(inp, f, dpix)
    dt = time.time() - t0
    return n_views / dt, dt, f"{n_views} view(s) fwd+bwd of P={g.P} {w}x{h} SH{SH_DEG} (same generator as the GPU workload)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2, help="views in flight per GPU (CUDA streams) in our arm")
    ap.add_argument("--per-stream-buffers", action="store_true", help="our arm: one flat gradient buffer per stream + a merge before the allreduce (round 1) instead of "
                    "ONE buffer that all streams add into atomically")
    ap.add_argument("--quick", action="store_true", help="headline numbers only (skip the unfused / like-for-like e2e variants, the refine step, the CPU baselines)")
    ap.add_argument("--workload", default="headline", choices=["headline", "config3"],
                    help="headline: 1 M Gaussians at 1920x1080 (BASELINE.json's metric); config3: the same Gaussians at 1352x1014 (ActorsHQ shape)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    global W, H
    if args.workload == "config3":
        W, H = 1352, 1014
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
    rank, world, local = gdist.init_from_env()
    have_gpu = torch.cuda.is_available()
    wl_name = "surface-1M-1080p-sh3" if args.workload == "headline" else "config3: surface-1M-1352x1014-sh3 (160 views / 8 GPUs)"
    config = {"workload": f"{wl_name} ({VIEWS_PER_GPU} views/GPU/step, dome cameras, SuGaR-bound Gaussians, L1 upstream grad)",
              "gaussians": None, "resolution": [W, H], "sh_degree": SH_DEG, "views_per_step": VIEWS_PER_GPU * world,
              "parallelism": f"view-sharded dp{world} + 1 allreduce/step; ours: {args.streams} views in flight on {args.streams} CUDA streams per GPU adding into " + ("one buffer per stream, merged" if args.per_stream_buffers else "ONE flat gradient buffer (reductions at L2)"), "l2_policy": "inputs (>= 236 MB of parameters per view) exceed the 126 MB L2"}

    if args.impl == "reference":
        use_gpu_ref = have_gpu and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_dgr_C.so"))
        if not use_gpu_ref:
            if rank != 0:
                return
            v, dt, sample = cpu_oracle_views_per_sec(3)
            line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                    "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": config, "device": "cpu",
                    "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
                    "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "note": "reference CUDA module unavailable here: timed the CPU oracle port instead"}
            print(json.dumps(line), flush=True)
            return
    if not have_gpu:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU oracle baseline)")

    res = run_gpu(args, args.impl, rank, world, local)
    if rank != 0:
        return
    config["gaussians"] = res["P"]
    line = {"metric": METRIC, "value": round(res["value"], 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(res["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "clocks": res["clocks"], "e2e": res["e2e"], "gpu_launches": res["launches"],
            "device": res["device_name"], "allreduce_bytes_per_step": res["allreduce_bytes"] if world > 1 else 0}
    if res.get("refine_step"):
        line["refine_step"] = res["refine_step"]
    if args.impl == "reference":
        line["impl"] = "reference"
        line["implementation"] = res["impl_desc"]
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                "sample": "n/a: the reference rasterizer is CUDA-only (no CPU path); this arm ran it unmodified on the GPU"}
        line["gpu_launches"] = 0
    else:
        line["roofline"] = res["roofline"]
        if world == 1 and not args.no_cpu_baseline and not args.quick:
            v, dt, sample = cpu_oracle_views_per_sec(3)
            line["cpu_baseline"] = {"value": round(v, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
                                    "seconds": round(dt, 1)}
            # the CPU baseline north_star / BASELINE.md 2.2 name: a naive PyTorch point-splat with autograd, at its two stated sizes
            from oracle import naive_splat
            naive = []
            for Pn, wn, hn, nv in ((30000, 128, 128, 4), (200000, 480, 270, 2)):
                v2, dt2, Pgot = naive_splat.time_fwd_bwd(Pn, wn, hn, sh_degree=SH_DEG, views=nv)
                naive.append({"value": round(v2, 4), "unit": UNIT, "cores": os.cpu_count(), "kind": "naive PyTorch point-splat (oracle/naive_splat.py)",
                              "sample": f"{nv} views fwd+bwd (autograd) of P={Pgot} {wn}x{hn} SH{SH_DEG}", "seconds": round(dt2, 1)})
            line["cpu_baseline_naive"] = naive
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _d
        if _d.is_available() and _d.is_initialized():
            _d.destroy_process_group()
