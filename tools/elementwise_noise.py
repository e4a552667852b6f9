"""Developer tool: distribution of the worst ELEMENT-WISE relative gradient error (entries above 1 % of the tensor's max) against the
fp64 CPU oracle, over repeated runs: this library's default backward, its deterministic mode, and the unmodified reference."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hh
from gaustar_b200 import capi
from oracle import refgpu, oracle as O
import test_parity_gpu as T

reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["random_closeup_odd", "random_big_sh2", "surface_sh3"]):
    d = T.SCENES[name]()
    kw = Hh.to_torch_kwargs(d)
    capi.set_hit_log(1)
    fwd = capi.forward(**kw); torch.cuda.synchronize()
    if not capi.hit_log_state(fwd)[2]:
        fwd = capi.forward(**kw); torch.cuda.synchronize()
    inp = Hh.oracle_inputs_from_dict(d)
    of = O.forward(inp)
    st = capi.image_state(fwd, d["W"], d["H"])
    of.n_contrib = st["n_contrib"].cpu().numpy().astype(np.uint32).reshape(of.n_contrib.shape)
    of.final_T = st["final_T"].cpu().numpy().reshape(of.final_T.shape).copy()
    dpix = np.random.default_rng(1).normal(0, 1, (3, d["H"], d["W"])).astype(np.float32)
    ob = O.backward(inp, of, dpix).__dict__
    dp = torch.from_numpy(dpix).cuda()
    rf = refgpu.forward(**kw)
    rows = {"ours": {}, "ours_det": {}, "reference": {}}
    for r in range(reps):
        for arm in rows:
            if arm == "reference":
                g = refgpu.backward(rf, dp, **Hh.bwd_kwargs(kw))
            else:
                capi.set_deterministic(arm == "ours_det")
                g = capi.backward(fwd, dp, **Hh.bwd_kwargs(kw))
                capi.set_deterministic(False)
            torch.cuda.synchronize()
            for k in Hh.GRAD_KEYS:
                ref = np.asarray(ob[k])
                if ref.size == 0: continue
                rows[arm].setdefault(k, []).append((T.elementwise_worst(g[k].cpu().numpy(), ref), Hh.rel_err(g[k].cpu().numpy().reshape(ref.shape), ref)))
    print(name)
    for k in rows["ours"]:
        print(f"  {k:14s}", "  ".join(f"{arm}: elem med {np.median([x[0] for x in v[k]]):.1e} max {np.max([x[0] for x in v[k]]):.1e} | maxnorm {np.max([x[1] for x in v[k]]):.1e}" for arm, v in rows.items()))
