"""Developer tool: run-to-run spread of the gradient error vs the live reference (both sides sum with fp32 atomics)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as Hh
from gaustar_b200 import capi
from oracle import refgpu, oracle as O
import test_parity_gpu as T

name = sys.argv[1] if len(sys.argv) > 1 else "random_closeup_odd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = T.SCENES[name]()
kw = Hh.to_torch_kwargs(d)
for mode in (0, 1):
    capi.set_hit_log(mode)
    worst = {}
    for r in range(reps):
        fwd = capi.forward(**kw); torch.cuda.synchronize()
        if mode and not capi.hit_log_state(fwd)[2]:
            fwd = capi.forward(**kw); torch.cuda.synchronize()
        ref = refgpu.forward(**kw)
        dpix = torch.randn(3, d["H"], d["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(2))
        mine = capi.backward(fwd, dpix, **Hh.bwd_kwargs(kw))
        rg = refgpu.backward(ref, dpix, **Hh.bwd_kwargs(kw))
        torch.cuda.synchronize()
        for k in Hh.GRAD_KEYS:
            rr = rg[k].cpu().numpy()
            if rr.size == 0: continue
            e = Hh.rel_err(mine[k].cpu().numpy().reshape(rr.shape), rr)
            worst.setdefault(k, []).append(e)
    print("hit_log" if mode else "walk", {k: f"{np.median(v):.1e}/{np.max(v):.1e}" for k, v in worst.items()}, "in_use", capi.hit_log_state(fwd)[2])
# reference vs itself
errs = {}
ref = refgpu.forward(**kw)
dpix = torch.randn(3, d["H"], d["W"], device="cuda", generator=torch.Generator("cuda").manual_seed(2))
base = {k: v.cpu().numpy() for k, v in refgpu.backward(ref, dpix, **Hh.bwd_kwargs(kw)).items()}
for r in range(reps):
    rg = refgpu.backward(ref, dpix, **Hh.bwd_kwargs(kw))
    for k in Hh.GRAD_KEYS:
        if base[k].size: errs.setdefault(k, []).append(Hh.rel_err(rg[k].cpu().numpy(), base[k]))
print("reference vs reference", {k: f"{np.max(v):.1e}" for k, v in errs.items()})
