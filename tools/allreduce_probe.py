"""Developer tool: time of the step's one collective -- a sum-allreduce of the flat gradient buffer (59 fp32 per Gaussian, 236 MB at
1 M Gaussians) -- under torchrun, NCCL as configured by the environment, and (optionally) torch's symmetric-memory kernels.

    torchrun --nproc-per-node 8 tools/allreduce_probe.py [--symm]
"""
import argparse, json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaustar_b200 import dist as gdist

ap = argparse.ArgumentParser()
ap.add_argument("--floats", type=int, default=59 * 1000002)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--symm", action="store_true")
a = ap.parse_args()
rank, world, local = gdist.init_from_env()
dev = torch.device("cuda", local)


def timed(fn, label):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = gdist.max_over_ranks(e0.elapsed_time(e1) / a.iters, dev)
    if rank == 0:
        gb = a.floats * 4 / 1e9
        print(json.dumps({"what": label, "world": world, "MB": round(gb * 1e3, 1), "ms": round(ms, 4), "algbw_GBs": round(gb / (ms * 1e-3), 1),
                          "busbw_GBs": round(gb / (ms * 1e-3) * 2 * (world - 1) / world, 1),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}), flush=True)


buf = torch.ones(a.floats, device=dev)
timed(lambda: dist.all_reduce(buf), "nccl all_reduce")
if a.symm:
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(a.floats, dtype=torch.float32, device=dev)
        symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        t.fill_(1.0)
        for op in ("multimem_all_reduce_", "two_shot_all_reduce_"):
            if hasattr(torch.ops.symm_mem, op):
                fn = getattr(torch.ops.symm_mem, op)
                timed(lambda: fn(t, "sum", dist.group.WORLD.group_name), "symm_mem." + op)
    except Exception as e:  # not available in this build / on this box
        if rank == 0:
            print(json.dumps({"what": "symm_mem", "error": repr(e)[:300]}), flush=True)
dist.barrier()
dist.destroy_process_group()
