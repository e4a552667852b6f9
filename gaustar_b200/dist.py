"""View-sharded data parallelism for multi-view per-frame optimisation.

The reference is single-GPU and consumes one camera view per iteration
(gaustar_trainers/refine.py:534-548).  The only natural parallel axis of the path is the camera
view: Gaussian parameters are replicated, each rank rasterizes its slice of a step's views, and ONE
sum-allreduce of the per-Gaussian parameter gradients (59 fp32 per Gaussian with SH degree 3:
3 mean + 3 scale + 4 rotation + 1 opacity + 48 SH) joins the ranks per optimiser step.  There is no
other exchange on the path, so nothing here touches the data path of the kernels.

One process per GPU (torchrun); backend "nccl" on GPUs (NVLS over NVSwitch), "gloo" in the CPU tests.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Sequence

import torch
import torch.distributed as dist

GRAD_FIELDS = ("dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dsh")


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def views_for_rank(num_views: int, rank: int, world: int) -> List[int]:
    """Round-robin shard: rank r renders views {v : v mod world == r} (SURVEY 8e 'Partitioning')."""
    return list(range(rank, num_views, world))


class FlatGrads:
    """All per-Gaussian parameter gradients of one step in a single flat fp32 buffer, so that the
    step needs exactly one collective.  Views of the buffer are handed out per field; accumulate()
    adds one view's op-level gradients; allreduce() sums over ranks."""

    def __init__(self, P: int, M: int, device):
        self.shapes = {"dL_dmeans3D": (P, 3), "dL_dscales": (P, 3), "dL_drotations": (P, 4), "dL_dopacity": (P, 1), "dL_dsh": (P, M, 3)}
        sizes = [int(torch.Size(s).numel()) for s in self.shapes.values()]
        # every field starts on a 128-byte boundary: the kernels' 128-bit load/store paths need 16-byte alignment
        starts, off = [], 0
        for n in sizes:
            starts.append(off)
            off += (n + 31) // 32 * 32
        raw = torch.zeros(off + 32, dtype=torch.float32, device=device)
        skip = (-raw.data_ptr() % 128) // 4  # CPU allocations are only 64-byte aligned; CUDA ones give skip == 0
        self.flat = raw[skip:skip + off]
        self.views: Dict[str, torch.Tensor] = {}
        for (k, s), n, st in zip(self.shapes.items(), sizes, starts):
            self.views[k] = self.flat[st:st + n].view(*s)

    def zero_(self):
        self.flat.zero_()

    def accumulate(self, grads: Dict[str, torch.Tensor]):
        dst = [self.views[k] for k in GRAD_FIELDS if self.views[k].numel()]
        src = [grads[k].view_as(self.views[k]) for k in GRAD_FIELDS if self.views[k].numel()]
        torch._foreach_add_(dst, src)

    def allreduce(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat

    @property
    def nbytes(self):
        return self.flat.numel() * 4


def max_over_ranks(value: float, device) -> float:
    """Timing rule: multi-GPU numbers are the max over ranks."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
