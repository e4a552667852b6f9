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


# ---- hooks for an UNMODIFIED single-GPU trainer ("optimizer-pre-step hook + per-rank seed") -------------------------------
# gaustar_trainers/refine.py draws `torch.randperm(len(training_cameras))` per pass over the cameras (:534), renders ONE view per
# iteration (:546-548) and ends the iteration with `optimizer.step(); optimizer.zero_grad(set_to_none=True)` (:794-795; the SuGaR
# optimizer wrapper forwards to a torch.optim.Adam over the groups of sugar_optimizer.py:67-87).  Launched once per GPU, the same loop
# becomes view-sharded data parallelism without an edit when (1) at every iteration the ranks hold DIFFERENT views and (2) the ranks'
# gradients are summed before Adam runs, so that every rank takes the identical step on its replica.

def rank_view_sequence(perm: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r's camera order from the permutation all ranks share: entry i is perm[(i*world + r) mod n], so at iteration i the ranks
    hold the `world` consecutive entries perm[i*world .. i*world+world-1] (mod n) -- pairwise different whenever world <= n -- and
    the job as a whole walks the permutation in order, `world` views per step."""
    n = perm.numel()
    if n == 0 or world <= 1:
        return perm
    return perm[(torch.arange(n, device=perm.device) * world + rank) % n]


class TrainerSharding:
    """Installs the two hooks; remove() (or leaving the `with` block) restores torch.randperm and unregisters the optimizer hook.

        with gaustar_b200.dist.TrainerSharding():      # after init_from_env(); a no-op in a single process
            refined_training(args)                     # the reference's loop, unchanged

    * torch.randperm(n) returns rank_view_sequence(shared permutation): the shared permutation comes from a private CPU generator
      seeded alike on every rank, so it does not depend on (or disturb) the global RNG streams the trainer uses elsewhere.  Calls that
      pass their own `generator=` or `out=` are left alone.
    * a global optimizer-step pre-hook (torch.optim.optimizer.register_optimizer_step_pre_hook) sums the `.grad` of every parameter of
      the stepping optimizer over the ranks with ONE allreduce of one flat fp32 buffer (plus one presence counter per parameter, so a
      parameter no rank produced a gradient for keeps grad None and Adam skips it exactly as in the single-process run).
      `average=True` divides by the world size (mean over the step's views instead of their sum).

    Not covered: anything else the trainer derives from its local views only -- the densifier's view-space gradient statistics
    (refine.py:769-787) would let the ranks' Gaussian sets diverge, so run sharded only where densification is off (GauSTAR's per-frame
    refinement after the first frame) or reduce those statistics as well."""

    def __init__(self, average: bool = False, shared_seed: int = 0, group=None):
        self.average, self.group = average, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._gen = torch.Generator().manual_seed(int(shared_seed))
        self._orig_randperm = None
        self._hook = None
        self.steps = 0
        self.bytes_last = 0

    # -- (1) view selection
    def _randperm(self, n, *args, **kw):
        if kw.get("generator") is not None or kw.get("out") is not None or args:
            return self._orig_randperm(n, *args, **kw)
        dev, dt = kw.pop("device", None), kw.pop("dtype", torch.int64)
        kw.pop("requires_grad", None); kw.pop("pin_memory", None); kw.pop("layout", None)
        seq = rank_view_sequence(self._orig_randperm(int(n), generator=self._gen), self.rank, self.world).to(dt)
        return seq.to(dev) if dev is not None else seq

    # -- (2) gradient sum before the step
    def _pre_step(self, optimizer, args, kwargs):
        params = [p for g in optimizer.param_groups for p in g["params"] if p.requires_grad]
        if not params:
            return
        dev = params[0].device
        n = sum(p.numel() for p in params)
        flat = torch.zeros(n + len(params), dtype=torch.float32, device=dev)
        off, have = 0, []
        for p in params:
            if p.grad is not None:
                flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
                have.append(1.0)
            else:
                have.append(0.0)
            off += p.numel()
        flat[n:] = torch.tensor(have, dtype=torch.float32).to(dev)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        present = flat[n:].cpu()
        if self.average:
            flat[:n].div_(self.world)
        off = 0
        for k, p in enumerate(params):
            if present[k] > 0:
                g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
            off += p.numel()
        self.steps += 1
        self.bytes_last = flat.numel() * 4

    def install(self):
        if self.world <= 1 or self._hook is not None:
            return self
        from torch.optim.optimizer import register_optimizer_step_pre_hook
        self._orig_randperm = torch.randperm
        torch.randperm = self._randperm
        self._hook = register_optimizer_step_pre_hook(self._pre_step)
        return self

    def remove(self):
        if self._hook is not None:
            self._hook.remove()
            self._hook = None
        if self._orig_randperm is not None:
            torch.randperm = self._orig_randperm
            self._orig_randperm = None

    __enter__ = install

    def __exit__(self, *exc):
        self.remove()
        return False
