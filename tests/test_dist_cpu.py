"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (SURVEY 8e parity check:
allreduced_grad(N ranks) == sum over the same views on 1 rank)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from gaustar_b200 import dist as gdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_view_sharding_is_a_partition():
    for V in (1, 7, 16, 160):
        for world in (1, 2, 4, 8):
            shards = [gdist.views_for_rank(V, r, world) for r in range(world)]
            flat = sorted(v for s in shards for v in s)
            assert flat == list(range(V))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_flat_grads_layout_single_process():
    fg = gdist.FlatGrads(P=10, M=4, device="cpu")
    assert fg.flat.numel() >= 10 * (3 + 3 + 4 + 1 + 12) and fg.nbytes == fg.flat.numel() * 4
    assert all(v.data_ptr() % 128 == 0 for v in fg.views.values())
    g = {"dL_dmeans3D": torch.ones(10, 3), "dL_dscales": 2 * torch.ones(10, 3), "dL_drotations": 3 * torch.ones(10, 4),
         "dL_dopacity": 4 * torch.ones(10, 1), "dL_dsh": 5 * torch.ones(10, 4, 3)}
    fg.accumulate(g); fg.accumulate(g)
    assert float(fg.views["dL_dsh"].min()) == 10.0 and float(fg.views["dL_dmeans3D"].max()) == 2.0
    assert float(fg.flat.sum()) == 2 * (30 * 1 + 30 * 2 + 40 * 3 + 10 * 4 + 120 * 5)
    fg.allreduce()  # no process group: no-op
    fg.zero_()
    assert float(fg.flat.abs().sum()) == 0.0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _view_grads(v, P, M):
    gen = torch.Generator().manual_seed(100 + v)
    return {"dL_dmeans3D": torch.randn(P, 3, generator=gen), "dL_dscales": torch.randn(P, 3, generator=gen), "dL_drotations": torch.randn(P, 4, generator=gen),
            "dL_dopacity": torch.randn(P, 1, generator=gen), "dL_dsh": torch.randn(P, M, 3, generator=gen)}


def _worker(rank, world, port, V, P, M, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = gdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    fg = gdist.FlatGrads(P, M, "cpu")
    for v in gdist.views_for_rank(V, rank, world):
        fg.accumulate(_view_grads(v, P, M))
    fg.allreduce()
    mx = gdist.max_over_ranks(float(rank + 1), "cpu")
    sm = gdist.sum_over_ranks(1.0, "cpu")
    gdist.barrier()
    if rank == 0:
        torch.save({"flat": fg.flat.clone(), "max": mx, "sum": sm}, out)
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_equals_single_rank_sum(tmp_path):
    V, P, M, world = 7, 50, 4, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), V, P, M, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = gdist.FlatGrads(P, M, "cpu")
    for v in range(V):
        ref.accumulate(_view_grads(v, P, M))
    assert torch.allclose(got["flat"], ref.flat, rtol=1e-6, atol=1e-6)
    assert got["max"] == 2.0 and got["sum"] == 2.0


# ---- TrainerSharding: the hooks that shard an UNMODIFIED one-view-per-iteration loop (refine.py:529-548, :794-795) ----
def test_rank_view_sequences_are_disjoint_per_iteration_and_cover_the_permutation():
    for n in (2, 5, 8, 21, 160):
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(n))
        for world in (1, 2, 4, 8):
            if world > n:
                continue
            seqs = [gdist.rank_view_sequence(perm, r, world) for r in range(world)]
            assert all(s.numel() == n for s in seqs)
            for i in range(n):
                held = [int(s[i]) for s in seqs]
                assert len(set(held)) == world  # different views on the ranks at every iteration
                assert held == [int(perm[(i * world + r) % n]) for r in range(world)]  # = the next `world` entries of the shared order


def _toy_targets(n):
    g = torch.Generator().manual_seed(7)
    return torch.randn(n, 5, 3, generator=g), torch.randn(n, 5, generator=g)


def _toy_loss(a, b, c, cam, ta, tb):
    loss = ((a - ta[cam]) ** 2).mean() + (torch.tanh(b) - tb[cam]).abs().mean()
    if cam % 3 == 0:  # a parameter only some views touch: its grad is None otherwise (set_to_none=True)
        loss = loss + (c * float(cam + 1)).sum()
    return loss


def _unmodified_loop(n_cams, passes, lr=0.05):
    """Shaped like refine.py: randperm per pass, ONE view per iteration, optimizer.step(); zero_grad(set_to_none=True).  Knows nothing of ranks."""
    ta, tb = _toy_targets(n_cams)
    a, b, c = [torch.nn.Parameter(torch.zeros(s)) for s in ((5, 3), (5,), (2,))]
    opt = torch.optim.Adam([{"params": [a], "lr": lr}, {"params": [b, c], "lr": lr / 2}])
    seen = []
    for _ in range(passes):
        shuffled_idx = torch.randperm(n_cams)
        for i in range(0, len(shuffled_idx), 1):
            cmr_i = shuffled_idx[i:i + 1].item()
            seen.append(cmr_i)
            _toy_loss(a, b, c, cmr_i, ta, tb).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    return [p.detach().clone() for p in (a, b, c)], seen


def _sharded_worker(rank, world, port, n_cams, passes, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    gdist.init_from_env(backend="gloo")
    torch.manual_seed(1234 + rank)  # the trainer's global RNG differs per rank; the shared permutation must not depend on it
    with gdist.TrainerSharding(shared_seed=3) as ts:
        params, seen = _unmodified_loop(n_cams, passes)
    assert ts.steps == n_cams * passes and ts.bytes_last == 4 * (15 + 5 + 2 + 3)
    assert torch.randperm.__module__ != gdist.__name__  # restored on exit
    torch.save({"params": params, "seen": seen}, out + f".{rank}")
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(180)
def test_unmodified_loop_sharded_over_two_ranks_equals_one_rank_stepping_on_both_views(tmp_path):
    n_cams, passes, world = 7, 2, 2
    out = str(tmp_path / "r")
    mp.spawn(_sharded_worker, args=(world, _free_port(), n_cams, passes, out), nprocs=world, join=True)
    got = [torch.load(out + f".{r}") for r in range(world)]
    # what one process does with the SAME shared order, two views per step, gradients summed
    gen = torch.Generator().manual_seed(3)
    ta, tb = _toy_targets(n_cams)
    a, b, c = [torch.nn.Parameter(torch.zeros(s)) for s in ((5, 3), (5,), (2,))]
    opt = torch.optim.Adam([{"params": [a], "lr": 0.05}, {"params": [b, c], "lr": 0.025}])
    want_seen = [[], []]
    for _ in range(passes):
        perm = torch.randperm(n_cams, generator=gen)
        for i in range(n_cams):
            for r in range(world):
                cam = int(perm[(i * world + r) % n_cams])
                want_seen[r].append(cam)
                _toy_loss(a, b, c, cam, ta, tb).backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
    for r in range(world):
        assert got[r]["seen"] == want_seen[r]
        for p_got, p_want in zip(got[r]["params"], (a, b, c)):
            assert torch.allclose(p_got, p_want.detach(), rtol=1e-5, atol=1e-6)
    for p0, p1 in zip(got[0]["params"], got[1]["params"]):
        assert torch.equal(p0, p1)  # the replicas never drift apart: identical reduced gradients, identical Adam state


def test_trainer_sharding_is_a_no_op_in_a_single_process():
    before = torch.randperm
    with gdist.TrainerSharding() as ts:
        assert torch.randperm is before and ts.world == 1
        params, seen = _unmodified_loop(4, 1)
    assert sorted(seen) == [0, 1, 2, 3] and ts.steps == 0


def test_bench_camera_windows_cover_the_pool_evenly():
    """bench.py: over a step every camera of the pool is rendered equally often at 8 GPUs (160 views / 32 cameras), every rank walks 20
    consecutive cameras, and one GPU renders the cameras it always did."""
    import collections
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    V, C = bench.VIEWS_PER_GPU, bench.CAM_POOL
    for step in (0, 1, 7):
        assert bench.cameras_of_step(step, 0, 1) == [(step * V + v) % C for v in range(V)]
        cnt = collections.Counter()
        for r in range(8):
            cams = bench.cameras_of_step(step, r, 8)
            assert len(cams) == V and all((b - a) % C == 1 for a, b in zip(cams, cams[1:]))  # a window of consecutive cameras
            cnt.update(cams)
        assert set(cnt.values()) == {V * 8 // C} and len(cnt) == C
        for world in (2, 4):
            per_rank = [bench.cameras_of_step(step, r, world) for r in range(world)]
            assert all(len(c) == V for c in per_rank)
            spread = collections.Counter(c for cams in per_rank for c in cams)
            assert max(spread.values()) - min(spread.values()) <= 1 and len(spread) == C
