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
