"""world_size-2 gloo test of the data-parallel host logic (sharding, global plan, gradient and
stats all-reduce, identical optimizer step) with a CPU stand-in for the CUDA engine whose "gradients"
are a known linear function of its rows, so the all-reduced result can be checked exactly."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icrl_b200.dp import DataParallelA2C, global_plan, shard_bounds


def test_shard_bounds_cover_rows():
    for n, w in ((4096, 8), (10, 4), (7, 8), (256, 1)):
        b = [shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


class _FakeEngine:
    def __init__(self, n):
        self.flat_grad = torch.zeros(n)
        self.param = torch.nn.Parameter(torch.zeros(n))
        self.param.grad = self.flat_grad

    def step(self, features, captions, uniforms=None, global_rows=None, level=None, plan=None, **kw):
        f = torch.as_tensor(features)
        self.flat_grad.copy_(f.sum(dim=0) / float(global_rows * plan[1]))
        return {"stats": torch.tensor([float(f.sum()), float(f.shape[0]), float(plan[1])]) / float(global_rows)}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    feats = rs.standard_normal((10, 16)).astype(np.float32)
    caps = np.full((10, 9), 5, dtype=np.int64)
    caps[:, 8] = 2
    caps[3, 4] = 2                      # an early <END> on rank 0's shard must not change the plan
    lo, hi = shard_bounds(10, rank, world)
    eng = _FakeEngine(16)
    opt = torch.optim.SGD([eng.param], lr=1.0)
    dp = DataParallelA2C(eng, opt)
    plan = global_plan(caps[lo:hi][:, :6] if rank == 0 else caps[lo:hi], None)
    res = dp.step(feats[lo:hi], caps[lo:hi], None, global_rows=10)
    first = (plan, eng.flat_grad.clone(), res["stats"].clone(), eng.param.detach().clone())
    # default path: global_rows=None must be resolved to the GLOBAL row count (sum over ranks), not the local shard
    eng2 = _FakeEngine(16)
    res2 = DataParallelA2C(eng2, None).step(feats[lo:hi], caps[lo:hi], None)
    out[rank] = first + (eng2.flat_grad.clone(), res2["stats"].clone())
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    rs = np.random.RandomState(0)
    feats = torch.from_numpy(rs.standard_normal((10, 16)).astype(np.float32))
    expect = feats.sum(dim=0) / (10 * 8)
    for r in (0, 1):
        plan, grad, stats, param, grad_default, stats_default = out[r]
        assert torch.allclose(grad_default, expect, atol=1e-6)  # global_rows=None: same normalisation as the explicit call
        assert torch.allclose(stats_default, stats, atol=1e-6)
        assert plan == (1, 8)                                   # caplen from the GLOBAL batch
        assert torch.allclose(grad, expect, atol=1e-6)
        assert torch.allclose(param, -expect, atol=1e-6)        # identical SGD step on both ranks
        assert abs(float(stats[0]) - float(feats.sum()) / 10) < 1e-5 and abs(float(stats[1]) - 1.0) < 1e-6
    assert torch.equal(out[0][3], out[1][3])
