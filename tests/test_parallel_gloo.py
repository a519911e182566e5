"""CPU, world_size 2, gloo: the host-side sharding / gathering / gradient-bucket / BatchNorm-sync logic of the N>1 path."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import ROOT


def _fake_run_local(ic, tsteps, srange, n_total=3):
    """Stand-in for engine.swag_rollout: sample s adds s to every state; returns the moments over samples [lo, hi)."""
    lo, hi = srange
    ns = hi - lo
    s = torch.arange(lo, hi, dtype=torch.float32).repeat_interleave(ic.shape[0]).view(-1, 1, 1, 1)
    traj = ic[:, None].repeat(ns, tsteps + 1, 1, 1) + s
    tr = traj.view(ns, ic.shape[0], tsteps + 1, *ic.shape[1:])
    return traj, tr.mean(0), tr.var(0, unbiased=False), torch.full((n_total,), 4.0)


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from arpde_b200 import parallel as PL
        # --- shard ranges partition the items
        for n in (1, 2, 7, 64):
            spans = [PL.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        # --- IC-sharded rollout: each rank processes its own ICs, everybody ends with all means/vars
        n_ic, T = 7, 3
        ics = torch.arange(n_ic, dtype=torch.float32).view(n_ic, 1, 1).repeat(1, 2, 4)
        traj, mean, var, betas = PL.sharded_swag_rollout(_fake_run_local, ics, T, 3)
        lo, hi = PL.shard_range(n_ic, rank, ws)
        assert traj.shape[0] == 3 * (hi - lo)
        assert mean.shape == (n_ic, T + 1, 2, 4) and var.shape == mean.shape
        assert torch.allclose(mean[:, 0, 0, 0], torch.arange(n_ic, dtype=torch.float32) + 1.0)
        assert torch.allclose(var, torch.full_like(var, 2.0 / 3.0 + 0.25))          # sample variance + 1/mean(beta)
        # gather=False keeps the local shard
        _, mloc, _, _ = PL.sharded_swag_rollout(_fake_run_local, ics, T, 3, gather=False)
        assert mloc.shape[0] == hi - lo
        # --- fewer ICs than ranks, IC mode forced: the rank with the empty shard must not hang the gather
        one = ics[:1]
        traj1, mean1, var1, betas1 = PL.sharded_swag_rollout(_fake_run_local, one, T, 3, mode='ic')
        assert (traj1 is None) == (rank == 1)
        assert mean1.shape == (1, T + 1, 2, 4) and torch.allclose(mean1, torch.full_like(mean1, 1.0))
        assert torch.allclose(var1, torch.full_like(var1, 2.0 / 3.0 + 0.25)) and torch.allclose(betas1, torch.full((3,), 4.0))
        # --- fewer ICs than ranks, auto -> samples are sharded and the partial moments merged (Chan)
        traj2, mean2, var2, _ = PL.sharded_swag_rollout(_fake_run_local, one, T, 3)
        slo, shi = PL.shard_range(3, rank, ws)
        assert traj2.shape[0] == (shi - slo) * 1
        assert torch.allclose(mean2, torch.full_like(mean2, 1.0)) and torch.allclose(var2, torch.full_like(var2, 2.0 / 3.0 + 0.25), atol=1e-6)
        # 1D variance formula: mean_s(1/beta_s)
        _, _, var3, _ = PL.sharded_swag_rollout(_fake_run_local, one, T, 3, formula='1d')
        assert torch.allclose(var3, torch.full_like(var3, 2.0 / 3.0 + 0.25), atol=1e-6)
        # --- flat gradient bucket as the gradient storage: p.grad are views, one all-reduce, AVERAGE by default
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
        bucket = PL.FlatGradBucket(ps)
        assert all(p.grad is not None and p.grad.data_ptr() == bucket.flat.data_ptr() + 4 * bucket.offset_of(p) for p in ps)
        loss = sum(((rank + 1.0) * (i + 1) * p).sum() for i, p in enumerate(ps))
        loss.backward()                                       # autograd accumulates into the views
        bucket.all_reduce_()
        assert torch.allclose(ps[0].grad, torch.full((3, 2), 1.5)) and torch.allclose(ps[1].grad, torch.full((5,), 3.0))
        assert torch.allclose(bucket.flat[-1:], torch.full((1,), 4.5))
        bucket.zero_()
        assert float(ps[0].grad.abs().sum()) == 0.0 and ps[0].grad.data_ptr() == bucket.flat.data_ptr()
        # --- detached gradients (plain autograd + zero_grad(set_to_none=True)): packed path; SUM on request;
        #     a parameter NO rank differentiated keeps grad = None, one that SOME rank differentiated gets the sum
        qs = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1)),
              torch.nn.Parameter(torch.zeros(2))]
        b2 = PL.FlatGradBucket(qs, attach=False)
        for i, p in enumerate(qs[:3]):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        if rank == 1:
            qs[2].grad = None
        b2.all_reduce_(average=False)
        assert torch.allclose(qs[0].grad, torch.full((3, 2), 3.0)) and torch.allclose(qs[1].grad, torch.full((5,), 6.0))
        assert torch.allclose(qs[2].grad, torch.full((1,), 3.0)) and qs[3].grad is None
        # --- BatchNorm sync hook: the callback all-reduces exactly the slice the library points at
        from arpde_b200 import _lib as L
        from arpde_b200 import program as P
        with PL.sync_batchnorm() as sync:
            assert P.BN_SYNC is sync and sync.world_size() == ws
            stats = torch.arange(10, dtype=torch.float64) * (rank + 1)
            cb = L.allreduce_callback(sync, stats)
            assert cb(0, stats.data_ptr() + 8 * 2, 4) == 0
            assert torch.allclose(stats[2:6], torch.arange(2, 6, dtype=torch.float64) * 3) and float(stats[6]) == 6.0 * (rank + 1)
            assert cb(0, stats.data_ptr() + 8 * 8, 4) == 1           # outside the registered scratch: refused
            assert sync.calls == 1 and sync.doubles == 4
        assert P.BN_SYNC is None
        q.put((rank, 'ok'))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, 'ok'), (1, 'ok')], res
