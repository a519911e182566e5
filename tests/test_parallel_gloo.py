"""CPU, world_size 2, gloo: the host-side sharding / gathering / gradient-bucket logic of the N>1 path."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import ROOT


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
        # --- sharded rollout: each rank processes its own ICs, everybody ends with all means/vars
        n_ic, T = 7, 3
        ics = torch.arange(n_ic, dtype=torch.float32).view(n_ic, 1, 1).repeat(1, 2, 4)

        def run_local(ic, tsteps, n_samples):
            s = torch.arange(n_samples, dtype=torch.float32).repeat_interleave(ic.shape[0]).view(-1, 1, 1, 1)
            traj = ic[:, None].repeat(n_samples, tsteps + 1, 1, 1) + s
            tr = traj.view(n_samples, ic.shape[0], tsteps + 1, 2, 4)
            return traj, tr.mean(0), tr.var(0, unbiased=False), torch.ones(n_samples)

        traj, mean, var, betas = PL.sharded_swag_rollout(run_local, ics, T, 3)
        lo, hi = PL.shard_range(n_ic, rank, ws)
        assert traj.shape[0] == 3 * (hi - lo)
        assert mean.shape == (n_ic, T + 1, 2, 4) and var.shape == mean.shape
        assert torch.allclose(mean[:, 0, 0, 0], torch.arange(n_ic, dtype=torch.float32) + 1.0)
        assert torch.allclose(var, torch.full_like(var, 2.0 / 3.0))
        # --- flat gradient bucket == sum of per-rank gradients
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(1))]
        for i, p in enumerate(ps):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        if rank == 1:
            ps[2].grad = None                                    # a rank without a gradient contributes zero
        PL.FlatGradBucket(ps).all_reduce_()
        assert torch.allclose(ps[0].grad, torch.full((3, 2), 3.0)) and torch.allclose(ps[1].grad, torch.full((5,), 6.0))
        assert torch.allclose(ps[2].grad, torch.full((1,), 3.0))
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
