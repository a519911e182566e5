"""Worker of tests/test_gpu_dp.py (T12): one data-parallel training step on `ws` ranks vs the same step on one rank at the global batch."""
import contextlib
import io
import os
import sys
import types


def build(dev, seed=0):
    import numpy as np
    import torch
    from arpde_b200.nn.densed import DenseED2d
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    from arpde_b200.nn.integrators import Burger2DIntegrate
    from arpde_b200.optim import Adam
    torch.manual_seed(seed); np.random.seed(seed)
    a = types.SimpleNamespace(dt=0.005, device=dev, nic=3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(dev)
        bayes = BayesNN(a, model)
        swag = SwagNN(a, bayes, full_cov=True, max_models=2)
        integ = Burger2DIntegrate(1.0 / 32, nu=0.005, grad_kernels=[3, 3], device=dev)
    opt = Adam([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
    swag.train()
    return swag, bayes, integ, opt


def train_step(swag, bayes, integ, opt, ic, dp, capture=True, tback=2, ntrain=5, peer='auto'):
    """2D-Burgers-SWAG/main.py:61-91 for one BPTT window; dp: use the data-parallel hooks (global BatchNorm + gradient all-reduce)."""
    import torch
    from arpde_b200 import parallel as PL
    bucket = PL.FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))
    bucket.zero_()
    inp = ic.repeat(1, 3, 1, 1)
    loss = 0
    ctx = PL.sync_batchnorm(peer=peer) if dp else contextlib.nullcontext()
    with ctx:
        for _ in range(tback):
            u_pred = swag(inp[:, -6:])
            ustar = integ.crankNicolson(u_pred, inp[:, -2:], 0.003)
            loss = loss + swag.calc_neg_log_joint(u_pred, ustar, ntrain)
            inp = torch.cat([inp, u_pred], 1)
        with (bucket.capture() if capture else contextlib.nullcontext()):
            loss.backward()
    if dp:
        bucket.all_reduce_()
    grads = bucket.flat.clone()
    opt.step()
    stats = {k: v.detach().clone() for k, v in bayes.state_dict().items() if 'running' in k or 'num_batches' in k}
    weights = torch.cat([p.detach().reshape(-1) for p in bayes.parameters()])
    return loss.detach().clone(), grads, stats, weights


def worker(rank, ws, port, backend, root, q):
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    ngpu = torch.cuda.device_count()
    dev = torch.device('cuda', rank % ngpu)
    torch.cuda.set_device(dev)
    if backend == 'nccl':
        dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=dev)
    else:
        dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from oracle import arpde_oracle as O
        from _util import rel_err
        B = 8
        ics = O.ic_burgers2d(32, range(B)).to(dev)
        per = B // ws
        # data parallel: every rank the same initial weights, its own shard of the mini-batch
        swag, bayes, integ, opt = build(dev)
        loss_dp, g_dp, st_dp, w_dp = train_step(swag, bayes, integ, opt, ics[rank * per:(rank + 1) * per], dp=True)
        lsum = loss_dp.clone()
        dist.all_reduce(lsum)
        loss_dp_mean = lsum / ws
        # single process, global batch, no hooks (plain autograd accumulation into the bucket views)
        swag1, bayes1, integ1, opt1 = build(dev)
        loss_1, g_1, st_1, w_1 = train_step(swag1, bayes1, integ1, opt1, ics, dp=False, capture=False)
        errs = {'loss': rel_err(loss_dp_mean, loss_1), 'grads': rel_err(g_dp, g_1), 'weights': rel_err(w_dp, w_1),
                'stats': max(rel_err(st_dp[k], st_1[k]) for k in st_1)}
        # every rank ends with identical weights and running statistics
        wcopy = w_dp.clone()
        dist.broadcast(wcopy, src=0)
        errs['rank_consistency'] = float((wcopy - w_dp).abs().max())
        # the NVLink peer-memory all-reduce of the BatchNorm sums (csrc/peer.cu; used by default when every rank has its own GPU) against
        # torch.distributed.all_reduce from the Python callback: two ranks add the same two numbers, so the step must be bit-identical
        from arpde_b200 import parallel as PL
        errs['peer_used'] = PL._peer_allreduce(None) is not None
        if errs['peer_used']:
            swag2, bayes2, integ2, opt2 = build(dev)
            loss_cb, g_cb, st_cb, w_cb = train_step(swag2, bayes2, integ2, opt2, ics[rank * per:(rank + 1) * per], dp=True, peer=False)
            errs['peer_vs_callback'] = max(float((g_cb - g_dp).abs().max()), float((w_cb - w_dp).abs().max()), float((loss_cb - loss_dp).abs()))
            errs['peer_failed'] = PL._peer_allreduce(None).failed()
        q.put((rank, errs))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def infer_worker(rank, ws, port, backend, root, q):
    """SWAG inference sharded over `ws` ranks (by initial condition and by weight sample) vs the single-process result."""
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    ngpu = torch.cuda.device_count()
    dev = torch.device('cuda', rank % ngpu)
    torch.cuda.set_device(dev)
    if backend == 'nccl':
        dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=dev)
    else:
        dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        from oracle import arpde_oracle as O
        from _util import rel_err, golden, sub
        import _models as M
        from arpde_b200 import engine
        from arpde_b200 import parallel as PL
        with contextlib.redirect_stdout(io.StringIO()):
            g, cfg, args, model, bayes, swag = M.build('burgers2d', str(dev), max_models=int(golden('burgers2d')['swag/max_models']))
        M.load_golden_weights(bayes, g)
        swag.load_state_dict({k: v for k, v in sub(g, 'swag/sd/').items()}, strict=True)
        swag.eval()
        S, T = 5, 3
        errs = {}
        for tag, n_ic, mode in (('ic_shard', 5, 'ic'), ('sample_shard', 1, 'auto'), ('ic_shard_one_ic', 1, 'ic')):
            ics = O.ic_burgers2d(32, range(n_ic)).to(dev)
            gen = torch.Generator(device=dev); gen.manual_seed(11)
            _, m1, v1, b1 = engine.swag_rollout(swag, ics, T, S, generator=gen)                 # this rank alone, everything
            traj, m, v, b = PL.swag_rollout(swag, ics, T, S, seed=11, mode=mode)
            errs[tag] = (rel_err(m, m1), rel_err(v, v1), rel_err(b, b1), None if traj is None else int(traj.shape[0]))
        q.put((rank, errs))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def peer_worker(rank, ws, port, backend, root, q):
    """parallel.PeerMoments (NCCL + symmetric memory, one GPU per rank): moments stored straight into rank 0's buffer == gather of local moments."""
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=ws, device_id=dev)
    try:
        from _util import rel_err
        from arpde_b200 import engine
        from arpde_b200 import parallel as PL
        assert PL.PeerMoments.available()
        S, n_loc, inner_shape = 5, 3, (4, 2, 8, 8)
        torch.manual_seed(100 + rank)
        traj = torch.randn((S * n_loc,) + inner_shape, device=dev)
        pm = PL.PeerMoments(ws * n_loc, inner_shape, dev)
        pm.reduce_into(traj, S, rank * n_loc, noise=0.25)
        pm.finish()
        mean, var = engine.predictive_moments(traj, S, noise=0.25)
        gm, gv = PL.gather_shards(mean, ws * n_loc), PL.gather_shards(var, ws * n_loc)
        torch.cuda.synchronize()
        errs = {}
        if rank == 0:
            errs = {'mean': float((pm.mean - gm).abs().max()), 'var': float((pm.var - gv).abs().max())}
        dist.barrier()
        q.put((rank, errs))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()
