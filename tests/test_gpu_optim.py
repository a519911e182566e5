"""GPU parity of the one-launch Adam (csrc/adam.cu, arpde_b200.optim.Adam) against torch.optim.Adam on the reference's
optimizer configuration (two parameter groups, ExponentialLR), including state_dict interchange.  fp32 update arithmetic with
a different instruction order than torch's multi-tensor path: tolerance 1e-6 relative after 25 steps."""
import copy

import pytest
import torch

from _util import rel_err
from arpde_b200.optim import Adam

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def make_params(seed):
    torch.manual_seed(seed)
    shapes = [(1,), (96, 6, 5, 5), (96,), (48, 96, 3, 3), (4, 52, 3, 3), (2, 16, 5, 5), (1300,)]
    return [torch.nn.Parameter(torch.randn(s, device=DEV) * 0.1) for s in shapes]


def run(opt_cls, steps, wd=0.0, swap_at=None):
    ps = make_params(0)
    groups = [{'params': [ps[0]], 'lr': 1e-2}, {'params': ps[1:]}]
    opt = opt_cls(groups, lr=1e-3, weight_decay=wd)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9)
    gen = torch.Generator(device=DEV).manual_seed(1)
    for it in range(steps):
        for i, p in enumerate(ps):
            p.grad = None if (i == 3 and it % 4 == 0) else torch.randn(p.shape, device=DEV, generator=gen) * (1.0 + i)     # a tensor without grad now and then
        opt.step(); opt.zero_grad()
        if it % 5 == 4:
            sched.step()
        if swap_at is not None and it == swap_at:          # continue with the other implementation from a checkpoint
            sd = copy.deepcopy(opt.state_dict())
            other = Adam if opt_cls is torch.optim.Adam else torch.optim.Adam
            opt = other(groups, lr=1e-3, weight_decay=wd)
            opt.load_state_dict(sd)
            sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.9, last_epoch=sched.last_epoch)
    return ps, opt


@pytest.mark.parametrize('wd', [0.0, 0.01])
def test_adam_matches_torch(wd):
    ref, ropt = run(torch.optim.Adam, 25, wd)
    got, gopt = run(Adam, 25, wd)
    for a, b in zip(got, ref):
        assert rel_err(a.detach(), b.detach().double()) < 1e-6
    for a, b in zip(got, ref):
        sa, sb = gopt.state[a], ropt.state[b]
        assert float(sa['step']) == float(sb['step'])
        assert rel_err(sa['exp_avg'], sb['exp_avg'].double()) < 1e-6 and rel_err(sa['exp_avg_sq'], sb['exp_avg_sq'].double()) < 1e-6


def test_adam_checkpoint_interchange_with_torch():
    ref, _ = run(torch.optim.Adam, 20)
    a, _ = run(Adam, 20, swap_at=9)                 # ours for 10 steps, then torch from our state_dict
    b, _ = run(torch.optim.Adam, 20, swap_at=9)     # torch for 10 steps, then ours from torch's state_dict
    for x, y, z in zip(a, b, ref):
        assert rel_err(x.detach(), z.detach().double()) < 1e-6 and rel_err(y.detach(), z.detach().double()) < 1e-6


def test_adam_rejects_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Adam([p]).step()


def test_graphed_training_window_matches_eager():
    """engine.GraphedStep: the BPTT window (forward with batch-statistics BatchNorm, Crank-Nicolson, negative log joint, backward into
    the flat gradient bucket) captured in a CUDA graph replays to the same loss and gradients as the eager calls, also for new inputs."""
    import types
    import numpy as np
    from arpde_b200 import engine, parallel as PL
    from arpde_b200.nn.densed import DenseED2d
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.integrators import Burger2DIntegrate
    torch.manual_seed(0); np.random.seed(0)
    a = types.SimpleNamespace(dt=0.005, device=torch.device(DEV), nic=3)
    model = DenseED2d(6, 2, [4], growth_rate=4, init_features=48).to(DEV)
    bayes = BayesNN(a, model)
    integ = Burger2DIntegrate(1.0 / 32, nu=0.005, grad_kernels=[3, 3], device=DEV)
    bayes.train()
    bucket = PL.FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))

    def window(ic):
        inp = ic.repeat(1, 3, 1, 1)
        loss = 0
        bucket.zero_()
        for _ in range(2):
            u = bayes(inp[:, -6:])
            ustar = integ.crankNicolson(u, inp[:, -2:], 0.003)
            loss = loss + bayes.calc_neg_log_joint(u, ustar, 10)
            inp = torch.cat([inp, u], 1)
        with bucket.capture():
            loss.backward()
        return loss.detach()

    ic0, ic1 = torch.randn(6, 2, 32, 32, device=DEV), torch.randn(6, 2, 32, 32, device=DEV)
    # running statistics advance with every call: compare loss and gradients (batch statistics), which do not depend on them
    ref = []
    for ic in (ic0, ic1):
        l = window(ic); ref.append((l.clone(), bucket.flat.clone()))
    gstep = engine.GraphedStep(window, (ic0.clone(),))
    for ic, (l_ref, g_ref) in zip((ic0, ic1), ref):
        l = gstep(ic)
        torch.cuda.synchronize()
        assert rel_err(l, l_ref) < 1e-6, (l, l_ref)                  # same launches; fp64 atomics may reorder
        assert rel_err(bucket.flat, g_ref) < 1e-6
    with pytest.raises(ValueError):
        gstep(torch.randn(5, 2, 32, 32, device=DEV))
