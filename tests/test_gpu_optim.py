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
