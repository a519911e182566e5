"""K3 (persistent 1D rollout kernel, csrc/rollout1d.cu): the reference's KS DenseED rolled out in ONE launch, one warp per trajectory,
against (a) the layer-by-layer path of the same library (arpde_rollout_layerwise: nine launches per step), (b) the CPU oracle of
DenseED.forward + the AR loop of 1D-KS-SWAG/main.py:117-131, (c) the rollout recorded from the reference with its shipped
epoch-200 checkpoint (tests/golden/ks1d_ckpt200.npz, also covered by test_gpu_parity.py::test_rollout_shipped_checkpoint).
KS is chaotic: single steps are held to the north_star's 1e-5 relative, longer horizons are reported per step and bounded loosely."""
import numpy as np
import pytest
import torch

from _util import golden, rel_err, sub
import _models as M
from oracle import arpde_oracle as O

from arpde_b200 import engine

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def _ks_swag(K=4, seed=0):
    g, cfg, args, model, bayes, swag = M.build('ks1d', DEV, max_models=K)
    M.load_golden_weights(bayes, g)
    torch.manual_seed(seed)
    with torch.no_grad():
        base = [p.detach().clone() for p in bayes.parameters()]
        for _ in range(K + 1):
            for p, b in zip(bayes.parameters(), base):
                p.copy_(b + 0.01 * torch.randn_like(b) * (b.abs().mean() + 1e-3))
            swag.collect()
        for p, b in zip(bayes.parameters(), base):
            p.copy_(b)
    swag.eval()
    return cfg, bayes, swag


def _ics(n, seed=3):
    rng = np.random.RandomState(seed)
    x = np.linspace(0, 22 * np.pi, 97)[:-1]
    out = [(2.5 + rng.rand()) * np.sin(2 * np.pi * rng.randint(2, 5) * x / (22 * np.pi) + 6 * rng.rand()) + rng.randn() * np.cos(2 * np.pi * x / (22 * np.pi)) for _ in range(n)]
    return torch.tensor(np.stack(out)[:, None], dtype=torch.float32, device=DEV)


@pytest.mark.parametrize('S,N', [(1, 1), (3, 5), (7, 3), (30, 64), (2, 300)])
def test_persistent_rollout_equals_layerwise(S, N):
    """batches of consecutive trajectories that start mid-group, span two (or, for tiny groups, several) weight groups, fill CTAs
    partially (N*S not a multiple of the batch) and exceed one wave (2 x 300)"""
    cfg, bayes, swag = _ks_swag()
    gen = torch.Generator(device=DEV); gen.manual_seed(5)
    flat, names, numels = swag.sample_flat(S, generator=gen)
    named = list(bayes.named_parameters())
    ic = _ics(N)
    T = 12
    a = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named)
    b = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named, layerwise=True)
    assert a.shape == b.shape == (S * N, T + 1, 1, 96)
    errs = [rel_err(a[:, t], b[:, t]) for t in range(T + 1)]
    print('S=%d N=%d per-step rel err persistent vs layerwise:' % (S, N), ['%.1e' % e for e in errs])
    assert errs[0] == 0.0 and errs[1] < TOL
    assert max(errs) < 1e-3
    # without sampled weights (the module's own parameters, one group)
    a1 = engine.rollout(swag, ic, 3)
    b1 = engine.rollout(swag, ic, 3, layerwise=True)
    assert rel_err(a1[:, 1], b1[:, 1]) < TOL and rel_err(a1, b1) < 1e-4


def test_persistent_rollout_vs_oracle_per_sample():
    cfg, bayes, swag = _ks_swag()
    S, N, T = 3, 4, 6
    gen = torch.Generator(device=DEV); gen.manual_seed(9)
    flat, names, numels = swag.sample_flat(S, generator=gen)
    named = list(bayes.named_parameters())
    ic = _ics(N, seed=8)
    traj = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named)
    sd0 = {k: v.detach().cpu() for k, v in bayes.model.state_dict().items()}
    hist = ic.cpu().repeat(1, 2, 1)
    for s in range(S):
        sd = dict(sd0)
        off = 0
        for (n, p), ne in zip(named, numels):
            sd[n[len('model.'):]] = flat[s, off:off + ne].view(p.shape).cpu()
            off += ne
        step = lambda inp: O.densed_forward('ks1d', sd, inp, cfg['blocks'], training=False)
        ref = O.rollout('ks1d', step, hist, cfg['nic'], T)
        for t in range(T + 1):
            assert rel_err(traj[s * N:(s + 1) * N, t], ref[:, t]) < (TOL if t <= 1 else 2e-4), (s, t)


def test_persistent_rollout_shipped_checkpoint_long_horizon():
    """200 steps from the reference's epoch-200 weights stay finite and bounded (the trained surrogate is stable), first 20 steps match
    the recorded reference rollout"""
    gk = golden('ks1d_ckpt200')
    g, cfg, args, model, bayes, swag = M.build('ks1d', DEV, max_models=int(gk['max_models']))
    swag.load_state_dict(sub(gk, 'sd/'), strict=True)
    swag.eval()
    ic = torch.from_numpy(gk['ic']).to(DEV)[:, :1]
    traj = engine.rollout(swag, ic, 200)
    assert torch.isfinite(traj).all() and float(traj.abs().max()) < 10.0
    T = gk['rollout'].shape[1] - 1
    errs = [rel_err(traj[:, t], gk['rollout'][:, t]) for t in range(T + 1)]
    print('ks ckpt per-step rel err:', ['%.1e' % e for e in errs])
    assert errs[1] < TOL and max(errs) < 1e-3
