"""CPU: pin oracle/arpde_oracle.py against fixtures generated from the real reference
(tests/golden/make_golden.py).  fp32 against fp32, same torch build => tight tolerances."""
import numpy as np
import pytest
import torch

from _util import SYSTEMS, cfg_of, golden, rel_err, sub
from oracle import arpde_oracle as O

TOL = 2e-6      # oracle vs real reference, fp32 both sides


def _model_sd(g, prefix='sd/'):
    return sub(g, prefix, strip='model.')


def _x(g, cfg):
    return torch.from_numpy(g['hist'])[:, -cfg['in_channels']:]


@pytest.mark.parametrize('system', SYSTEMS)
def test_forward_eval(system):
    g = golden(system); cfg = cfg_of(g)
    y = O.densed_forward(system, _model_sd(g), _x(g, cfg), cfg['blocks'], training=False)
    assert rel_err(y, g['y_eval']) < TOL


@pytest.mark.parametrize('system', SYSTEMS)
def test_forward_train_and_running_stats(system):
    g = golden(system); cfg = cfg_of(g)
    sd = _model_sd(g)
    y = O.densed_forward(system, sd, _x(g, cfg), cfg['blocks'], training=True)
    assert rel_err(y, g['y_train0']) < TOL
    for k, v in sub(g, 'sd_after1/', strip='model.').items():
        assert rel_err(sd[k], v) < TOL, k


def _integ(system, g):
    dx, nu, dt = float(g['phys/dx']), float(g['phys/nu']), float(g['phys/dt'])
    gk = [int(v) for v in g['phys/grad_kernels']]
    if system == 'burgers2d':
        return (lambda a, b: O.burgers2d_cn(a, b, dt, dx, nu)), (lambda a, b: O.burgers2d_be(a, b, dt, dx, nu))
    if system == 'burgers1d':
        return (lambda a, b: O.burgers1d_cn(a, b, dt, dx, nu, *gk)), (lambda a, b: O.burgers1d_be(a, b, dt, dx, nu, *gk))
    return (lambda a, b: O.ks_cn(a, b, dt, dx, nu, *gk)), (lambda a, b: O.ks_be(a, b, dt, dx, nu, *gk))


@pytest.mark.parametrize('system', SYSTEMS)
@pytest.mark.parametrize('meth', ['crankNicolson', 'backwardEuler'])
def test_integrators_fwd_bwd(system, meth):
    g = golden(system)
    cn, be = _integ(system, g)
    fn = cn if meth == 'crankNicolson' else be
    a = torch.from_numpy(g['integ/uPred']).requires_grad_(True)
    b = torch.from_numpy(g['integ/uPred0']).requires_grad_(True)
    us = fn(a, b)
    assert rel_err(us, g['integ/%s/ustar' % meth]) < TOL
    (us * torch.from_numpy(g['integ/gout'])).sum().backward()
    assert rel_err(a.grad, g['integ/%s/g_uPred' % meth]) < 5e-6
    assert rel_err(b.grad, g['integ/%s/g_uPred0' % meth]) < 5e-6


@pytest.mark.parametrize('system', ['burgers1d', 'ks1d'])
def test_filters_all_orders(system):
    g = golden(system)
    dx = float(g['phys/dx'])
    a = torch.from_numpy(g['integ/uPred'])
    tabs = {'grad': (O.GRAD_TAPS, dx), 'lap': (O.LAP_TAPS, dx ** 2), 'd4': (O.D4_TAPS, dx ** 4)}
    n = 0
    for key in g:
        if key.startswith('filt/'):
            nm = next(t for t in ('grad', 'lap', 'd4') if key[5:].startswith(t)); k = int(key[5 + len(nm):])
            tab, sc = tabs[nm]
            assert rel_err(O.filt1d(a, tab[k], sc), g[key]) < TOL, key
            n += 1
    assert n >= 6


def test_filters_2d():
    g = golden('burgers2d'); dx = float(g['phys/dx'])
    a = torch.from_numpy(g['integ/uPred'])[:, :1]
    assert rel_err(O.filt2d(a, O.SOBEL_H, dx), g['filt/grad_h']) < TOL
    assert rel_err(O.filt2d(a, O.SOBEL_V, dx), g['filt/grad_v']) < TOL
    assert rel_err(O.filt2d(a, O.LAPLACE_2D, dx ** 2), g['filt/lap']) < TOL


def test_filter_convergence_order():
    """analytic check: k3 gradient is 2nd order, k5 4th order on sin(2 pi x)."""
    errs = {}
    for k in (3, 5):
        e = []
        for n in (32, 64):
            x = torch.linspace(0, 1, n + 1, dtype=torch.float64)[:-1]
            u = torch.sin(2 * np.pi * x)[None, None]
            du = O.filt1d(u, O.GRAD_TAPS[k], 1.0 / n)
            e.append(float((du - 2 * np.pi * torch.cos(2 * np.pi * x)).abs().max()))
        errs[k] = np.log2(e[0] / e[1])
    assert 1.9 < errs[3] < 2.1 and 3.8 < errs[5] < 4.2


@pytest.mark.parametrize('system', SYSTEMS)
def test_train_step_loss_and_grads(system):
    """model(train) -> crankNicolson -> calc_neg_log_joint -> backward, as main.py train()."""
    g = golden(system); cfg = cfg_of(g)
    sd = _model_sd(g)
    for k, v in sub(g, 'sd_after1/', strip='model.').items():
        sd[k] = v.clone()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
    sdp = dict(sd); sdp.update(params)
    x = _x(g, cfg).clone().requires_grad_(True)
    uPred = O.densed_forward(system, sdp, x, cfg['blocks'], training=True)
    assert rel_err(uPred, g['train/uPred']) < TOL
    cn, _ = _integ(system, g)
    nout = cfg['out_channels']
    ustar = cn(uPred, x[:, -nout:])
    assert rel_err(ustar, g['train/ustar']) < TOL
    feats = [params[k] for k in params if k.startswith('features.')]
    loss = O.neg_log_joint(uPred, ustar, int(g['train/ntrain']), params['log_beta'], feats, float(g['phys/dt_args']))
    assert loss.shape == (1,)
    assert rel_err(loss, g['train/loss']) < 1e-5
    loss.backward()
    assert rel_err(x.grad, g['train/g_x']) < 2e-4
    for k, p in params.items():
        assert rel_err(p.grad, g['train/g/model.' + k]) < 2e-4, k
    for k, v in sub(g, 'sd_after2/', strip='model.').items():
        assert rel_err(sd[k], v) < TOL, k


@pytest.mark.parametrize('system', SYSTEMS)
@pytest.mark.parametrize('tag', ['diag', 'full'])
def test_swag_sample(system, tag):
    g = golden(system)
    K = int(g['swag/max_models'])
    sw = sub(g, 'swag/sd/')
    names = [k[len('swag/%s/w/' % tag):] for k in g if k.startswith('swag/%s/w/' % tag)]
    noise = [torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith('swag/%s/noise' % tag))]
    means, sqs, covs, z1s, z2s = [], [], [], [], []
    it = iter(noise)
    for n in names:
        u = n.replace('.', '_')
        means.append(sw[u + '_mean']); sqs.append(sw[u + '_sq_mean']); covs.append(sw[u + '_cov_mat_sqrt'])
        z1s.append(next(it))
        z2s.append(next(it) if tag == 'full' else None)
    ws = O.swag_sample(means, sqs, covs, z1s, z2s, K, diag=(tag == 'diag'))
    for n, w in zip(names, ws):
        assert rel_err(w, g['swag/%s/w/%s' % (tag, n)]) < TOL, n


@pytest.mark.parametrize('system', ['burgers1d', 'ks1d'])
def test_shipped_checkpoint_rollout(system):
    """Shipped epoch-200 weights (reference post/networks/) -> deterministic AR rollout."""
    g = golden(system + '_ckpt200'); cfg = cfg_of(golden(system))
    sd = sub(g, 'sd/', strip='base.model.')
    ic = torch.from_numpy(g['ic'])
    step = lambda inp: O.densed_forward(system, sd, inp, cfg['blocks'], training=False)
    ro = O.rollout(system, step, ic, cfg['nic'], g['rollout'].shape[1] - 1)
    errs = [rel_err(ro[:, t], g['rollout'][:, t]) for t in range(ro.shape[1])]
    assert max(errs[:2]) < TOL
    assert max(errs) < (1e-4 if system == 'burgers1d' else 1e-3), errs


def test_shipped_ks_swag_sample():
    g = golden('ks1d_ckpt200'); cfg = cfg_of(golden('ks1d'))
    sd = sub(g, 'sd/')
    names = [k[len('sample/w/'):] for k in g if k.startswith('sample/w/')]
    noise = iter([torch.from_numpy(g[k]) for k in sorted(k for k in g if k.startswith('sample/noise'))])
    means, sqs, covs, z1s, z2s = [], [], [], [], []
    for n in names:
        u = n.replace('.', '_')
        means.append(sd[u + '_mean']); sqs.append(sd[u + '_sq_mean']); covs.append(sd[u + '_cov_mat_sqrt'])
        z1s.append(next(noise)); z2s.append(next(noise))
    ws = O.swag_sample(means, sqs, covs, z1s, z2s, int(g['max_models']))
    msd = {k[len('base.model.'):]: v for k, v in sd.items() if k.startswith('base.model.')}
    for n, w in zip(names, ws):
        assert rel_err(w, g['sample/w/' + n]) < TOL, n
        msd[n[len('model.'):]] = w
    y = O.densed_forward('ks1d', msd, torch.from_numpy(g['ic'])[:, -cfg['nic']:], cfg['blocks'])
    assert rel_err(y, g['sample/y']) < TOL


def test_ic_generator_2d_is_deterministic():
    a = O.ic_burgers2d(16, [0, 1]); b = O.ic_burgers2d(16, [1])
    assert a.shape == (2, 2, 16, 16) and torch.equal(a[1], b[0])
    assert float(a.abs().max()) <= 3.0 + 1e-6


def test_ic_generator_2d_vs_reference_golden():
    """oracle.ic_burgers2d == the reference's InitPeriodicCond2d (tests/golden/ic2d.npz, made by make_golden.py ic2d)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ic2d.npz'))
    for ncells in (16, 64):
        got = O.ic_burgers2d(ncells, [int(i) for i in g['ic/%d/index' % ncells]]).numpy()
        assert np.abs(got - g['ic/%d/u' % ncells]).max() <= 1e-6


# ----------------------------------------------------------------------------- P4 variants (5x5 stencils, upwind, WENO limiter)
def _p4_cases():
    g = golden('p4')
    for key in sorted(set(k.rsplit('/', 1)[0] for k in g if k.endswith('/ustar'))):
        yield key


@pytest.mark.parametrize('key', list(_p4_cases()))
def test_p4_oracle_vs_reference_fixture(key):
    g = golden('p4')
    dim, ks, meth = key.split('/')
    k1, k2 = int(ks[1]), int(ks[2])
    u = torch.from_numpy(g[dim + '/u']).double().requires_grad_(True)
    u0 = torch.from_numpy(g[dim + '/u0']).double().requires_grad_(True)
    dx, nu, dt = float(g[dim + '/dx']), float(g[dim + '/nu']), float(g[dim + '/dt'])
    if dim == '2d':
        us = O.burgers2d_integrate_k(u, u0, dt, dx, nu, k1, k2, cn=(meth == 'crankNicolson'))
    elif meth == 'crankNicolsonLimiter':
        us = O.burgers1d_cn_limiter(u, u0, dt, dx, nu, k2)
    else:
        us = O.burgers1d_upwind(u, u0, dt, dx, nu, k2, cn=(meth == 'crankNicolson'))
    (us * torch.from_numpy(g[dim + '/gout']).double()).sum().backward()
    assert rel_err(us, g[key + '/ustar']) < 1e-5
    assert rel_err(u.grad, g[key + '/g_u']) < 1e-5
    assert rel_err(u0.grad, g[key + '/g_u0']) < 1e-5


def test_p4_filters_oracle_vs_reference_fixture():
    g = golden('p4')
    u = torch.from_numpy(g['1d/u']).double()
    assert rel_err(O.weno_flux_grad(u, float(g['1d/dx'])), g['1d/weno']) < 1e-5
    assert rel_err(O.upwind_grad(u, float(g['1d/dx'])), g['1d/upwind']) < 1e-5
    a = torch.from_numpy(g['2d/u']).double()[:, :1]
    dx = float(g['2d/dx'])
    assert rel_err(O.filt2d_k(a, O.SOBEL_H5, dx), g['2d/filt5/grad_h']) < 1e-5
    assert rel_err(O.filt2d_k(a, O.SOBEL_V5, dx), g['2d/filt5/grad_v']) < 1e-5
    assert rel_err(O.filt2d_k(a, O.LAPLACE_5, dx ** 2), g['2d/filt5/lap']) < 1e-5


# ----------------------------------------------------------------------------- ground-truth FD solver
@pytest.mark.parametrize('go', [(1, 2), (2, 4), (1, 4)])
def test_fd_solver_oracle_vs_reference_fixture(go):
    g = golden('fd_burgers1d')
    rows = g['rows']
    u = O.fd_burgers1d(g['u0'], float(g['dt']), int(g['nsteps']), 1.0 / int(g['ncell']), float(g['nu']), grad_order=go, rows=rows)
    ref = g['u/g%dd%d' % go]
    assert u.shape == ref.shape and u.dtype == np.float32
    # fp64 integration on both sides, results stored in fp32: agreement to fp32 rounding
    assert rel_err(u, ref) < 2e-7
