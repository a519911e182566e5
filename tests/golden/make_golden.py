"""
Generate golden fixtures from the REAL reference (cics-nd/ar-pde-cnn at /root/reference).

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # all three systems, one subprocess each

Shims applied (none edits a reference file):
  1. torch<=1.2 circular padding: every Conv{1,2}d built with padding_mode='circular' gets
     padding //= 2 (reference README.md:40 prescribes exactly this for newer torch).
  2. matplotlib is not installed: MagicMock modules are placed in sys.modules.
  3. noise draws of SwagNN.sample are recorded by wrapping torch.randn_like / Tensor.normal_.

Outputs: tests/golden/{burgers2d,burgers1d,ks1d}.npz (+ *_ckpt200.npz for the shipped 1D
checkpoints) and ic2d.npz (`make_golden.py ic2d`: samples of the on-the-fly initial-condition data set).  The npz files are committed; this script documents how they were made.
"""
import os
import subprocess
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
PROJ = {'burgers2d': '2D-Burgers-SWAG', 'burgers1d': '1D-Burger-SWAG', 'ks1d': '1D-KS-SWAG'}


sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shims import install_shims          # noqa: E402  (padding + matplotlib shims, shared with bench.py / T13)


def gen(system):
    import numpy as np
    import torch
    import torch.nn.functional as F
    install_shims()
    proj = os.path.join(REF, PROJ[system])
    sys.path.insert(0, proj)
    torch.set_num_threads(1)
    torch.manual_seed(12345)
    np.random.seed(12345)

    from nn.bayesNN import BayesNN
    from nn.swag import SwagNN
    if system == 'burgers2d':
        from nn.denseEDcirc2d import DenseED
        from nn.burger2DFiniteDifference import Burger2DIntegrate as Integ
        cfg = dict(in_channels=6, out_channels=2, blocks=[4], growth_rate=4, init_features=96)
        B, shape, nic = 2, (32, 32), 3
        dx, nu, dt_args, dt = 1.0 / 32, 0.005, 0.005, 0.003
        gk = [3, 3]
    elif system == 'burgers1d':
        from nn.denseEDcirc import DenseED
        from nn.burgerFiniteDifference import BurgerIntegrate as Integ
        cfg = dict(in_channels=5, out_channels=1, blocks=[1], growth_rate=4, init_features=64)
        B, shape, nic = 3, (64,), 5
        dx, nu, dt_args, dt = 1.0 / 65, 0.0025, 0.005, 0.002
        gk = [3, 3]
    else:
        from nn.denseEDcirc import DenseED
        from nn.ksFiniteDifference import KSIntegrate as Integ
        cfg = dict(in_channels=2, out_channels=1, blocks=[4], growth_rate=4, init_features=32)
        B, shape, nic = 3, (96,), 2
        dx, nu, dt_args, dt = 22 * np.pi / 96, 1.0, 0.1, 0.05
        gk = [5, 5, 7]

    args = types.SimpleNamespace(dt=dt_args, device=torch.device('cpu'), nic=nic)
    out = {}

    def put_sd(prefix, sd):
        for k, v in sd.items():
            out[prefix + k] = v.detach().cpu().numpy().copy()

    model = DenseED(bn_size=8, drop_rate=0, bottleneck=False, out_activation=None, **cfg)
    bayes = BayesNN(args, model)
    # make BN affine params and running stats non-trivial
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
                m.running_mean.uniform_(-0.2, 0.2)
                m.running_var.uniform_(0.5, 1.5)
    put_sd('sd/', bayes.state_dict())                 # keys 'model.features...', 'model.log_beta'
    out['cfg/blocks'] = np.array(cfg['blocks'])
    out['cfg/growth_rate'] = np.array(cfg['growth_rate'])
    out['cfg/init_features'] = np.array(cfg['init_features'])
    out['cfg/in_channels'] = np.array(cfg['in_channels'])
    out['cfg/out_channels'] = np.array(cfg['out_channels'])
    out['cfg/nic'] = np.array(nic)

    nch_hist = cfg['in_channels'] + (2 if system == 'burgers2d' else 3)
    hist = torch.randn(B, nch_hist, *shape)           # a longer history; model sees a channel slice
    x = hist[:, -cfg['in_channels']:]
    out['hist'] = hist.numpy().copy()

    # ---- eval forward
    bayes.eval()
    with torch.no_grad():
        out['y_eval'] = bayes(x).numpy().copy()

    # ---- train forward x2 (running stats move), then loss + grads like main.py train()
    bayes.train()
    with torch.no_grad():
        out['y_train0'] = bayes(x).numpy().copy()
    put_sd('sd_after1/', {k: v for k, v in bayes.state_dict().items() if 'running' in k or 'num_batches' in k})

    integ = Integ(dx, nu=nu, grad_kernels=gk, device='cpu')
    xin = x.clone().requires_grad_(True)
    uPred = bayes(xin)
    nout = 2 if system == 'burgers2d' else 1
    uPred0 = xin[:, -nout:]
    ustar = integ.crankNicolson(uPred, uPred0, dt)
    ntrain = 7
    loss = bayes.calc_neg_log_joint(uPred, ustar, ntrain)
    loss.backward()
    out['train/uPred'] = uPred.detach().numpy().copy()
    out['train/ustar'] = ustar.detach().numpy().copy()
    out['train/loss'] = loss.detach().numpy().copy()
    out['train/ntrain'] = np.array(ntrain)
    out['train/g_x'] = xin.grad.numpy().copy()
    for k, p in bayes.named_parameters():
        out['train/g/' + k] = p.grad.numpy().copy()
    put_sd('sd_after2/', {k: v for k, v in bayes.state_dict().items() if 'running' in k or 'num_batches' in k})
    out['phys/dx'], out['phys/nu'], out['phys/dt'], out['phys/dt_args'] = map(np.array, (dx, nu, dt, dt_args))
    out['phys/grad_kernels'] = np.array(gk)

    # ---- integrators alone (forward + grads wrt both inputs), smooth fields
    torch.manual_seed(7)
    grid = torch.linspace(0, 1, shape[-1] + 1)[:-1]
    if system == 'burgers2d':
        yy, xx = torch.meshgrid(grid, grid, indexing='ij')
        base = torch.stack([torch.sin(2 * np.pi * (xx + 2 * yy)), torch.cos(2 * np.pi * (2 * xx - yy))], 0)
        a = (base[None] * torch.randn(B, 2, 1, 1) + 0.3 * torch.randn(B, 2, *shape)).requires_grad_(True)
        b = (base[None].flip(1) * torch.randn(B, 2, 1, 1) + 0.3 * torch.randn(B, 2, *shape)).requires_grad_(True)
    else:
        base = torch.sin(2 * np.pi * grid)[None, None]
        a = (base * torch.randn(B, 1, 1) * 2 + 0.3 * torch.randn(B, 1, *shape)).requires_grad_(True)
        b = (base.roll(5, -1) * torch.randn(B, 1, 1) * 2 + 0.3 * torch.randn(B, 1, *shape)).requires_grad_(True)
    gout = torch.randn_like(a)
    for meth in ('crankNicolson', 'backwardEuler'):
        a.grad = b.grad = None
        us = getattr(integ, meth)(a, b, dt)
        (us * gout).sum().backward()
        out['integ/%s/ustar' % meth] = us.detach().numpy().copy()
        out['integ/%s/g_uPred' % meth] = a.grad.numpy().copy()
        out['integ/%s/g_uPred0' % meth] = (b.grad if b.grad is not None else torch.zeros_like(b)).numpy().copy()
    out['integ/uPred'], out['integ/uPred0'], out['integ/gout'] = a.detach().numpy().copy(), b.detach().numpy().copy(), gout.numpy().copy()
    if system != 'burgers2d':
        # other stencil orders through the filter objects directly
        if system == 'ks1d':
            from nn.ksFiniteDifference import GradFilter1D, LaplaceFilter1D, FourthOrderFilter1D
            combos = [('grad', GradFilter1D, [3, 5]), ('lap', LaplaceFilter1D, [3, 5]), ('d4', FourthOrderFilter1D, [5, 7])]
        else:
            from nn.burgerFiniteDifference import GradFilter1D, LaplaceFilter1D, FourthOrderFilter1D
            combos = [('grad', GradFilter1D, [3, 5, 7]), ('lap', LaplaceFilter1D, [3, 5]), ('d4', FourthOrderFilter1D, [5, 7])]
        for nm, cls, ks in combos:
            for k in ks:
                out['filt/%s%d' % (nm, k)] = cls(dx, kernel_size=k, device='cpu')(a.detach()).numpy().copy()
    else:
        out['filt/grad_h'] = integ.grad1.grad_h(a.detach()[:, :1]).numpy().copy()
        out['filt/grad_v'] = integ.grad1.grad_v(a.detach()[:, :1]).numpy().copy()
        out['filt/lap'] = integ.grad2(a.detach()[:, :1]).numpy().copy()

    # ---- one reference train() window: tsteps=4, tback=2, Adam -- final parameters
    bayes.zero_grad()
    import importlib
    main = importlib.import_module('main')
    swag = SwagNN(args, bayes, full_cov=True, max_models=3)
    params = [{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}]
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=0.0)
    if system == 'burgers2d':
        ics = torch.randn(B, 2, *shape) * 0.5
        loader = [ics]
    elif system == 'burgers1d':
        ics = (torch.randn(B, 1, *shape) * 0.5).repeat(1, 20, 1)
        loader = [ics]
    else:
        ics = (torch.randn(B, 1, *shape) * 0.5).repeat(1, 5, 1)
        loader = [ics]
    out['trainwin/ics'] = ics.numpy().copy()
    # state before the window == 'sd/' params + 'sd_after2/' BN statistics (not stored twice)
    main.epoch = 1
    if system == 'ks1d':    # 1D-KS-SWAG/main.py:26 has a different argument order
        lt, mt = main.train(args, swag, integ, loader, opt, dt=dt, tsteps=np.array([4]), tback=np.array([2]), tstart=np.array([0]))
    else:
        lt, mt = main.train(args, swag, integ, loader, opt, np.array([4]), np.array([2]), np.array([0]), dt)
    out['trainwin/loss_total'], out['trainwin/mse_total'] = np.array(lt), np.array(mt)
    put_sd('trainwin/sd_after/', bayes.state_dict())

    # ---- SWAG: collect a few perturbed models, then sample with recorded noise
    for it in range(5):
        with torch.no_grad():
            for p in bayes.parameters():
                p.add_(0.01 * torch.randn_like(p))
        swag.collect()
    put_sd('swag/sd/', swag.state_dict())
    out['swag/max_models'] = np.array(3)
    for tag, diag in (('diag', True), ('full', False)):
        rec = []
        o_rl, o_n = torch.randn_like, torch.Tensor.normal_

        def rl(t, *a_, **k_):
            z = o_rl(t, *a_, **k_); rec.append(z.clone()); return z

        def nn_(self, *a_, **k_):
            r = o_n(self, *a_, **k_); rec.append(self.clone()); return r
        torch.randn_like, torch.Tensor.normal_ = rl, nn_
        try:
            sm = swag.sample(diagCov=diag)
        finally:
            torch.randn_like, torch.Tensor.normal_ = o_rl, o_n
        for i, z in enumerate(rec):
            out['swag/%s/noise%03d' % (tag, i)] = z.numpy().copy()
        put_sd('swag/%s/w/' % tag, dict(sm.named_parameters()))

    np.savez_compressed(os.path.join(HERE, system + '.npz'), **out)
    print(system, 'fixture keys:', len(out))

    # ---- shipped checkpoints (1D only; 2D blobs are absent, .MISSING_LARGE_BLOBS)
    if system in ('burgers1d', 'ks1d'):
        ck = {}
        K = 30 if system == 'burgers1d' else 10
        sargs = types.SimpleNamespace(dt=dt_args, device=torch.device('cpu'), nic=nic)
        m2 = DenseED(bn_size=8, drop_rate=0, bottleneck=False, out_activation=None, **cfg)
        b2 = BayesNN(sargs, m2)
        s2 = SwagNN(sargs, b2, full_cov=True, max_models=K)
        s2.loadModel(200, file_dir=os.path.join(proj, 'post', 'networks'))
        full = s2.state_dict()
        for k, v in full.items():
            if system == 'burgers1d' and k.endswith('_cov_mat_sqrt'):
                continue        # 1.6 MB of low-rank rows: KS fixture carries the full-cov case
            ck['sd/' + k] = v.numpy().copy()
        ck['max_models'] = np.array(K)
        # deterministic rollout of the base model on a synthetic Fourier IC
        L = 512 if system == 'burgers1d' else 96
        T = 50 if system == 'burgers1d' else 20
        np.random.seed(12345)
        if system == 'burgers1d':
            xg = np.linspace(0, 1, L + 1)
            lam = np.random.randn(2, 7); c = np.random.rand() - 0.5
            kx = np.outer(np.arange(-3, 4), xg) * 2 * np.pi
            f = np.dot(lam[[0]], np.cos(kx)) + np.dot(lam[[1]], np.sin(kx))
            f = 2 * f / np.amax(np.abs(f)) + 2 * c
            ic = torch.Tensor(f[:, :-1]).unsqueeze(0).repeat(1, 20, 1)
        else:
            xg = np.linspace(0, 22 * np.pi, L + 1)
            ic = torch.Tensor(3.0 * np.sin(2 * np.pi * 3 * xg[:-1] / (22 * np.pi)) + np.sin(2 * np.pi * 5 * xg[:-1] / (22 * np.pi))
                              ).reshape(1, 1, -1).repeat(1, 5, 1)
        ck['ic'] = ic.numpy().copy()
        b2.eval()
        inp = ic.clone()
        outs = [inp[:, :1].clone()]
        with torch.no_grad():
            for t in range(T):
                up = b2(inp[:, -nic:])
                outs.append(up[:, :1].clone())
                inp = torch.cat([inp[:, -(nic - 1):], up[:, :1]], 1)
        ck['rollout'] = torch.stack(outs, 1).numpy().copy()
        if system == 'ks1d':
            rec = []
            o_rl, o_n = torch.randn_like, torch.Tensor.normal_

            def rl(t, *a_, **k_):
                z = o_rl(t, *a_, **k_); rec.append(z.clone()); return z

            def nn_(self, *a_, **k_):
                r = o_n(self, *a_, **k_); rec.append(self.clone()); return r
            torch.randn_like, torch.Tensor.normal_ = rl, nn_
            try:
                sm = s2.sample(diagCov=False)
            finally:
                torch.randn_like, torch.Tensor.normal_ = o_rl, o_n
            for i, z in enumerate(rec):
                ck['sample/noise%03d' % i] = z.numpy().copy()
            for k, p in sm.named_parameters():
                ck['sample/w/' + k] = p.detach().numpy().copy()
            sm.eval()
            with torch.no_grad():
                ck['sample/y'] = sm(ic[:, -nic:]).numpy().copy()
        np.savez_compressed(os.path.join(HERE, system + '_ckpt200.npz'), **ck)
        print(system, 'checkpoint fixture keys:', len(ck))


def gen_ic2d():
    """Initial conditions of the reference's on-the-fly data set (2D-Burgers-SWAG/utils/burgerLoader2D.py:80-116)."""
    import numpy as np
    install_shims()
    sys.path.insert(0, os.path.join(REF, PROJ['burgers2d']))
    from utils.burgerLoader2D import InitPeriodicCond2d
    out = {}
    for ncells, idx in [(16, [0, 1, 7]), (64, [3, 12345])]:
        ds = InitPeriodicCond2d(ncells, 100000)
        out['ic/%d/index' % ncells] = np.array(idx)
        out['ic/%d/u' % ncells] = np.stack([ds[i].numpy() for i in idx], 0)
    np.savez_compressed(os.path.join(HERE, 'ic2d.npz'), **out)
    print('wrote ic2d.npz', {k: v.shape for k, v in out.items()})


def gen_p4():
    """integrator / filter variants no reference driver calls -- 5x5 2D stencils, kernel_size=2 "conditional
    upwind" gradient, WENO flux derivative and crankNicolsonLimiter -- forward values and gradients w.r.t. both inputs."""
    import importlib
    import numpy as np
    import torch
    from oracle import ref_shims
    out = {}
    torch.manual_seed(11)
    ref_shims.import_reference('burgers1d')
    R = importlib.import_module('nn.burgerFiniteDifference')
    L_ = 64
    dx, nu, dt = 1.0 / (L_ + 1), 0.0025, 0.002
    grid = torch.linspace(0, 1, L_ + 1)[:-1]
    u = (2 * torch.sin(2 * np.pi * grid)[None, None] * torch.randn(3, 1, 1) + 0.3 * torch.randn(3, 1, L_)).requires_grad_(True)
    u0 = (2 * torch.cos(2 * np.pi * grid)[None, None] * torch.randn(3, 1, 1) + 0.3 * torch.randn(3, 1, L_)).requires_grad_(True)
    g = torch.randn(3, 1, L_)
    out['1d/u'], out['1d/u0'], out['1d/gout'] = u.detach().numpy().copy(), u0.detach().numpy().copy(), g.numpy().copy()
    out['1d/dx'], out['1d/nu'], out['1d/dt'] = np.array(dx), np.array(nu), np.array(dt)
    for gk in ([2, 3], [2, 5], [3, 3]):
        integ = R.BurgerIntegrate(dx, nu=nu, grad_kernels=gk, device='cpu')
        meths = ('crankNicolson', 'backwardEuler', 'crankNicolsonLimiter') if gk[0] == 2 else ('crankNicolsonLimiter',)
        for meth in meths:
            u.grad = u0.grad = None
            us = getattr(integ, meth)(u, u0, dt)
            (us * g).sum().backward()
            key = '1d/k%d%d/%s/' % (gk[0], gk[1], meth)
            out[key + 'ustar'] = us.detach().numpy().copy()
            out[key + 'g_u'] = u.grad.numpy().copy()
            out[key + 'g_u0'] = u0.grad.numpy().copy()
    out['1d/weno'] = R.FluxWENOFilter1D(dx)(u.detach()).numpy().copy()
    out['1d/upwind'] = R.GradFilter1D(dx, kernel_size=2)(u.detach()).numpy().copy()
    for k in [k for k in sys.modules if k == 'nn' or k.startswith('nn.')]:
        del sys.modules[k]
    ref_shims.import_reference('burgers2d')
    R2 = importlib.import_module('nn.burger2DFiniteDifference')
    n = 16
    dx2, nu2, dt2 = 1.0 / n, 0.005, 0.003
    a = torch.randn(2, 2, n, n).requires_grad_(True)
    b = torch.randn(2, 2, n, n).requires_grad_(True)
    g2 = torch.randn(2, 2, n, n)
    out['2d/u'], out['2d/u0'], out['2d/gout'] = a.detach().numpy().copy(), b.detach().numpy().copy(), g2.numpy().copy()
    out['2d/dx'], out['2d/nu'], out['2d/dt'] = np.array(dx2), np.array(nu2), np.array(dt2)
    for gk in ([5, 5], [5, 3], [3, 5]):
        integ = R2.Burger2DIntegrate(dx2, nu=nu2, grad_kernels=gk, device='cpu')
        for meth in ('crankNicolson', 'backwardEuler'):
            a.grad = b.grad = None
            us = getattr(integ, meth)(a, b, dt2)
            (us * g2).sum().backward()
            key = '2d/k%d%d/%s/' % (gk[0], gk[1], meth)
            out[key + 'ustar'] = us.detach().numpy().copy()
            out[key + 'g_u'] = a.grad.numpy().copy()
            out[key + 'g_u0'] = (b.grad if b.grad is not None else torch.zeros_like(b)).numpy().copy()
    s5, l5 = R2.SobelFilter2d(dx2, kernel_size=5), R2.LapLaceFilter2d(dx2, kernel_size=5)
    out['2d/filt5/grad_h'] = s5.grad_h(a.detach()[:, :1]).numpy().copy()
    out['2d/filt5/grad_v'] = s5.grad_v(a.detach()[:, :1]).numpy().copy()
    out['2d/filt5/lap'] = l5(a.detach()[:, :1]).numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'p4.npz'), **out)
    print('wrote p4.npz,', len(out), 'keys')


def gen_fd():
    """burgers1d() of the reference's ground-truth script 1D-Burger-SWAG/solver/fd_burger1D.py:55-98 (RK4 + upwind,
    fp64), called as a function with the module globals its __main__ block would set (args.nu, dx, idx)."""
    import importlib
    import types
    import numpy as np
    import torch
    from oracle import ref_shims
    proj = ref_shims.import_reference('burgers1d')
    sys.path.insert(0, os.path.join(proj, 'solver'))
    for k in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')]:
        del sys.modules[k]
    fd = importlib.import_module('fd_burger1D')
    out = {}
    ncell, nu, dt, nsteps = 128, 0.0025, 5e-4, 600
    fd.args = types.SimpleNamespace(nu=nu)
    fd.dx = 1.0 / ncell
    fd.idx = 0
    rng = np.random.RandomState(3)
    x = np.linspace(0, 1, ncell + 1)[:-1]
    u0 = np.stack([a * np.sin(2 * np.pi * (x + b)) + c * np.cos(4 * np.pi * x) + d for a, b, c, d in rng.randn(3, 4)])
    out['u0'] = u0
    out['ncell'], out['nu'], out['dt'], out['nsteps'] = map(np.array, (ncell, nu, dt, nsteps))
    rows = np.array([0, 1, 9, 99, 299, 599])
    out['rows'] = rows
    for go in ([1, 2], [2, 4], [1, 4]):
        res = []
        for i in range(u0.shape[0]):
            with torch.no_grad():
                u_fd, _ = fd.burgers1d(u0[i:i + 1], dt, nsteps, grad_order=go, device='cpu')
            res.append(u_fd[rows])
        out['u/g%dd%d' % tuple(go)] = np.stack(res)             # float32 [3, len(rows), ncell]
    np.savez_compressed(os.path.join(HERE, 'fd_burgers1d.npz'), **out)
    print('wrote fd_burgers1d.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'fd':
        gen_fd()
    elif len(sys.argv) > 1 and sys.argv[1] == 'p4':
        gen_p4()
    elif len(sys.argv) > 1 and sys.argv[1] == 'ic2d':
        gen_ic2d()
    elif len(sys.argv) > 1:
        gen(sys.argv[1])
    else:
        for s in PROJ:
            cwd = '/tmp/arpde_golden_' + s
            os.makedirs(cwd, exist_ok=True)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), s], cwd=cwd)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), 'ic2d'], cwd='/tmp')
        subprocess.check_call([sys.executable, os.path.abspath(__file__), 'p4'], cwd='/tmp')
        subprocess.check_call([sys.executable, os.path.abspath(__file__), 'fd'], cwd='/tmp')
