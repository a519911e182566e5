"""Build arpde_b200 models from golden fixtures (shared by CPU and GPU tests)."""
import types

import numpy as np
import torch

from _util import cfg_of, golden, sub

from arpde_b200.nn import densed as D
from arpde_b200.nn.bayesnn import BayesNN
from arpde_b200.nn.swag import SwagNN
from arpde_b200.nn import integrators as I

CLS = {'burgers2d': D.DenseED2d, 'burgers1d': D.DenseED1dBurgers, 'ks1d': D.DenseED1dKS}


def build(system, device='cpu', max_models=3, dt_args=None):
    g = golden(system)
    cfg = cfg_of(g)
    dt_args = float(g['phys/dt_args']) if dt_args is None else dt_args
    args = types.SimpleNamespace(dt=dt_args, device=torch.device(device), nic=cfg['nic'])
    np.random.seed(0)
    model = CLS[system](cfg['in_channels'], cfg['out_channels'], cfg['blocks'], growth_rate=cfg['growth_rate'],
                        init_features=cfg['init_features'], bn_size=8, drop_rate=0, bottleneck=False, out_activation=None)
    model = model.to(device)
    bayes = BayesNN(args, model)
    swag = SwagNN(args, bayes, full_cov=True, max_models=max_models)
    return g, cfg, args, model, bayes, swag


def load_golden_weights(bayes, g, prefix='sd/', extra=()):
    sd = {k: v for k, v in sub(g, prefix).items()}
    for p in extra:
        sd.update(sub(g, p))
    bayes.load_state_dict(sd, strict=True)


def integrator(system, g, device):
    dx, nu = float(g['phys/dx']), float(g['phys/nu'])
    gk = [int(v) for v in g['phys/grad_kernels']]
    if system == 'burgers2d':
        return I.Burger2DIntegrate(dx, nu=nu, grad_kernels=gk, device=device)
    if system == 'burgers1d':
        return I.BurgerIntegrate(dx, nu=nu, grad_kernels=gk, device=device)
    return I.KSIntegrate(dx, nu=nu, grad_kernels=gk, device=device)
