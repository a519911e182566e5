import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
from _util import golden, sub, rel_err, cfg_of
import _models as M
from oracle import arpde_oracle as O
DEV='cuda:0'
system='burgers2d'
g, cfg, args, model, bayes, swag = M.build(system, DEV)
M.load_golden_weights(bayes, g, extra=('sd_after2/',))
bayes.train()
integ = M.integrator(system, g, DEV)
inp = torch.from_numpy(g['trainwin/ics']).to(DEV).repeat(1,3,1,1)
dt = float(g['phys/dt'])
# ours: first two steps then backward
loss = 0
x = inp
ups = []
for i in range(2):
    uPred = swag(x[:, -6:]); ups.append(uPred)
    ustar = integ.crankNicolson(uPred, x[:, -2:], dt)
    loss = loss + swag.calc_neg_log_joint(uPred, ustar, 1)
    x = torch.cat([x, uPred], 1)
loss.backward()
# oracle fp64
sd = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in sub(g, 'sd/', strip='model.').items()}
for k, v in sub(g, 'sd_after2/', strip='model.').items():
    sd[k] = v.double() if v.dtype.is_floating_point else v.clone()
params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
sdp = dict(sd); sdp.update(params)
xd = torch.from_numpy(g['trainwin/ics']).double().repeat(1,3,1,1)
lossd = 0
dx, nu = float(g['phys/dx']), float(g['phys/nu'])
for i in range(2):
    up = O.densed_forward(system, sdp, xd[:, -6:], cfg['blocks'], training=True)
    print('step', i, 'fwd rel err', rel_err(ups[i], up))
    us = O.burgers2d_cn(up, xd[:, -2:], dt, dx, nu)
    lossd = lossd + O.neg_log_joint(up, us, 1, params['log_beta'], [params[k] for k in params if k.startswith('features.')], float(g['phys/dt_args']))
    xd = torch.cat([xd, up], 1)
lossd.backward()
print('loss', float(loss), float(lossd))
for k, p in bayes.named_parameters():
    print('%-55s %.2e' % (k, rel_err(p.grad, params[k[len('model.'):]].grad)))
