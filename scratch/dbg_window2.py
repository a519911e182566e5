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
dt = float(g['phys/dt']); dx, nu = float(g['phys/dx']), float(g['phys/nu'])
opt = torch.optim.Adam([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
sd = {k: (v.double() if v.dtype.is_floating_point else v.clone()) for k, v in sub(g, 'sd/', strip='model.').items()}
for k, v in sub(g, 'sd_after2/', strip='model.').items():
    sd[k] = v.double() if v.dtype.is_floating_point else v.clone()
params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and 'running' not in k}
sdp = dict(sd); sdp.update(params)
feats = [params[k] for k in params if k.startswith('features.')]
opt64 = torch.optim.Adam([{'params': [params['log_beta']], 'lr': 1e-3}, {'params': feats}], lr=1e-3)
x = torch.from_numpy(g['trainwin/ics']).to(DEV).repeat(1,3,1,1)
xd = torch.from_numpy(g['trainwin/ics']).double().repeat(1,3,1,1)
loss = 0; l64 = 0
def cmp(tag, grads=False):
    w = 0; wk=''
    for k, p in bayes.named_parameters():
        a = p.grad if grads else p
        b = params[k[len('model.'):]].grad if grads else params[k[len('model.'):]]
        e = rel_err(a, b)
        if e > w: w, wk = e, k
    print(tag, 'worst %.2e at %s' % (w, wk))
for i in range(4):
    uPred = swag(x[:, -6:]); ustar = integ.crankNicolson(uPred, x[:, -2:], dt)
    loss = loss + swag.calc_neg_log_joint(uPred, ustar, 1)
    up = O.densed_forward(system, sdp, xd[:, -6:], cfg['blocks'], training=True)
    l64 = l64 + O.neg_log_joint(up, O.burgers2d_cn(up, xd[:, -2:], dt, dx, nu), 1, params['log_beta'], feats, float(g['phys/dt_args']))
    print('step', i, 'uPred err %.2e' % rel_err(uPred, up), 'loss', float(loss.detach()), float(l64.detach()))
    if (i+1) % 2 == 0:
        loss.backward(); l64.backward(); loss = 0; l64 = 0
        cmp('grads after window %d' % (i//2), True)
        opt.step(); opt64.step()
        cmp('params after step %d' % (i//2))
        opt.zero_grad(); opt64.zero_grad()
        x = torch.cat([x[:, -4:].detach(), uPred.detach()], 1); xd = torch.cat([xd[:, -4:].detach(), up.detach()], 1)
    else:
        x = torch.cat([x, uPred], 1); xd = torch.cat([xd, up], 1)
print('---- detail')
# redo from scratch: one window, inspect elementwise
g, cfg, args, model, bayes, swag = M.build(system, DEV)
M.load_golden_weights(bayes, g, extra=('sd_after2/',))
bayes.train()
x = torch.from_numpy(g['trainwin/ics']).to(DEV).repeat(1,3,1,1)
loss = 0
for i in range(2):
    uPred = swag(x[:, -6:]); ustar = integ.crankNicolson(uPred, x[:, -2:], dt)
    loss = loss + swag.calc_neg_log_joint(uPred, ustar, 1)
    x = torch.cat([x, uPred], 1)
loss.backward()
p = bayes.model.features.In_conv3.weight
gr = p.grad.detach().cpu().double().flatten()
print('grad absmax %.3e, min abs %.3e, #|g|<1e-5*max: %d, #|g|<1e-8: %d of %d' % (gr.abs().max(), gr.abs().min(), int((gr.abs() < 1e-5*gr.abs().max()).sum()), int((gr.abs()<1e-8).sum()), gr.numel()))
q = torch.quantile(gr.abs(), torch.tensor([0.01,0.1,0.5,0.9,0.99], dtype=torch.float64))
print('quantiles', q.tolist())
for k, pp in bayes.named_parameters():
    gg = pp.grad.detach().abs()
    print('%-50s max %.3e  median %.3e' % (k, float(gg.max()), float(gg.median())))
