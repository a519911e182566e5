"""GPU debug: per-parameter gradient errors of the (2,3,2) 2D topology vs fp64 oracle autograd."""
import os, sys, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _models as M
from oracle import arpde_oracle as O
from _util import rel_err
dev = 'cuda:0'
for blocks, act, shape, seed in (([2, 3, 2], 'tanh', (3, 6, 32, 32), 3), ([4], None, (3, 6, 32, 32), 3), ([4], None, (3, 6, 32, 32), 4), ([4], None, (3, 6, 32, 32), 5), ([4], None, (3, 6, 32, 32), 6)):
  for training in (True, False):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.CLS['burgers2d'](6, 2, blocks, growth_rate=4, init_features=16 if len(blocks) > 1 else 96, out_activation=act).to(dev)
    with torch.no_grad():
        for mod in m.modules():
            if hasattr(mod, 'running_var'):
                mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
    m.train(training)
    x = torch.randn(*shape, device=dev)
    sd = {'features.' + k: v.detach().cpu().clone() for k, v in m.features.state_dict().items()}
    xg = x.clone().requires_grad_(True)
    y = m(xg); gy = torch.randn_like(y); (y * gy).sum().backward()
    sd64 = {k: (v.double().requires_grad_('running' not in k) if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    x64 = x.cpu().double().requires_grad_(True)
    ref = O.densed_forward('burgers2d', sd64, x64, blocks, training=training, out_activation=act)
    (ref * gy.cpu().double()).sum().backward()
    print('seed', seed, 'blocks', blocks, 'act', act, 'training', training, 'fwd %.1e gx %.1e' % (rel_err(y, ref), rel_err(xg.grad, x64.grad)))
    for name, p_ in m.features.named_parameters():
        e = rel_err(p_.grad, sd64['features.' + name].grad)
        if e > 5e-6: print('   %-45s %.1e' % (name, e))
