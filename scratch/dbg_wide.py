import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from _util import rel_err
import _models as M
from oracle import arpde_oracle as O
DEV = 'cuda:0'
H = int(sys.argv[1]) if len(sys.argv) > 1 else 32
growth = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(5)
m = M.CLS['burgers2d'](6, 2, [4], growth_rate=growth, init_features=96).to(DEV)
with torch.no_grad():
    for mod in m.modules():
        if hasattr(mod, 'running_var'):
            mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
x = torch.randn(2, 6, H, H, device=DEV)
sd = {'features.' + k: v.detach().cpu().clone() for k, v in m.features.state_dict().items()}
m.eval()
xg = x.clone().requires_grad_(True)
y = m(xg)
(y ** 2).sum().backward()
sd64 = {k: (v.double().requires_grad_('running' not in k) if v.dtype.is_floating_point else v) for k, v in sd.items()}
x64 = x.cpu().double().requires_grad_(True)
y64 = O.densed_forward('burgers2d', sd64, x64, [4], training=False)
(y64 ** 2).sum().backward()
print('H', H, 'growth', growth, 'fwd', rel_err(y, y64), 'gx', rel_err(xg.grad, x64.grad))
for name, p_ in m.features.named_parameters():
    print('  %-40s %s %.2e' % (name, tuple(p_.shape), rel_err(p_.grad, sd64['features.' + name].grad)))
