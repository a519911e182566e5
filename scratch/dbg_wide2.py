import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import _models as M
DEV = 'cuda:0'
H = 128
tag = sys.argv[1]
torch.manual_seed(5)
m = M.CLS['burgers2d'](6, 2, [4], growth_rate=32, init_features=96).to(DEV)
with torch.no_grad():
    for mod in m.modules():
        if hasattr(mod, 'running_var'):
            mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
x = torch.randn(2, 6, H, H, device=DEV)
m.eval()
xg = x.clone().requires_grad_(True)
y = m(xg)
st = y.grad_fn.saved_tensors
work = st[1].clone(); ss = st[2].clone()
if len(sys.argv) > 2:      # perturb the allocator state before backward
    junk = [torch.empty(1 << 20, device=DEV) for _ in range(7)]
(y ** 2).sum().backward()
torch.cuda.synchronize()
torch.save({'y': y.detach().cpu(), 'work': work.cpu(), 'ss': ss.cpu(), 'gx': xg.grad.cpu(), 'gw1': m.features.In_conv1.weight.grad.cpu(),
            'work_after': st[1].cpu()}, '/tmp/dbg_%s.pt' % tag)
print('saved', tag)
