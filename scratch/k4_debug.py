"""GPU debug: gradients of the k4/s2 circular convolution (TransDown) against fp64 autograd."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import arpde_b200
from arpde_b200.nn import densed as D
from oracle import arpde_oracle as O
dev = 'cuda:0'
def rel(a, b): return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max() / b.detach().double().abs().max())
torch.manual_seed(0)
for (N, cin, cout, H, W, k, s) in [(3, 12, 12, 16, 16, 4, 2), (3, 8, 8, 32, 32, 4, 2), (2, 40, 40, 16, 16, 4, 2), (3, 12, 12, 16, 16, 3, 2), (3, 12, 12, 1, 32, 4, 2)]:
    one_d = H == 1
    conv = (D.CircConv1d if one_d else D.CircConv2d)(cin, cout, k, s).to(dev)
    x = torch.randn(N, cin, W, device=dev) if one_d else torch.randn(N, cin, H, W, device=dev)
    xg = x.clone().requires_grad_(True)
    y = conv(xg)
    gy = torch.randn_like(y)
    y.backward(gy)
    x64 = x.double().cpu().requires_grad_(True); w64 = conv.weight.detach().double().cpu().requires_grad_(True)
    yr = O.conv_circ(x64, w64, stride=s)
    yr.backward(gy.double().cpu())
    e = (xg.grad.double().cpu() - x64.grad).abs()
    print((N, cin, cout, H, W, k, s), 'fwd %.1e  gx %.1e  gw %.1e' % (rel(y, yr), rel(xg.grad, x64.grad), rel(conv.weight.grad, w64.grad)),
          'worst gx idx', [int(v) for v in torch.nonzero(e == e.max())[0]], 'n>1e-4:', int((e > 1e-4 * float(x64.grad.abs().max())).sum()))
