import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
DEV = 'cuda:0'
NB, cin, cout, k, H = 1920, 16, 2, 5, 64
w = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).to(DEV)
sc = (torch.rand(cin) + 0.5).to(DEV); sh = (torch.randn(cin) * 0.3).to(DEV)
x = torch.randn(NB, cin, H, H, device=DEV); y = torch.empty(NB, cout, H, H, device=DEV)
for _ in range(3):
    L.call('arpde_conv_fwd', L.ptr(x), cin * H * H, L.ptr(w), 0, L.ptr(sc), L.ptr(sh), 0, L.ptr(y), 0, NB, cin, H, H, cout, k, k, 1, 0, 1, 0, cout, 0, NB, L.stream())
torch.cuda.synchronize(); print('ok')
