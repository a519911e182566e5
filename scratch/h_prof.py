"""GPU: ncu target -- conv_h on the two decoder layers at the bench batch (2 launches each)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arpde_b200 import _lib as L
dev = torch.device('cuda:0')
N, G = 1920, 30
for (cin, cout, k, up, H) in [(64, 32, 1, 0, 32), (32, 16, 3, 1, 32)]:
    xs = [torch.randn(N, cin, H, H, device=dev) for _ in range(2)]
    w = torch.randn(G, cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    sc = torch.rand(G, cin, device=dev) + 0.5; sh = torch.randn(G, cin, device=dev) * 0.2
    f = 2 if up else 1
    y = torch.empty(N, cout, H * f, H * f, device=dev)
    ns = int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, G))
    scr = torch.empty(ns, device=dev)
    for i in range(2):
        L.call('arpde_conv_fwd_h', L.ptr(xs[i]), cin * H * H, L.ptr(w), w.stride(0), L.ptr(sc), L.ptr(sh), cin, L.ptr(y), N, cin, H, H, cout, k, up, 1,
               cout, 0, N // G, L.ptr(scr), ns, L.stream())
    torch.cuda.synchronize()
print('ok')
