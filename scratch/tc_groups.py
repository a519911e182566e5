"""Stress: tensor-core conv with many weight groups and many items per persistent CTA vs the FFMA kernel."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from _util import rel_err
from arpde_b200 import _lib as L
DEV = 'cuda:0'
G, NPG = 30, 64
N = G * NPG
LAYERS = [('In_conv1', 6, 96, 5, 1, 0, 64, 0), ('In_conv3', 96, 48, 3, 2, 0, 64, 0), ('dense4', 60, 4, 3, 1, 0, 32, 1),
          ('conv1x1', 64, 32, 1, 1, 0, 32, 1), ('conv2up', 32, 16, 3, 1, 1, 32, 1)]
for name, cin, cout, k, s, up, H, relu in LAYERS:
    torch.manual_seed(0)
    Ho = (H << up) // s
    w = (torch.randn(G, cout, cin, k, k) / (cin * k * k) ** 0.5).to(DEV)
    sc = (torch.rand(G, cin) + 0.5).to(DEV); sh = (torch.randn(G, cin) * 0.3).to(DEV)
    x = torch.randn(N, cin, H, H, device=DEV)
    ys = []
    for fn in ('arpde_conv_fwd', 'arpde_conv_fwd_tc'):
        y = torch.zeros(N, cout, Ho, Ho, device=DEV)
        a = [L.ptr(x), cin * H * H, L.ptr(w), cout * cin * k * k, L.ptr(sc) if relu else 0, L.ptr(sh) if relu else 0, cin, L.ptr(y), 0, N, cin, H, H, cout, k, k, s, up, relu, 0, cout, 0, NPG]
        if fn.endswith('tc'):
            scratch = torch.empty(int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, k, k, G)), device=DEV)
            a += [L.ptr(scratch), scratch.numel()]
        for rep in range(3):
            L.call(fn, *a, L.stream())
        torch.cuda.synchronize()
        ys.append(y)
    print('%-9s groups=%d N=%d  tc vs ffma rel err %.2e' % (name, G, N, rel_err(ys[1], ys[0])), flush=True)
