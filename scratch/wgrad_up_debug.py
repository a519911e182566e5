"""GPU debug: weight gradient of folded upsampling layers for small / odd channel counts (auto path 0 vs generic path 1 vs fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import arpde_b200
from arpde_b200 import _lib as L
from oracle import arpde_oracle as O
from _util import rel_err
DEV = 'cuda:0'
import ctypes
plan = (ctypes.c_int32 * 400)()
for (cin, cout, H, W, N, up) in [(9, 4, 16, 16, 3, 1), (9, 8, 16, 16, 3, 1), (12, 4, 16, 16, 3, 1), (32, 4, 16, 16, 3, 1), (9, 16, 16, 16, 3, 1), (32, 16, 32, 32, 3, 1), (10, 5, 16, 16, 3, 1),
                               (24, 12, 16, 16, 2, 1), (9, 4, 8, 8, 3, 1), (9, 4, 32, 32, 2, 1), (9, 4, 16, 16, 3, 0), (9, 16, 16, 16, 3, 0)]:
    torch.manual_seed(cin + cout)
    x = torch.randn(N, cin, H, W); Ho, Wo = H << up, W << up
    gy = torch.randn(N, cout, Ho, Wo)
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    a = torch.relu(x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    if up: a = O._up2(a)
    w = torch.zeros((cout, cin, 3, 3), dtype=torch.float64, requires_grad=True)
    O.conv_circ(a, w).backward(gy.double())
    xd, gd, scd, shd = x.to(DEV), gy.to(DEV), sc.to(DEV), sh.to(DEV)
    d64 = torch.empty((cout * cin * 9,), device=DEV, dtype=torch.float64)
    n_scr = int(L.lib.arpde_conv_wgrad_tc_scratch(N, cin, H, W, cout, 3, 3, 1, up))
    scr = torch.empty((max(n_scr, 1),), device=DEV)
    errs = []
    for path in (0, 1):
        dW = torch.full((cout, cin, 3, 3), 5.0, device=DEV)
        L.call('arpde_conv_wgrad', L.ptr(xd), cin * H * W, L.ptr(scd), L.ptr(shd), L.ptr(gd), cout * Ho * Wo, L.ptr(d64), L.ptr(dW), N, cin, H, W, cout, 3, 3, 1, up, 1,
               L.ptr(scr) if n_scr else 0, n_scr, path, L.stream())
        torch.cuda.synchronize()
        errs.append(rel_err(dW, w.grad))
    rc = L.lib.arpde_conv_wgrad_tc_plan(cin, H, W, cout, 3, 3, 1, up, plan)
    print((cin, cout, H, W, N, up), 'tc scratch', n_scr, 'plan rc', rc, 'auto %.1e generic %.1e' % tuple(errs))
