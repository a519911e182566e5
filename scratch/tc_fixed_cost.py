"""GPU: conv_tc launch time vs batch size for In_conv1 / In_conv3 (fixed cost per launch = intercept)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import arpde_b200._lib as L
dev = torch.device('cuda:0')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, (cin, cout, k, s) in (('In_conv1', (6, 96, 5, 1)), ('In_conv3', (96, 48, 3, 2)), ('conv1x1', (64, 32, 1, 1))):
    H = 64 if name != 'conv1x1' else 32
    for NT, npg in ((30, 1), (60, 2), (120, 4), (240, 8), (480, 16), (960, 32), (1920, 64)):
        G = NT // npg
        x = torch.randn((NT, cin, H, H), device=dev)
        w = torch.randn((G, cout, cin, k, k), device=dev) / (cin * k * k) ** 0.5
        y = torch.empty((NT, cout, H // s, H // s), device=dev)
        n_scr = int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, k, k, G))
        scr = torch.empty((n_scr,), device=dev)
        def f():
            L.call('arpde_conv_fwd_tc', L.ptr(x), cin * H * H, L.ptr(w), cout * cin * k * k, 0, 0, 0, L.ptr(y), 0, NT, cin, H, H, cout, k, k, s, 0, 0, 0, cout, 0, npg, L.ptr(scr), n_scr, L.stream())
        us = t(f)
        print('%-9s NT=%4d groups=%2d  %.1f us  (%.3f us/sample)' % (name, NT, G, us, us / NT))
