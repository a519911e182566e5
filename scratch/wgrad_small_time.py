"""Narrow-layer weight gradient: sliding-window kernel (path 0) vs generic FFMA kernel (path 1) vs tensor core, N = 128."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
DEV = 'cuda:0'
N = 128
for name, cin, cout, k, H in [('last 16->2 5x5', 16, 2, 5, 64), ('dense1 48->4 3x3', 48, 4, 3, 32), ('dense4 60->4 3x3', 60, 4, 3, 32)]:
    torch.manual_seed(0)
    x = torch.randn(N, cin, H, H, device=DEV); gy = torch.randn(N, cout, H, H, device=DEV)
    sc = torch.rand(cin, device=DEV) + 0.5; sh = torch.randn(cin, device=DEV) * 0.3
    d64 = torch.empty((cout * cin * k * k,), device=DEV, dtype=torch.float64)
    res = {}
    for tag, path in (('small', 0), ('generic', 1)):
        dW = torch.empty((cout, cin, k, k), device=DEV)
        def run():
            L.call('arpde_conv_wgrad', L.ptr(x), cin * H * H, L.ptr(sc), L.ptr(sh), L.ptr(gy), cout * H * H, L.ptr(d64), L.ptr(dW), N, cin, H, H, cout, k, k, 1, 0, 1, 0, 0, path, L.stream())
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        res[tag] = (e0.elapsed_time(e1) / 10, dW.clone())
    a = torch.relu(x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    p = (k - 1) // 2
    ref = torch.nn.grad.conv2d_weight(torch.nn.functional.pad(a, (p, p, p, p), mode='circular'), (cout, cin, k, k), gy.double())
    err = lambda t: float((t.double() - ref).abs().max() / ref.abs().max())
    print('%-18s small %.3f ms (err %.1e) | generic %.3f ms (err %.1e)' % (name, res['small'][0], err(res['small'][1]), res['generic'][0], err(res['generic'][1])))
