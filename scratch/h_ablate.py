"""GPU (development build): conv_h ablations on the two decoder layers at the bench batch (ARPDE_H_DBG bits: 1 no loads, 2 no conversion, 4 no MMAs, 8 no stores)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for dbg in (0, 1, 2, 3, 4, 8, 12, 14, 15):
        env = dict(os.environ, ARPDE_H_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, 'run'], env=env, capture_output=True, text=True).stdout
        print('dbg', dbg, out.strip().replace('\n', ' | '))
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
from arpde_b200 import _lib as L
dev = torch.device('cuda:0')
N, G = 1920, 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (cin, cout, k, up, H) in [(64, 32, 1, 0, 32), (32, 16, 3, 1, 32)]:
    xs = [torch.randn(N, cin, H, H, device=dev) for _ in range(3)]
    w = torch.randn(G, cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    sc = torch.rand(G, cin, device=dev) + 0.5; sh = torch.randn(G, cin, device=dev) * 0.2
    f = 2 if up else 1
    y = torch.empty(N, cout, H * f, H * f, device=dev)
    ns = int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, G))
    scr = torch.empty(ns, device=dev)
    def h(i):
        L.call('arpde_conv_fwd_h', L.ptr(xs[i % 3]), cin * H * H, L.ptr(w), w.stride(0), L.ptr(sc), L.ptr(sh), cin, L.ptr(y), N, cin, H, H, cout, k, up, 1,
               cout, 0, N // G, L.ptr(scr), ns, L.stream())
    for i in range(3): h(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(10): h(i)
    e1.record(); torch.cuda.synchronize()
    print('k%d: %.3f ms' % (k, e0.elapsed_time(e1) / 10))
