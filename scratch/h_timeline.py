"""GPU (development build): clock64 timeline of CTA 0 of conv_h (converter / epilogue / MMA roles), first items."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arpde_b200 import _lib as L
dev = torch.device('cuda:0')
N, G = 1920, 30
which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cin, cout, k, up, H = [(64, 32, 1, 0, 32), (32, 16, 3, 1, 32)][which]
xs = torch.randn(N, cin, H, H, device=dev)
w = torch.randn(G, cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
sc = torch.rand(G, cin, device=dev) + 0.5; sh = torch.randn(G, cin, device=dev) * 0.2
f = 2 if up else 1
y = torch.empty(N, cout, H * f, H * f, device=dev)
ns = int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, G))
scr = torch.empty(ns, device=dev)
tl = torch.zeros(3 * 512, dtype=torch.int64, device=dev)
def h():
    L.call('arpde_conv_fwd_h', L.ptr(xs), cin * H * H, L.ptr(w), w.stride(0), L.ptr(sc), L.ptr(sh), cin, L.ptr(y), N, cin, H, H, cout, k, up, 1,
           cout, 0, N // G, L.ptr(scr), ns, L.stream())
for _ in range(3): h()
torch.cuda.synchronize()
L.lib.arpde_conv_tc_set_debug.argtypes = [C.c_void_p]
L.lib.arpde_conv_tc_set_debug(tl.data_ptr())
h(); torch.cuda.synchronize()
L.lib.arpde_conv_tc_set_debug(None)
t = tl.cpu().numpy().reshape(3, 512)
t0 = min(int(t[r][0]) for r in range(3) if t[r][0] > 0)
ntile = 2 if which == 0 else 3
print('converter (per batch: load begins, store begins, buffer free, stored, arrived), cycles from start')
c = t[0]; c = c[c > 0] - t0
for i in range(0, min(len(c), 60), 5): print('  batch %2d (item %d):' % (i // 5, i // 10), list(c[i:i + 5]))
print('MMA warp (per item: wait, patch in place, tiles issued)')
m = t[2]; m = m[m > 0] - t0
for i in range(0, min(len(m), 8 * (2 + ntile)), 2 + ntile): print('  item %2d:' % (i // (2 + ntile)), list(m[i:i + 2 + ntile]))
print('epilogue warp 0 (per tile: waiting, complete, drained)')
e = t[1]; e = e[e > 0] - t0
for i in range(0, min(len(e), 3 * 8 * ntile), 3): print('  tile %2d:' % (i // 3), list(e[i:i + 3]))
