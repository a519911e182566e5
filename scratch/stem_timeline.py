"""GPU (development build): clock64 timeline of CTA 0 of conv_stem_kernel (epilogue warp 0 / MMA warp) at the bench batch."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arpde_b200 import _lib as L
dev = torch.device('cuda:0')
N, G = 1920, 30
x = torch.randn(N, 6, 64, 64, device=dev)
w1 = torch.randn(G, 96, 6, 5, 5, device=dev) / 12
sc = torch.rand(G, 96, device=dev) + 0.5; sh = torch.randn(G, 96, device=dev) * 0.2
zf = int(L.lib.arpde_conv_z_floats(96, 64, 64)); z = torch.empty(N * zf, device=dev)
ns = int(L.lib.arpde_conv_fwd_tc_scratch_floats(6, 96, 5, 5, G)); scr = torch.empty(ns, device=dev)
tl = torch.zeros(3 * 512, dtype=torch.int64, device=dev)
def run():
    L.call('arpde_stem_pair_producer', L.ptr(x), 6 * 4096, L.ptr(w1), w1.stride(0), L.ptr(sc), L.ptr(sh), 96, N, 6, 64, 64, 96, 5, N // G, L.ptr(z), z.numel(),
           L.ptr(scr), ns, L.stream())
for _ in range(3): run()
torch.cuda.synchronize()
L.lib.arpde_conv_tc_set_debug.argtypes = [C.c_void_p]
L.lib.arpde_conv_tc_set_debug(tl.data_ptr())
run(); torch.cuda.synchronize()
L.lib.arpde_conv_tc_set_debug(None)
t = tl.cpu().numpy().reshape(3, 512)
t0 = min(int(t[r][0]) for r in (1, 2) if t[r][0] > 0)
m = t[2]; m = m[m > 0] - t0
print('MMA warp, item 0..: [wait patch, patch in place] then per tile [slot free, issued]')
per_item = 2 + 2 * 17
for it in range(min(3, len(m) // per_item)):
    seg = m[it * per_item:(it + 1) * per_item]
    print(' item %d: patch wait %d -> %d' % (it, seg[0], seg[1]))
    print('   slot-free :', [int(v) for v in seg[2::2]])
    print('   issued    :', [int(v) for v in seg[3::2]])
    print('   issue time per tile:', [int(b - a) for a, b in zip(seg[2::2], seg[3::2])], ' wait for slot:', [int(a - b) for a, b in zip(seg[4::2], seg[3::2])])
e = t[1]; e = e[e > 0] - t0
print('epilogue warp 0: per tile [waiting, complete, drained]')
rows = [(int(e[i]), int(e[i + 1]), int(e[i + 2])) for i in range(0, min(len(e) - 2, 3 * 40), 3)]
print('   complete-to-complete:', [rows[i + 1][1] - rows[i][1] for i in range(len(rows) - 1)])
print('   drain time          :', [r[2] - r[1] for r in rows])
print('   idle before tile    :', [r[1] - r[0] for r in rows])
print('per item: [patch ready - prev item end], [item duration], tiles mean issue, mean slot wait')
nit = len(m) // per_item
prev_end = None
for it in range(nit):
    seg = m[it * per_item:(it + 1) * per_item]
    sf, iss = seg[2::2], seg[3::2]
    gap = int(seg[1] - prev_end) if prev_end is not None else int(seg[1])
    print('  item %2d: gap %6d  dur %6d  issue %5.0f  slotwait %5.0f' % (it, gap, int(iss[-1] - seg[1]), float((iss - sf).mean()), float((sf[1:] - iss[:-1]).mean())))
    prev_end = int(iss[-1])
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): run()
torch.cuda.synchronize(); e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print('producer call: %.3f ms' % (e0.elapsed_time(e1) / 10))
