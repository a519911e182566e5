"""Time the fp32 FFMA conv kernel on the narrow layers (row-tiled vs generic: ARPDE_CONV_RT_OFF=1)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
from oracle import arpde_oracle as O
DEV = 'cuda:0'
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
for name, cin, cout, k, H in [('last 16->2 5x5', 16, 2, 5, 64), ('dense 52->4 3x3', 52, 4, 3, 32)]:
    torch.manual_seed(0)
    x = torch.randn(N, cin, H, H, device=DEV); w = torch.randn(cout, cin, k, k, device=DEV) / (cin * k * k) ** 0.5
    sc = torch.rand(cin, device=DEV) + 0.5; sh = torch.randn(cin, device=DEV) * 0.3
    y = torch.empty(N, cout, H, H, device=DEV)
    def run():
        L.call('arpde_conv_fwd', L.ptr(x), cin * H * H, L.ptr(w), 0, L.ptr(sc), L.ptr(sh), 0, L.ptr(y), 0, N, cin, H, H, cout, k, k, 1, 0, 1, 0, cout, 0, N, L.stream())
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    a = torch.relu(x[:2].double().cpu() * sc.double().cpu().view(1, -1, 1, 1) + sh.double().cpu().view(1, -1, 1, 1))
    ref = O.conv_circ(a, w.double().cpu(), stride=1)
    err = float((y[:2].cpu().double() - ref).abs().max() / ref.abs().max())
    print('%-18s N=%d: %.3f ms  %.1f TFLOP/s  rel err %.2e' % (name, N, ms, 2.0 * N * H * H * cout * cin * k * k / ms / 1e9, err))
