"""Per-layer timing of the weight-gradient kernels (tensor-core vs FFMA) at the training batch (N=128, C4 layers)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
DEV = 'cuda:0'
dbg = torch.zeros(16, dtype=torch.int64, device=DEV)
L.call('arpde_conv_tc_set_debug', dbg.data_ptr())
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
LAYERS = [('In_conv1', 6, 96, 5, 1, 0, 64, 64), ('In_conv3', 96, 48, 3, 2, 0, 64, 64), ('up3x3', 32, 16, 3, 1, 1, 32, 32),
          ('dense4', 60, 4, 3, 1, 0, 32, 32), ('1x1', 64, 32, 1, 1, 0, 32, 32), ('last', 16, 2, 5, 1, 0, 64, 64)]
for name, cin, cout, k, s, up, H, W in LAYERS:
    x = torch.randn(N, cin, H, W, device=DEV)
    Ho, Wo = (H << up) // s, (W << up) // s
    gy = torch.randn(N, cout, Ho, Wo, device=DEV)
    sc, sh = torch.rand(cin, device=DEV) + 0.5, torch.randn(cin, device=DEV) * 0.3
    d64 = torch.empty((cout * cin * k * k,), device=DEV, dtype=torch.float64)
    dW = torch.empty((cout, cin, k, k), device=DEV)
    n = int(L.lib.arpde_conv_wgrad_tc_scratch(N, cin, H, W, cout, k, k, s, up))
    if n == 0:
        print('%-9s not covered by the tensor-core kernel' % name); continue
    scratch = torch.empty((n,), device=DEV)
    def run():
        L.call('arpde_conv_wgrad_tc', L.ptr(x), cin * H * W, L.ptr(sc), L.ptr(sh), L.ptr(gy), cout * Ho * Wo, L.ptr(d64), L.ptr(dW), N, cin, H, W,
               cout, k, k, s, up, 1, L.ptr(scratch), n, L.stream())
    for _ in range(int(os.environ.get('WG_WARM', '3'))): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = int(os.environ.get('WG_REPS', '10'))
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * N * Ho * Wo * cout * cin * k * k
    print('%-9s %7.3f ms  %6.1f TFLOP/s (fp32-equivalent)  scratch %.1f MB' % (name, ms, fl / ms / 1e9, n * 4 / 1e6))
    d = dbg.tolist()
    print('   converter: total %d, wait ring-empty %d, wait dO-empty %d, loads+index %d, chunks %d | MMA: total %d, wait acc_free %d, wait full %d, issue %d, chunks %d | drain: total %d, wait %d, segments %d' % (d[0], d[1], d[2], d[3], d[4], d[8], d[9], d[10], d[11], d[12], d[5], d[6], d[7]))
