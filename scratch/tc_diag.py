"""Per-layer timing (FFMA vs tensor-core kernel) and error vs the fp64 oracle for the default 2D network layers."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from _util import rel_err
from oracle import arpde_oracle as O
from arpde_b200 import _lib as L
DEV = 'cuda:0'
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 1920
FLAGS = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L.lib.arpde_conv_tc_set_debug_flags(FLAGS)
LAYERS = [('In_conv1', 6, 96, 5, 1, 0, 64, 0), ('In_conv3', 96, 48, 3, 2, 0, 64, 0), ('dense1', 48, 4, 3, 1, 0, 32, 1), ('dense4', 60, 4, 3, 1, 0, 32, 1),
          ('conv1x1', 64, 32, 1, 1, 0, 32, 1), ('conv2up', 32, 16, 3, 1, 1, 32, 1), ('conv3', 16, 2, 5, 1, 0, 64, 1)]

def run(fn, x, w, sc, sh, y, N, cin, H, cout, k, s, up, relu, scratch):
    a = [L.ptr(x), cin * H * H, L.ptr(w), 0, L.ptr(sc) if relu else 0, L.ptr(sh) if relu else 0, 0, L.ptr(y), 0, N, cin, H, H, cout, k, k, s, up, relu, 0, cout, 0, N]
    if fn == 'arpde_conv_fwd_tc':
        a += [L.ptr(scratch), scratch.numel()]
    L.call(fn, *a, L.stream())

for name, cin, cout, k, s, up, H, relu in LAYERS:
    torch.manual_seed(0)
    Ho = (H << up) // s
    w = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).to(DEV)
    sc = (torch.rand(cin) + 0.5).to(DEV); sh = (torch.randn(cin) * 0.3).to(DEV)
    scratch = torch.empty(int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, k, k, 1)), device=DEV)
    # accuracy on a small batch
    xs = torch.randn(2, cin, H, H, device=DEV)
    a = xs.cpu().double()
    if relu: a = torch.relu(a * sc.cpu().double().view(1, -1, 1, 1) + sh.cpu().double().view(1, -1, 1, 1))
    if up: a = O._up2(a)
    ref = O.conv_circ(a, w.cpu().double(), stride=s)
    errs = {}
    for fn in ('arpde_conv_fwd', 'arpde_conv_fwd_tc'):
        y = torch.empty(2, cout, Ho, Ho, device=DEV)
        run(fn, xs, w, sc, sh, y, 2, cin, H, cout, k, s, up, relu, scratch)
        torch.cuda.synchronize()
        errs[fn] = rel_err(y, ref)
    # timing on the rollout batch
    x = torch.randn(NB, cin, H, H, device=DEV)
    y = torch.empty(NB, cout, Ho, Ho, device=DEV)
    ms = {}
    for fn in ('arpde_conv_fwd', 'arpde_conv_fwd_tc'):
        for _ in range(2): run(fn, x, w, sc, sh, y, NB, cin, H, cout, k, s, up, relu, scratch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(fn, x, w, sc, sh, y, NB, cin, H, cout, k, s, up, relu, scratch)
        e1.record(); torch.cuda.synchronize()
        ms[fn] = e0.elapsed_time(e1) / 5
    fl = 2.0 * NB * Ho * Ho * cout * cin * k * k
    import ctypes
    dbg = torch.zeros(16, dtype=torch.int64, device=DEV)
    L.lib.arpde_conv_tc_set_debug.argtypes = [ctypes.c_void_p]
    L.lib.arpde_conv_tc_set_debug(dbg.data_ptr())
    run('arpde_conv_fwd_tc', x, w, sc, sh, y, NB, cin, H, cout, k, s, up, relu, scratch); torch.cuda.synchronize()
    L.lib.arpde_conv_tc_set_debug(0)
    d = dbg.cpu().tolist(); it = max(d[12], 1)
    print('   per item (clk): stager total %d wait_empty %d store %d fence %d load_issue %d enter %d | epi total %d wait_full %d tmem_ld %d | issuer total %d wait_tmem_empty %d wait_full_p %d wait_w %d issue_block %d (items %d)' % (
        d[0] // it, d[1] // it, d[2] // it, d[3] // it, d[13] // it, d[14] // it, d[4] // it, d[5] // it, d[6] // it, d[8] // it, d[9] // it, d[10] // it, d[11] // it, d[15] // it, d[12]))
    print('%-9s N=%d  ffma %.3f ms (%.1f TF/s) err %.2e | tc %.3f ms (%.1f TF/s) err %.2e' % (
        name, NB, ms['arpde_conv_fwd'], fl / ms['arpde_conv_fwd'] / 1e9, errs['arpde_conv_fwd'],
        ms['arpde_conv_fwd_tc'], fl / ms['arpde_conv_fwd_tc'] / 1e9, errs['arpde_conv_fwd_tc']), flush=True)
