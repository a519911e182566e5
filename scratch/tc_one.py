"""Time one layer of the default net on the tensor-core kernel (for planner A/B experiments).  usage: tc_one.py <layer idx> [N]"""
import sys, os, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
DEV = 'cuda:0'
LAYERS = [('In_conv1', 6, 96, 5, 1, 0, 64, 0), ('In_conv3', 96, 48, 3, 2, 0, 64, 0), ('dense1', 48, 4, 3, 1, 0, 32, 1), ('dense4', 60, 4, 3, 1, 0, 32, 1),
          ('conv1x1', 64, 32, 1, 1, 0, 32, 1), ('conv2up', 32, 16, 3, 1, 1, 32, 1), ('conv3', 16, 2, 5, 1, 0, 64, 1),
          ('wide144', 144, 32, 3, 1, 0, 16, 1), ('wide_in3_128', 96, 48, 3, 2, 0, 128, 1), ('dgrad_like', 48, 96, 3, 1, 0, 32, 0)]
name, cin, cout, k, s, up, H, relu = LAYERS[int(sys.argv[1])]
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
out = (ctypes.c_int32 * 160)()
L.call('arpde_conv_tc_plan', cin, H, H, cout, k, k, s, up, ctypes.addressof(out))
names = 'pitch min_ry min_rx ntile ctas PP PHP nplanes cchunk ncc tchunk ntc ntaps coutP tmem nslot_p smem Ho Wo Hc Wc sy sx wanted resident nacc kp'.split()
p = dict(zip(names, out[:27]))
torch.manual_seed(0)
Ho = (H << up) // s
w = (torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5).to(DEV)
sc = (torch.rand(cin) + 0.5).to(DEV); sh = (torch.randn(cin) * 0.3).to(DEV)
scratch = torch.empty(int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, k, k, 1)), device=DEV)
x = torch.randn(NB, cin, H, H, device=DEV); y = torch.empty(NB, cout, Ho, Ho, device=DEV)
def run():
    L.call('arpde_conv_fwd_tc', L.ptr(x), cin * H * H, L.ptr(w), 0, L.ptr(sc) if relu else 0, L.ptr(sh) if relu else 0, 0, L.ptr(y), 0, NB, cin, H, H, cout, k, k, s, up, relu, 0, cout, 0, NB, L.ptr(scratch), scratch.numel(), L.stream())
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print('%-9s %.3f ms  plan: %s' % (name, e0.elapsed_time(e1) / 5, {k_: p[k_] for k_ in 'ntile nslot_p cchunk ncc tchunk resident nacc smem kp'.split()}))
