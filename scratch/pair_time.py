"""GPU: time the stem pair (In_conv1 -> Z -> In_conv3) alone at the bench batch and the AR step with / without it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from arpde_b200 import engine, _lib as L
dev = torch.device('cuda:0')
N = int(os.environ.get('N', '1920'))
x = torch.randn(N, 6, 64, 64, device=dev)
w1 = torch.randn(96, 6, 5, 5, device=dev) / 12; w3 = torch.randn(48, 96, 3, 3, device=dev) / 30
sc = torch.rand(96, device=dev) + 0.5; sh = torch.randn(96, device=dev) * 0.2
y = torch.empty(N, 64, 32, 32, device=dev)
zf = int(L.lib.arpde_conv_z_floats(96, 64, 64)); z = torch.empty(N * zf, device=dev)
ns = int(L.lib.arpde_conv_fwd_tc_scratch_floats(6, 96, 5, 5, 1)) + int(L.lib.arpde_conv_fwd_tc_scratch_floats(96, 48, 3, 3, 1))
scratch = torch.empty(ns, device=dev)
def pair():
    L.call('arpde_stem_pair_fwd', L.ptr(x), x.stride(0), L.ptr(w1), 0, L.ptr(sc), L.ptr(sh), 0, L.ptr(w3), 0, L.ptr(y), N, 6, 64, 64, 96, 5, 48, 64, 0, N,
           L.ptr(z), z.numel(), L.ptr(scratch), scratch.numel(), L.stream())
for _ in range(3): pair()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): pair()
e1.record(); torch.cuda.synchronize()
print('pair (In_conv1 zout + conv_z + 2 preps): %.3f ms' % (e0.elapsed_time(e1) / 10))
if N == 1920:
    swag, _ = bench.build_swag(dev)
    named = list(swag.base.named_parameters())
    gen = torch.Generator(device=dev); gen.manual_seed(1234)
    ic = bench.make_ics(64).to(dev)
    flat, _, _ = swag.sample_flat(30, generator=gen)
    buf = torch.empty((1920, 6 + 2 * 25, 64, 64), device=dev)
    for lw in (False, True, False):
        engine.rollout(swag, ic, 25, flat_weights=flat, named_params=named, out=buf, layerwise=lw)
        torch.cuda.synchronize(); e0.record()
        engine.rollout(swag, ic, 25, flat_weights=flat, named_params=named, out=buf, layerwise=lw)
        e1.record(); torch.cuda.synchronize()
        print('layerwise=%s: %.3f ms/step' % (lw, e0.elapsed_time(e1) / 25))
