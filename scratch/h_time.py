"""GPU: conv1x1 (64 -> 32, 32x32) and the upsampling conv3x3 (32 -> 16) at the bench batch on conv_tc (3xTF32) vs conv_h (fp16 two-term), and the AR step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import bench
from arpde_b200 import engine, _lib as L
dev = torch.device('cuda:0')
N = int(os.environ.get('N', '1920')); G = 30 if N % 30 == 0 else 1
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, nrep=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(nrep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep
for (cin, cout, k, up, H) in [(64, 32, 1, 0, 32), (32, 16, 3, 1, 32)]:
    xs = [torch.randn(N, cin, H, H, device=dev) for _ in range(3)]
    w = torch.randn(G, cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    sc = torch.rand(G, cin, device=dev) + 0.5; sh = torch.randn(G, cin, device=dev) * 0.2
    f = 2 if up else 1
    y = torch.empty(N, cout, H * f, H * f, device=dev); y2 = torch.empty_like(y)
    ns = max(int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, G)), int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, k, k, G)))
    scr = torch.empty(ns, device=dev)
    cnt = [0]
    def h():
        cnt[0] += 1
        L.call('arpde_conv_fwd_h', L.ptr(xs[cnt[0] % 3]), cin * H * H, L.ptr(w), w.stride(0), L.ptr(sc), L.ptr(sh), cin, L.ptr(y), N, cin, H, H, cout, k, up, 1,
               cout, 0, N // G, L.ptr(scr), ns, L.stream())
    def tc():
        cnt[0] += 1
        L.call('arpde_conv_fwd_tc', L.ptr(xs[cnt[0] % 3]), cin * H * H, L.ptr(w), w.stride(0), L.ptr(sc), L.ptr(sh), cin, L.ptr(y2), 0, N, cin, H, H, cout, k, k, 1, up,
               1, 0, cout, 0, N // G, L.ptr(scr), ns, L.stream())
    t_h = timeit(h); t_tc = timeit(tc)
    cnt[0] = 0; h(); cnt[0] = 0; tc(); torch.cuda.synchronize()
    byts = (xs[0].numel() + y.numel()) * 4
    print('cin %d cout %d k %d up %d: conv_h %.3f ms (%.0f GB/s), conv_tc %.3f ms, max diff %.2e (max |y| %.2f)' % (
        cin, cout, k, up, t_h, byts / t_h / 1e6, t_tc, float((y - y2).abs().max()), float(y2.abs().max())))
    del xs, y, y2
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
