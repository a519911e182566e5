"""GPU: is the C4 rollout linear in T?  (job 651 ms vs 100 x 5.85 ms AR step)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from arpde_b200 import engine
dev = torch.device('cuda:0')
swag, _ = bench.build_swag(dev)
named = list(swag.base.named_parameters())
gen = torch.Generator(device=dev); gen.manual_seed(1234)
NIC = int(os.environ.get('NIC', '64'))
ic = bench.make_ics(NIC).to(dev)
flat, _, _ = swag.sample_flat(30, generator=gen)
bufs = {T: torch.empty((30 * NIC, 6 + 2 * T, 64, 64), device=dev) for T in (5, 25, 50, 100)}
def t_rollout(T, layerwise=False, ns=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.rollout(swag, ic, T, flat_weights=flat, named_params=named, out=bufs[T], layerwise=layerwise, n_streams=ns); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for T in (5, 25, 50, 100): t_rollout(T)
for rep in range(2):
    for T in (5, 25, 50, 100):
        ms = t_rollout(T); print('T=%3d  %.1f ms  %.3f ms/step' % (T, ms, ms / T))
print('layerwise T=25: %.3f ms/step' % (t_rollout(25, True) / 25))
ref = bufs[50].clone()
for ns in (2, 3, 4):
    t_rollout(50, ns=ns); ms = t_rollout(50, ns=ns); print('n_streams=%d T=50: %.3f ms/step, max diff vs 1 stream %.1e' % (ns, ms / 50, float((bufs[50] - ref).abs().max())))
# the other parts of the job
torch.cuda.synchronize(); e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
e0.record(); flat2, _, _ = swag.sample_flat(30, generator=gen); e1.record()
traj = engine.rollout(swag, ic, 100, flat_weights=flat, named_params=named, out=bufs[100]); e2.record()
m, v = engine.predictive_moments(traj, 30, flat[:, 0].exp(), '2d'); e3.record(); torch.cuda.synchronize()
print('sample %.2f ms, rollout %.1f ms, moments %.2f ms' % (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3)))
print('traj finite', bool(torch.isfinite(traj).all()), 'absmax', float(traj.abs().max()))
