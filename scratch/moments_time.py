"""GPU: predictive-moments kernel on the C4 trajectory tensor (30 x 64 trajectories, 101 saved states of 2 x 64 x 64)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import engine
dev = 'cuda:0'
S, N = 30, 64
traj = torch.randn(S * N, 206, 64, 64, device=dev)
betas = torch.rand(S, device=dev) + 1
for form in ('2d', '1d'):
    for _ in range(2): m, v = engine.predictive_moments(traj[:, 4:], S, betas, form)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): m, v = engine.predictive_moments(traj[:, 4:], S, betas, form)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byts = traj[:, 4:].numel() * 4 + 2 * m.numel() * 4
    print('moments %s: %.3f ms, %.0f GB/s' % (form, ms, byts / ms / 1e6))
y = traj[:, 4:].reshape(S, N, -1)[:, :2, :4096].double()
mm = y.mean(0); print('check mean', float((m.reshape(N, -1)[:2, :4096].double() - mm).abs().max()))
