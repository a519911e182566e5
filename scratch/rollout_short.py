"""GPU: a short C4 rollout (NIC initial conditions x 30 weight samples, T steps) -- target of the ncu launch list / full captures; prints the
AR-step time (CUDA events) for n_streams = 1 and 2.  NIC=8 is the shard every rank runs when the 64-IC job is split over 8 GPUs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from arpde_b200 import engine
dev = torch.device('cuda:0')
swag, _ = bench.build_swag(dev)
named = list(swag.base.named_parameters())
gen = torch.Generator(device=dev); gen.manual_seed(1234)
nic = int(os.environ.get("NIC", "64"))
ic = bench.make_ics(nic).to(dev)
flat, _, _ = swag.sample_flat(30, generator=gen)
T = int(os.environ.get('T', '3'))
traj = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named)
torch.cuda.synchronize()
print('ok', float(traj.abs().max()))
if os.environ.get('TIME'):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    traj = torch.empty((30 * nic, 6 + 2 * T, 64, 64), device=dev)
    for ns in [int(v) for v in os.environ.get("NS", "1,2").split(",")]:
        for _ in range(2):
            engine.rollout(swag, ic, T, flat_weights=flat, named_params=named, out=traj, n_streams=ns)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            engine.rollout(swag, ic, T, flat_weights=flat, named_params=named, out=traj, n_streams=ns)
        e1.record(); torch.cuda.synchronize()
        print('NIC %d n_streams %d: %.4f ms per AR step (%d trajectories)' % (nic, ns, e0.elapsed_time(e1) / 3 / T, 30 * nic))
