"""GPU: a 3-step C4-batch rollout (1920 trajectories) -- target of the ncu launch list / full captures."""
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
ic = bench.make_ics(int(os.environ.get("NIC", "64"))).to(dev)
flat, _, _ = swag.sample_flat(30, generator=gen)
T = int(os.environ.get('T', '3'))
traj = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named)
torch.cuda.synchronize()
print('ok', float(traj.abs().max()))
