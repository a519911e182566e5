import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
dev = 'cuda:0'
N, H, CT = 1920, 64, 26
traj = torch.randn(N, CT, H, H, device=dev)
up, u0 = traj[:, 6:8], traj[:, 4:6]
ustar = torch.empty(N, 2, H, H, device=dev)
bs = traj.stride(0)
for _ in range(4):
    L.call('arpde_cn_burgers2d_fwd', L.ptr(up), bs, L.ptr(u0), bs, L.ptr(ustar), 0, N, H, H, 1.0 / H, 0.005, 0.005, 0, L.stream())
torch.cuda.synchronize()
print('ok')
