"""K2 timing on the C4 rollout batch (1920 x 2 x 64 x 64, channel-slice views of a trajectory tensor) and on C5 (256^2)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
dev = 'cuda:0'
for (N, H, CT) in ((1920, 64, 26), (256, 256, 8)):
    traj = torch.randn(N, CT, H, H, device=dev)
    up, u0 = traj[:, 6:8], traj[:, 4:6]
    ustar = torch.empty(N, 2, H, H, device=dev)
    sse = torch.zeros(1, dtype=torch.float64, device=dev)
    bs = traj.stride(0)
    for name, out, s in (('api form (ustar)', ustar, None), ('fused loss (sse only)', None, sse), ('both', ustar, sse)):
        def k2():
            L.call('arpde_cn_burgers2d_fwd', L.ptr(up), bs, L.ptr(u0), bs, L.ptr(out), L.ptr(s), N, H, H, 1.0 / H, 0.005, 0.005, 0, L.stream())
        for _ in range(3): k2()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): k2()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        bpp = 24 if out is not None else 16
        print('N=%d %dx%d %-22s %.4f ms  %.0f GB/s algorithmic (%d B/pt)' % (N, H, H, name, ms, N * H * H * bpp / ms / 1e6, bpp), flush=True)
