"""GPU: small targets for ncu captures (python scratch/prof_targets.py <ks|stencil|train|c4step>)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device('cuda:0')
what = sys.argv[1]
if what == 'ks':
    print(json.dumps(bench.ks_probe(dev)))
elif what == 'stencil':
    import arpde_b200._lib as L
    for B, n in ((1920, 64), (256, 256)):
        s = [torch.randn((B, 2, n, n), device=dev) for _ in range(3)] + [torch.empty((B, 2, n, n), device=dev) for _ in range(3)]
        bs = 2 * n * n
        for _ in range(2):
            L.call('arpde_cn_burgers2d_fwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[3]), 0, B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())
            L.call('arpde_cn_burgers2d_bwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[2]), L.ptr(s[4]), L.ptr(s[5]), B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())
        torch.cuda.synchronize()
    print('ok')
elif what == 'train':
    print(json.dumps(bench.train_step_probe(dev, nwarm=2, nrep=1)))
