"""K2 on the bulk-copy ring kernels: rotating buffer sets, C4 batch / C5 / training batch; sweeps TY and NS in a dev build."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200 import _lib as L
dev = 'cuda:0'
BPP = {'fwd': 24, 'fwd_sse': 16, 'bwd': 40}

def run(tag, B, n, names=('fwd', 'fwd_sse', 'bwd'), nsets=4, reps=20):
    sets = [[torch.randn((B, 2, n, n), device=dev) for _ in range(3)] + [torch.empty((B, 2, n, n), device=dev) for _ in range(3)] for _ in range(nsets)]
    sse = torch.zeros((nsets,), device=dev, dtype=torch.float64)
    bs = 2 * n * n
    def fwd(i):
        s = sets[i % nsets]
        L.call('arpde_cn_burgers2d_fwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[3]), 0, B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())
    def fwd_sse(i):
        s = sets[i % nsets]
        L.call('arpde_cn_burgers2d_fwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, 0, sse.data_ptr() + 8 * (i % nsets), B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())
    def bwd(i):
        s = sets[i % nsets]
        L.call('arpde_cn_burgers2d_bwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[2]), L.ptr(s[4]), L.ptr(s[5]), B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())
    fns = {'fwd': fwd, 'fwd_sse': fwd_sse, 'bwd': bwd}
    res = []
    for name in names:
        fn = fns[name]
        for i in range(4): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps): fn(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append('%s %.1f us %.0f GB/s' % (name, ms * 1e3, B * n * n * BPP[name] / ms / 1e6))
    print('%-28s B=%d %dx%d | %s' % (tag, B, n, n, ' | '.join(res)), flush=True)

def setenv(**kw):
    for k in ('ARPDE_STENCIL_NOBULK', 'ARPDE_BULK_TY', 'ARPDE_BULK_NS'):
        os.environ.pop(k, None)
    for k, v in kw.items(): os.environ[k] = str(v)

def setenv(**kw):
    for k in ('ARPDE_STENCIL_NOBULK', 'ARPDE_BULK_TY', 'ARPDE_BULK_NS', 'ARPDE_BULK_W8'):
        os.environ.pop(k, None)
    for k, v in kw.items(): os.environ[k] = str(v)

for (B, n) in ((1920, 64), (256, 256), (128, 64), (240, 64), (64, 128)):
    setenv(ARPDE_STENCIL_NOBULK=1); run('old kernels', B, n)
    setenv(ARPDE_BULK_W8=1); run('bulk 8 warps x 4 rows', B, n)
    setenv(); run('bulk 16 warps x 2 rows', B, n)
for ty in (16, 32, 64):
    setenv(ARPDE_BULK_TY=ty); run('bulk16 TY=%d' % ty, 1920, 64, names=('fwd', 'fwd_sse'))
for ty in (8, 16, 32):
    setenv(ARPDE_BULK_TY=ty); run('bulk16 TY=%d' % ty, 256, 256, names=('fwd', 'fwd_sse'))
