"""GPU: pure-write, pure-read and copy bandwidth of HBM on this B200 (torch kernels, 3.2 GB buffers, CUDA events)."""
import torch
dev = 'cuda:0'
n = 800 * 1024 * 1024          # 3.2 GB of fp32
a = torch.empty(n, device=dev); b = torch.empty(n, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: a.fill_(1.0)); print('write-only  fill_   : %.3f ms  %.0f GB/s' % (ms, n * 4 / ms / 1e6))
ms = t(lambda: a.zero_()); print('write-only  zero_   : %.3f ms  %.0f GB/s' % (ms, n * 4 / ms / 1e6))
ms = t(lambda: torch.cuda.memset(a.data_ptr(), 0, n * 4) if hasattr(torch.cuda, 'memset') else a.zero_())
ms = t(lambda: a.sum()); print('read-only   sum     : %.3f ms  %.0f GB/s' % (ms, n * 4 / ms / 1e6))
ms = t(lambda: b.copy_(a)); print('copy (r + w) copy_  : %.3f ms  %.0f GB/s (both directions counted)' % (ms, 2 * n * 4 / ms / 1e6))
