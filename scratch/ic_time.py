import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200.data import InitPeriodicCond2d
from oracle import arpde_oracle as O
for n in (64, 256):
    ds = InitPeriodicCond2d(n, 100000, device='cuda:0')
    idx = list(range(128))
    ds.batch(idx); torch.cuda.synchronize()
    t0 = time.time(); x = ds.batch(idx); torch.cuda.synchronize(); t1 = time.time()
    lam, c = ds.coefficients(idx)
    t2 = time.time(); ds.coefficients(idx); t3 = time.time()
    t4 = time.time(); O.ic_burgers2d(n, idx[:8]); t5 = time.time()
    print('ncells %d: 128 ICs on device %.2f ms (of which host RNG %.2f ms); numpy reference path %.1f ms per sample' % (n, (t1 - t0) * 1e3, (t3 - t2) * 1e3, (t5 - t4) / 8 * 1e3))
