import torch
a = torch.load('/tmp/dbg_tc1.pt'); b = torch.load('/tmp/dbg_tc0.pt'); c = torch.load('/tmp/dbg_tc1j.pt')
def re(u, v): return float((u - v).abs().max() / v.abs().max())
for k in a:
    print(k, 'tc1 vs tc0 %.3e' % re(a[k], b[k]), ' tc1(junk alloc) vs tc0 %.3e' % re(c[k], b[k]))
d = (a['work'] - b['work']).abs()
idx = (d > 1e-3 * b['work'].abs().max()).nonzero().flatten()
print('work mismatches', idx.numel(), idx[:10].tolist(), idx[-5:].tolist() if idx.numel() else '', 'of', d.numel())
print('work changed by backward (tc1):', re(a['work_after'], a['work']))
