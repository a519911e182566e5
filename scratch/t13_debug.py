"""GPU debug: T13 1D Burgers -- where do the drop-in and the reference part ways?"""
import os, sys, json, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import make_t13 as T13
for system in ('burgers1d',):
    a = T13.run(system, 'dropin'); b = T13.run(system, 'ref')
    print(system, 'dropin  ', json.dumps(a)); print(system, 'ref cuda', json.dumps(b))
# module-level: same weights, train-mode forwards move the BN statistics, then eval forward
from oracle import ref_shims
import arpde_b200
from arpde_b200.nn import densed as D
torch.backends.cudnn.allow_tf32 = False
dev = 'cuda:0'
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    ours = D.DenseED1dBurgers(5, 1, [1], growth_rate=4, init_features=64).to(dev)
    with ref_shims.reference_modules('burgers1d'):
        from nn.denseEDcirc import DenseED as Ref
        ref = Ref(5, 1, [1], growth_rate=4, init_features=64).to(dev)
ref.load_state_dict(ours.state_dict())
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
for B in (8, 5):
    ours.train(); ref.train()
    for it in range(10):
        x = torch.randn(B, 20, 64, device=dev)[:, -5:]
        ya, yb = ours(x), ref(x)
        if it in (0, 9): print('B', B, 'train fwd', it, rel(ya, yb))
    sa, sb = ours.state_dict(), ref.state_dict()
    print(' max running-stat rel err', max(rel(sa[k].float(), sb[k].float()) for k in sa if 'running' in k), [int(sa[k]) for k in sa if 'num_batches' in k][:3], [int(sb[k]) for k in sb if 'num_batches' in k][:3])
    ours.eval(); ref.eval()
    with torch.no_grad():
        inp = torch.randn(5, 20, 64, device=dev)
        ia, ib = inp.clone(), inp.clone()
        for t in range(4):
            ua, ub = ours(ia[:, -5:]), ref(ib[:, -5:])
            print(' eval step', t, rel(ua, ub), tuple(ua.shape))
            ia = torch.cat([ia[:, -4:].detach(), ua[:, 0].unsqueeze(1).detach()], 1)
            ib = torch.cat([ib[:, -4:].detach(), ub[:, 0].unsqueeze(1).detach()], 1)
