"""C1: 1D Kuramoto-Sivashinsky SWAG rollout, 30 posterior samples x N ICs x 200 AR steps (96-point grid), through engine.swag_rollout."""
import sys, os, time, types, contextlib, io, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arpde_b200.nn.densed import DenseED1dKS as DenseED1d
from arpde_b200.nn.bayesnn import BayesNN
from arpde_b200.nn.swag import SwagNN
from arpde_b200 import engine
dev = torch.device('cuda:0')
S, NIC, T, L_ = 30, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 200, 96
torch.manual_seed(0); np.random.seed(0)
args = types.SimpleNamespace(dt=0.1, device=dev, nic=2)
with contextlib.redirect_stdout(io.StringIO()):
    model = DenseED1d(2, 1, [4], growth_rate=4, init_features=32).to(dev)
    bayes = BayesNN(args, model); swag = SwagNN(args, bayes, full_cov=True, max_models=30)
with torch.no_grad():
    base = [p.detach().clone() for p in bayes.parameters()]
    for it in range(30):
        for p, b in zip(bayes.parameters(), base):
            p.copy_(b + 0.002 * torch.randn_like(b) * (b.abs().mean() + 1e-3))
        swag.collect()
swag.eval()
ic = torch.randn(NIC, 1, L_, device=dev) * 0.1
g = torch.Generator(device=dev); g.manual_seed(1)
def job():
    return engine.swag_rollout(swag, ic, T, S, generator=g, formula='1d')
for _ in range(2): job()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); n = 3
for _ in range(n): out = job()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print('KS C1: %d samples x %d ICs x %d steps: %.1f ms per job, %.1f us per AR step, %.0f sample-steps/s, finite=%s' % (S, NIC, T, ms, ms * 1e3 / T, S * NIC * T / ms * 1e3, bool(torch.isfinite(out[1]).all())))
