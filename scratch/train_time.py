"""C3-style training step timing: B=128 random Fourier ICs, tback=2 window (2 AR steps fwd + CN + loss, backward, Adam)."""
import sys, os, types, time, contextlib, io, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + '/tests')
import bench as Bn
from arpde_b200.nn.densed import DenseED2d
from arpde_b200.nn.bayesnn import BayesNN
from arpde_b200.nn.swag import SwagNN
from arpde_b200.nn.integrators import Burger2DIntegrate
dev = torch.device('cuda:0')
torch.manual_seed(0); np.random.seed(0)
args = types.SimpleNamespace(dt=0.005, device=dev, nic=3)
with contextlib.redirect_stdout(io.StringIO()):
    model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(dev)
    bayes = BayesNN(args, model); swag = SwagNN(args, bayes, full_cov=True, max_models=30)
    integ = Burger2DIntegrate(1.0 / 64, nu=0.005, grad_kernels=[3, 3], device=dev)
from arpde_b200.optim import Adam as _A
opt = (torch.optim.Adam if os.environ.get('TORCH_ADAM') else _A)([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ic = Bn.make_ics(B).to(dev)
swag.train()
def step():
    inp = ic.repeat(1, 3, 1, 1)
    loss = 0
    for i in range(2):
        uPred = swag(inp[:, -6:])
        ustar = integ.crankNicolson(uPred, inp[:, -2:], 0.003)
        loss = loss + swag.calc_neg_log_joint(uPred, ustar, 10)
        inp = torch.cat([inp, uPred], 1)
    loss.backward(); opt.step(); opt.zero_grad()
NW = int(os.environ.get("NW", "3")); NS = int(os.environ.get("NS", "10"))
for _ in range(NW): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NS): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / NS
print('train step B=%d, 2 AR steps: %.2f ms -> %.0f sample-steps/s (TC=%s)' % (B, ms, 2 * B / ms * 1e3, os.environ.get('ARPDE_TC', '1')))
# forward-only share
torch.cuda.synchronize(); e0.record()
with torch.no_grad():
    for _ in range(10):
        inp = ic.repeat(1, 3, 1, 1)
        for i in range(2):
            uPred = swag(inp[:, -6:]); inp = torch.cat([inp, uPred], 1)
e1.record(); torch.cuda.synchronize()
print('  forward only (train-mode BN, 2 AR steps): %.2f ms' % (e0.elapsed_time(e1) / 10))
