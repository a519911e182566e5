"""GPU: is the training step capturable in a CUDA graph?  (forward + Crank-Nicolson + neg log joint + backward into the flat bucket)
Compares eager and replayed gradients and times both, for C1 (1D Burgers, B=256) and C3 (2D Burgers, B=128)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from arpde_b200.nn.densed import DenseED1dBurgers, DenseED2d
from arpde_b200.nn.bayesnn import BayesNN
from arpde_b200.nn.swag import SwagNN
from arpde_b200.nn.integrators import BurgerIntegrate, Burger2DIntegrate
from arpde_b200 import parallel as PL
dev = torch.device('cuda:0')

def build(which):
    torch.manual_seed(0); np.random.seed(0)
    if which == 'c1':
        nic, B = 5, 256
        a = types.SimpleNamespace(dt=0.005, device=dev, nic=nic)
        with bench._quiet():
            model = DenseED1dBurgers(nic, 1, [1], growth_rate=4, init_features=64).to(dev)
            bayes = BayesNN(a, model); swag = SwagNN(a, bayes, full_cov=True, max_models=30)
            integ = BurgerIntegrate(1.0 / 513, nu=0.0025, grad_kernels=[3, 3], device=dev)
        ic = torch.randn(B, 1, 512, device=dev)
        feed = lambda inp: inp[:, -nic:]; last = lambda inp: inp[:, -1:]; app = lambda inp, u: torch.cat([inp, u[:, :1]], 1); rep = (1, nic, 1)
    else:
        B = 128
        a = types.SimpleNamespace(dt=0.005, device=dev, nic=3)
        with bench._quiet():
            model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(dev)
            bayes = BayesNN(a, model); swag = SwagNN(a, bayes, full_cov=True, max_models=30)
            integ = Burger2DIntegrate(1.0 / 64, nu=0.005, grad_kernels=[3, 3], device=dev)
        ic = bench.make_ics(B).to(dev)
        feed = lambda inp: inp[:, -6:]; last = lambda inp: inp[:, -2:]; app = lambda inp, u: torch.cat([inp, u], 1); rep = (1, 3, 1, 1)
    swag.train()
    bucket = PL.FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))
    def fb():
        inp = ic.repeat(*rep)
        loss = 0
        bucket.zero_()
        for _ in range(2):
            u = swag(feed(inp))
            ustar = integ.crankNicolson(u, last(inp), 0.003)
            loss = loss + swag.calc_neg_log_joint(u, ustar, 10)
            inp = app(inp, u)
        with bucket.capture():
            loss.backward()
        return loss.detach()
    return fb, bucket

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for which in ('c1', 'c3'):
    fb, bucket = build(which)
    fb(); flat = bucket.flat
    fb(); torch.cuda.synchronize()
    g_eager = flat.clone()
    t_eager = timeit(fb)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fb()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            loss = fb()
    except Exception as e:
        print(which, 'capture failed:', repr(e)[:500]); continue
    graph.replay(); torch.cuda.synchronize()
    err = float((flat - g_eager).abs().max() / g_eager.abs().max())
    t_graph = timeit(graph.replay)
    print('%s: eager %.3f ms, graph replay %.3f ms, max gradient difference (relative to max) %.2e, loss %s' % (which, t_eager, t_graph, err, loss.cpu().numpy()))
