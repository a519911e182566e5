"""GPU: small invocations of the round-2 kernels for `compute-sanitizer --tool memcheck` (one tool, one call)."""
import os, sys, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import _models as M
from _util import golden, sub
from arpde_b200 import engine, solver
from arpde_b200 import _lib as L
dev = 'cuda:0'
# KS persistent rollout: 2 weight groups x 5 ICs (a batch spans both groups), 3 steps
with contextlib.redirect_stdout(io.StringIO()):
    g, cfg, args, model, bayes, swag = M.build('ks1d', dev, max_models=3)
M.load_golden_weights(bayes, g)
for _ in range(3): swag.collect()
swag.eval()
flat, _, _ = swag.sample_flat(2)
ic = torch.randn(5, 1, 96, device=dev)
t = engine.rollout(swag, ic, 3, flat_weights=flat, named_params=list(bayes.named_parameters()))
print('ks', bool(torch.isfinite(t).all()))
# 2D default net, 80 samples of 64x64 (fused dense block streaming kernel + output conv), 2 groups
with contextlib.redirect_stdout(io.StringIO()):
    g2, cfg2, args2, model2, bayes2, swag2 = M.build('burgers2d', dev, max_models=3)
M.load_golden_weights(bayes2, g2)
for _ in range(3): swag2.collect()
swag2.eval()
flat2, _, _ = swag2.sample_flat(2)
ic2 = torch.randn(40, 2, 64, 64, device=dev)
t2 = engine.rollout(swag2, ic2, 1, flat_weights=flat2, named_params=list(bayes2.named_parameters()))
print('2d stream', bool(torch.isfinite(t2).all()))
# tiled dense block (3-layer block -> tiled kernel), 96 samples
with contextlib.redirect_stdout(io.StringIO()):
    m3 = M.CLS['burgers2d'](6, 2, [3], growth_rate=4, init_features=64).to(dev).eval()
with torch.no_grad():
    y3 = m3(torch.randn(96, 6, 64, 64, device=dev))
print('2d tiled', bool(torch.isfinite(y3).all()))
# stencil adjoint (tiled) at 64x64 and 256x256, strided views
for N, n in ((3, 64), (2, 256), (2, 24)):
    hist = torch.randn(N, 6, n, n, device=dev); gg = torch.randn(N, 2, n, n, device=dev)
    ga = torch.empty(N, 2, n, n, device=dev); gb = torch.empty_like(ga)
    for mode in (0, 1):
        L.call('arpde_cn_burgers2d_bwd', L.ptr(hist[:, 4:6]), hist.stride(0), L.ptr(hist[:, 2:4]), hist.stride(0), L.ptr(gg), L.ptr(ga), L.ptr(gb), N, n, n, 1.0 / n, 0.005, 0.003, mode, L.stream())
torch.cuda.synchronize(); print('stencil bwd ok')
# stencil forward (bulk-copy ring kernels incl. EDGE variant, ragged strips, many items per CTA), API form + fused loss
for N, H, W in ((3, 64, 64), (2, 256, 256), (400, 16, 16), (3, 18, 64), (2, 10, 384), (1, 8, 1024), (2, 17, 23)):
    hist = torch.randn(N, 6, H, W, device=dev); up = torch.randn(N, 2, H, W, device=dev)
    out = torch.empty(N, 2, H, W, device=dev); sse = torch.zeros((), device=dev, dtype=torch.float64)
    for mode in (0, 1):
        L.call('arpde_cn_burgers2d_fwd', L.ptr(up), 2 * H * W, L.ptr(hist[:, -2:]), 6 * H * W, L.ptr(out), L.ptr(sse), N, H, W, 1.0 / W, 0.005, 0.004, mode, L.stream())
        L.call('arpde_cn_burgers2d_fwd', L.ptr(up), 2 * H * W, L.ptr(hist[:, -2:]), 6 * H * W, 0, L.ptr(sse), N, H, W, 1.0 / W, 0.005, 0.004, mode, L.stream())
        ga = torch.empty(N, 2, H, W, device=dev); gb = torch.empty_like(ga)
        L.call('arpde_cn_burgers2d_bwd', L.ptr(up), 2 * H * W, L.ptr(hist[:, -2:]), 6 * H * W, L.ptr(out), L.ptr(ga), L.ptr(gb), N, H, W, 1.0 / W, 0.005, 0.003, mode, L.stream())
torch.cuda.synchronize(); print('stencil fwd ok')
# FD solver
u = solver.fd_burgers1d(torch.randn(2, 128, device=dev, dtype=torch.float64), 1e-4, 20, save_every=5)
u2 = solver.fd_burgers1d(torch.randn(1, 2048, device=dev, dtype=torch.float64) * 0.1, 1e-5, 5, grad_order=(2, 4))
print('fd', bool(torch.isfinite(u).all()), bool(torch.isfinite(u2).all()))
torch.cuda.synchronize()
print('done')
