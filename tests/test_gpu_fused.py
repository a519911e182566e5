"""Fused dense-block kernel (csrc/dense_block.cu: all 4-channel dense layers of a block in one launch, eval mode, 32-column maps) against
the layer-by-layer path of the same library and against the CPU oracle of DenseED.forward (2D-Burgers-SWAG/nn/denseEDcirc2d.py:34-79).
The kernel is only dispatched once the grid fills the machine (N * H/16 >= number of SMs), so these tests use >= 80 samples."""
import contextlib
import io

import pytest
import torch

from _util import golden, rel_err, sub
import _models as M
from oracle import arpde_oracle as O

from arpde_b200 import engine

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def _model(blocks, seed=0, init_features=96):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = M.CLS['burgers2d'](6, 2, blocks, growth_rate=4, init_features=init_features).to(DEV)
    with torch.no_grad():
        for mod in m.modules():
            if hasattr(mod, 'running_var'):
                mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
    return m.eval()


@pytest.mark.parametrize('blocks,init_features,N', [([4], 96, 80), ([2], 96, 150), ([3], 64, 96)])
def test_fused_dense_block_forward_vs_oracle(blocks, init_features, N):
    m = _model(blocks, init_features=init_features)
    x = torch.randn(N, 6, 64, 64, device=DEV)
    with torch.no_grad():
        y = m(x)
    sd = {'features.' + k: v.detach().cpu() for k, v in m.features.state_dict().items()}
    idx = [0, 1, N // 2, N - 1]                              # eval-mode BatchNorm: samples are independent
    ref = O.densed_forward('burgers2d', sd, x[idx].cpu(), blocks, training=False)
    assert rel_err(y[idx], ref) < TOL


def test_fused_rollout_equals_layerwise_with_weight_groups():
    """S weight samples x N ICs (sample-major groups): the fused kernel picks each sample's weight group and folded BatchNorm"""
    g, cfg, args, model, bayes, swag = M.build('burgers2d', DEV, max_models=int(golden('burgers2d')['swag/max_models']))
    M.load_golden_weights(bayes, g)
    swag.load_state_dict({k: v for k, v in sub(g, 'swag/sd/').items()}, strict=True)
    swag.eval()
    S, N, T = 3, 30, 3
    ic = O.ic_burgers2d(64, range(N)).to(DEV)
    gen = torch.Generator(device=DEV); gen.manual_seed(2)
    flat, names, numels = swag.sample_flat(S, generator=gen)
    named = list(bayes.named_parameters())
    a = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named)
    b = engine.rollout(swag, ic, T, flat_weights=flat, named_params=named, layerwise=True)
    errs = [rel_err(a[:, t], b[:, t]) for t in range(T + 1)]
    print('fused vs layerwise per-step rel err', ['%.1e' % e for e in errs])
    assert errs[0] == 0.0 and errs[1] < TOL and max(errs) < 1e-4
    # continued rollout (pieces) == one-shot rollout
    buf = torch.empty((S * N, 6 + 2 * T, 64, 64), device=DEV)
    engine.rollout(swag, ic, 2, flat_weights=flat, named_params=named, out=buf, t_start=0, total_steps=T)
    c = engine.rollout(swag, ic, 1, flat_weights=flat, named_params=named, out=buf, t_start=2, total_steps=T)
    assert c.shape == a.shape and rel_err(c, a) < 1e-6


@pytest.mark.parametrize('cin,H,N,npg', [(16, 64, 3, 3), (4, 16, 2, 1), (12, 80, 2, 2), (8, 32, 5, 1)])
def test_output_conv_kernel_vs_oracle(cin, H, N, npg):
    """conv_out.cu through the C ABI (arpde_conv_fwd): 5x5 circular conv to 2 channels on 64-column maps with the BN + ReLU prologue,
    input as a channel-slice view, output written at a channel offset of a wider tensor (the rollout's trajectory), per-group weights,
    tanh output activation."""
    from arpde_b200 import _lib as L
    torch.manual_seed(cin + H)
    W = 64
    groups = N // npg
    xbuf = torch.randn(N, cin + 3, H, W, device=DEV)
    x = xbuf[:, 3:]
    w = torch.randn(groups, 2, cin, 5, 5, device=DEV) / (cin * 25) ** 0.5
    sc = torch.rand(groups, cin, device=DEV) + 0.5
    sh = torch.randn(groups, cin, device=DEV) * 0.3
    ybuf = torch.full((N, 7, H, W), 9.0, device=DEV)
    L.call('arpde_conv_fwd', L.ptr(x), (cin + 3) * H * W, L.ptr(w), 2 * cin * 25, L.ptr(sc), L.ptr(sh), cin, L.ptr(ybuf), 0,
           N, cin, H, W, 2, 5, 5, 1, 0, 1, 1, 7, 4, npg, L.stream())
    torch.cuda.synchronize()
    for n in range(N):
        g = n // npg
        a = torch.relu(x[n:n + 1].double().cpu() * sc[g].double().cpu().view(1, -1, 1, 1) + sh[g].double().cpu().view(1, -1, 1, 1))
        ref = torch.tanh(O.conv_circ(a, w[g].double().cpu()))
        assert rel_err(ybuf[n:n + 1, 4:6], ref) < TOL, n
    assert float((ybuf[:, :4] - 9.0).abs().max()) == 0.0 and float((ybuf[:, 6:] - 9.0).abs().max()) == 0.0      # neighbours untouched


def test_launch_counter_counts_the_rollout_launches():
    """arpde_launch_count(): a T-step rollout of the default 2D net at a machine-filling batch is 6 launches per step (In_conv1 as
    conv_stem_kernel, In_conv3 as conv_z_kernel, ONE fused dense block, conv1x1 and the folded up-conv as conv_h_kernel, output conv) + 13
    one-off launches (8 TF32 weight splits -- every layer the per-layer path could run on the tensor core --, the 2 fp16 weight
    preparations of the stem pair, the 2 of conv_h, 1 BatchNorm fold); the per-layer path: 9 per step + 9 (8 weight splits + the fold)."""
    from arpde_b200 import _lib as L
    m = _model([4])
    ic = torch.randn(100, 2, 64, 64, device=DEV)
    engine.rollout(m, ic, 1)
    counts = {}
    for lw in (False, True):
        for T in (1, 3):
            L.lib.arpde_launch_count(1)
            engine.rollout(m, ic, T, layerwise=lw)
            counts[lw, T] = int(L.lib.arpde_launch_count(1))
    print(counts)
    per_step = {lw: (counts[lw, 3] - counts[lw, 1]) // 2 for lw in (False, True)}
    once = {lw: counts[lw, 1] - per_step[lw] for lw in (False, True)}
    assert per_step[False] == 6 and once[False] == 13, counts
    assert per_step[True] == 9 and once[True] == 9, counts
