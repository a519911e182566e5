"""GPU parity of the rows SURVEY.md 8a marks "generic / not called by any driver" and of the large-grid configuration:
  P4  5x5 2D stencils, kernel_size=2 upwind gradient, WENO flux derivative + crankNicolsonLimiter  vs tests/golden/p4.npz (real reference)
  M4  encoder-decoder topology (TransDown k4/s2, TransUp) BACKWARD vs fp64 oracle autograd; dropout / bilinear upsampling /
      bottleneck dense layers vs the reference's own modules on the same device and RNG stream; forward_test in train mode
  C5  wide DenseED (growth 32) at the stated 256x256: forward vs the CPU oracle, per-layer weight / data gradients at 256x256
      including the layers whose tensor-core weight-gradient ring does not fit shared memory (generic-kernel fallback)
Tolerance: the north_star's 1e-5 relative (max|a-b|/max|b|) in fp32 unless stated at the assertion."""
import contextlib
import io

import numpy as np
import pytest
import torch

from _util import golden, rel_err
import _models as M
from oracle import arpde_oracle as O
from oracle import ref_shims

from arpde_b200 import _lib as L
from arpde_b200.nn import densed as D
from arpde_b200.nn import integrators as I

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


# ----------------------------------------------------------------------------- P4
def _p4_keys():
    g = golden('p4')
    return sorted(set(k.rsplit('/', 1)[0] for k in g if k.endswith('/ustar')))


@pytest.mark.parametrize('key', _p4_keys())
def test_p4_integrator_variants_vs_reference_fixture(key):
    g = golden('p4')
    dim, ks, meth = key.split('/')
    gk = [int(ks[1]), int(ks[2])]
    dx, nu, dt = float(g[dim + '/dx']), float(g[dim + '/nu']), float(g[dim + '/dt'])
    with quiet():
        integ = (I.Burger2DIntegrate if dim == '2d' else I.BurgerIntegrate)(dx, nu=nu, grad_kernels=gk, device=DEV)
    u = cu(g[dim + '/u']).requires_grad_(True)
    u0 = cu(g[dim + '/u0']).requires_grad_(True)
    us = getattr(integ, meth)(u, u0, dt)
    (us * cu(g[dim + '/gout'])).sum().backward()
    assert rel_err(us, g[key + '/ustar']) < TOL
    assert rel_err(u.grad, g[key + '/g_u']) < TOL
    assert rel_err(u0.grad if u0.grad is not None else torch.zeros_like(u0), g[key + '/g_u0']) < TOL


def test_p4_filter_objects_vs_reference_fixture():
    g = golden('p4')
    u = cu(g['1d/u'])
    dx = float(g['1d/dx'])
    assert rel_err(I.FluxWENOFilter1D(dx, device=DEV)(u), g['1d/weno']) < TOL
    assert rel_err(I.GradFilter1D(dx, kernel_size=2, device=DEV)(u), g['1d/upwind']) < TOL
    a = cu(g['2d/u'])[:, :1].contiguous()
    dx2 = float(g['2d/dx'])
    with quiet():
        s5, l5 = I.SobelFilter2d(dx2, kernel_size=5, device=DEV), I.LapLaceFilter2d(dx2, kernel_size=5, device=DEV)
    assert rel_err(s5.grad_h(a), g['2d/filt5/grad_h']) < TOL
    assert rel_err(s5.grad_v(a), g['2d/filt5/grad_v']) < TOL
    assert rel_err(l5(a), g['2d/filt5/lap']) < TOL
    with pytest.raises(RuntimeError):
        l5(a.cpu())                      # no CPU path, also on the roll path


# ----------------------------------------------------------------------------- M4
def _randomise_bn(m):
    with torch.no_grad():
        for mod in m.modules():
            if hasattr(mod, 'running_var'):
                mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)


@pytest.mark.parametrize('sysname,xshape', [('burgers2d', (3, 6, 32, 32)), ('ks1d', (4, 2, 64))])
@pytest.mark.parametrize('training', [True, False])
def test_m4_encoder_decoder_backward_vs_oracle(sysname, xshape, training):
    """blocks=(2,3,2): EncBlock, TransDown (1x1 + k4/s2 circular conv, pad 1), DecBlocks, TransUp (1x1 + nearest x2 + conv3), tanh output:
    gradient of every parameter and of the input against fp64 autograd through the oracle (k4/s2 data / weight gradients,
    zero-insertion path, output-activation backward)."""
    torch.manual_seed(3)
    with quiet():
        m = M.CLS[sysname](xshape[1], 1 if sysname == 'ks1d' else 2, [2, 3, 2], growth_rate=4, init_features=16, out_activation='tanh').to(DEV)
    _randomise_bn(m)
    m.train(training)
    x = torch.randn(*xshape, device=DEV)
    sd = {'features.' + k: v.detach().cpu().clone() for k, v in m.features.state_dict().items()}
    xg = x.clone().requires_grad_(True)
    y = m(xg)
    gy = torch.randn_like(y)
    (y * gy).sum().backward()
    sd64 = {k: (v.double().requires_grad_('running' not in k) if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    x64 = x.cpu().double().requires_grad_(True)
    ref = O.densed_forward(sysname, sd64, x64, [2, 3, 2], training=training, out_activation='tanh')
    (ref * gy.cpu().double()).sum().backward()
    assert rel_err(y, ref) < TOL
    # 5-layer-deep BPTT-free single pass; gradients through 14 BatchNorm layers: 2e-5 as for the 3-step BPTT gradients
    assert rel_err(xg.grad, x64.grad) < 2e-5
    for name, p_ in m.features.named_parameters():
        assert rel_err(p_.grad, sd64['features.' + name].grad) < 2e-5, name


def test_forward_test_in_train_mode_matches_forward():
    """DenseED.forward_test on a fresh (train-mode) model, as the reference's own __main__ runs it (denseEDcirc2d.py:316-323,355-):
    stand-alone sub-trees that start with BatchNorm compute the batch statistics of their input; outputs and running statistics equal
    those of the fused forward."""
    torch.manual_seed(4)
    with quiet():
        a = D.DenseED2d(6, 2, [2, 3, 2], growth_rate=4, init_features=16).to(DEV)
        b = D.DenseED2d(6, 2, [2, 3, 2], growth_rate=4, init_features=16).to(DEV)
    _randomise_bn(a)
    b.load_state_dict(a.state_dict())
    x = torch.randn(3, 6, 32, 32, device=DEV)
    a.train(); b.train()
    with torch.no_grad():
        ya = a(x)
    with quiet():
        yb = b.forward_test(x)
    assert rel_err(yb, ya) < TOL
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert rel_err(sb[k], sa[k]) < TOL, k
    assert int(sb['features.LastTransUp.norm1.num_batches_tracked']) == 1


def _ref_or_skip():
    if ref_shims.ref_root() is None:
        pytest.skip('reference modules not staged (baseline/_ref is written by __graft_entry__.build() in the build container)')


def _compare_with_reference_module(ours, ref, x, seed, check_grads=True, tol=TOL):
    """same weights, same input, same torch RNG seed (dropout masks are drawn by the same torch calls in the same order)"""
    ref.load_state_dict(ours.state_dict())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    outs = []
    for m in (ours, ref):
        torch.manual_seed(seed)
        xi = x.clone().requires_grad_(True)
        y = m(xi)
        (y ** 2).sum().backward()
        outs.append((y.detach(), xi.grad, {n: p.grad.clone() for n, p in m.named_parameters()}, {k: v.clone() for k, v in m.state_dict().items()}))
        m.zero_grad()
    (ya, ga, pa, sa), (yb, gb, pb, sb) = outs
    assert rel_err(ya, yb) < tol
    if check_grads:
        # Both sides are fp32 autograd chains with different summation orders, so a pre-activation within ~1e-6 of zero can flip its
        # ReLU mask on one side: one pixel of one gradient map then moves by its full value (measured on the default net: 3 seeds of 4
        # agree to 6e-6 in every tensor, the 4th shows 2e-3 in the tensors upstream of one flipped pixel).  A flip is local, an indexing
        # or mask-stream error is not.  The partner here is cuDNN in fp32, whose weight gradients are atomically accumulated: two runs of
        # the REFERENCE differ by 4e-4 (relative L2) in In_conv1.weight of this net (a 5e-4-sized gradient left over from cancelling
        # terms), while this library accumulates weight gradients in fp64 and reproduces bit for bit.  Bounds: relative L2 2e-3, max-norm
        # 2e-2; the tight check of the same topology is test_m4_encoder_decoder_backward_vs_oracle (fp64 autograd, 2e-5).
        l2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        assert l2(ga, gb) < 2e-3 and rel_err(ga, gb) < 2e-2
        for n in pa:
            assert l2(pa[n], pb[n]) < 2e-3 and rel_err(pa[n], pb[n]) < 2e-2, n
    for k in sa:
        assert rel_err(sa[k], sb[k]) < tol, k


@pytest.mark.parametrize('training', [True, False])
def test_m4_dropout_model_vs_reference_modules(training):
    """drop_rate > 0 (Dropout2d after every dense-layer / transition conv, denseEDcirc2d.py:61-62,102-127): train mode runs module by
    module -- convolutions on libarpde with their own autograd nodes, dropout / BatchNorm by the torch ops the reference uses, hence the
    same masks for the same seed; eval mode runs the fused program."""
    _ref_or_skip()
    kw = dict(blocks=[2, 3, 2], growth_rate=4, init_features=16, drop_rate=0.25)
    torch.manual_seed(6)
    with quiet():
        ours = D.DenseED2d(6, 2, **kw).to(DEV)
        with ref_shims.reference_modules('burgers2d'):
            from nn.denseEDcirc2d import DenseED as RefDenseED
            ref = RefDenseED(6, 2, **kw).to(DEV)
    _randomise_bn(ours)
    ours.train(training); ref.train(training)
    _compare_with_reference_module(ours, ref, torch.randn(3, 6, 32, 32, device=DEV), seed=123)


def test_m4_bottleneck_dense_layer_vs_reference_modules():
    """bottleneck=True: the reference's dense layer ends in BatchNorm(bn_size*growth)+ReLU (its 'conv1' is registered twice,
    denseEDcirc2d.py:50-56); works only for bn_size == 1, which is what is compared; otherwise both raise."""
    _ref_or_skip()
    kw = dict(blocks=[3], growth_rate=2, init_features=16, bn_size=1, bottleneck=True)
    torch.manual_seed(7)
    with quiet():
        ours = D.DenseED2d(6, 2, **kw).to(DEV)
        with ref_shims.reference_modules('burgers2d'):
            from nn.denseEDcirc2d import DenseED as RefDenseED
            ref = RefDenseED(6, 2, **kw).to(DEV)
            bad_ref = RefDenseED(6, 2, blocks=[3], growth_rate=2, init_features=16, bn_size=2, bottleneck=True).to(DEV)
        bad = D.DenseED2d(6, 2, blocks=[3], growth_rate=2, init_features=16, bn_size=2, bottleneck=True).to(DEV)
    _randomise_bn(ours)
    ours.train(); ref.train()
    x = torch.randn(2, 6, 16, 16, device=DEV)
    _compare_with_reference_module(ours, ref, x, seed=5)
    with pytest.raises(RuntimeError):
        bad_ref(x)
    with pytest.raises(RuntimeError):
        bad(x)


def test_m4_bilinear_transition_vs_reference_modules():
    """_Transition(upsample='linear') (denseEDcirc2d.py:135-139): bilinear align_corners=True upsampling between the 1x1 and the 3x3 conv."""
    _ref_or_skip()
    torch.manual_seed(8)
    ours = D._Transition(2, 12, 6, down=False, upsample='linear').to(DEV)
    with ref_shims.reference_modules('burgers2d'):
        from nn.denseEDcirc2d import _Transition as RefTransition
        ref = RefTransition(12, 6, down=False, upsample='linear').to(DEV)
    _randomise_bn(ours)
    ours.train(); ref.train()
    _compare_with_reference_module(ours, ref, torch.randn(2, 12, 16, 16, device=DEV), seed=1)


# ----------------------------------------------------------------------------- C5 at the stated 256 x 256
def _c5_model():
    torch.manual_seed(5)
    with quiet():
        m = M.CLS['burgers2d'](6, 2, [4], growth_rate=32, init_features=96).to(DEV)
    _randomise_bn(m)
    return m


@pytest.mark.parametrize('training', [False, True])
def test_c5_forward_256_vs_oracle(training):
    """BASELINE.json configs[4]: wide DenseED (growth 32, init_features 96) on the 256x256 grid (one sample keeps the fp32 CPU oracle to
    ~12 GFLOP)."""
    m = _c5_model()
    x = torch.randn(1, 6, 256, 256, device=DEV)
    sd = {'features.' + k: v.detach().cpu().clone() for k, v in m.features.state_dict().items()}
    m.train(training)
    with torch.no_grad():
        y = m(x)
    ref = O.densed_forward('burgers2d', sd, x.cpu(), [4], training=training)
    assert y.shape == ref.shape == (1, 2, 256, 256)
    assert rel_err(y, ref) < TOL
    if training:
        for k, v in m.features.state_dict().items():
            assert rel_err(v, sd['features.' + k]) < TOL, k


C5_LAYERS = [
    # name, cin, cout, k, s, up, H, W   (input resolution of the layer at a 256x256 network input)
    ('In_conv1', 6, 96, 5, 1, 0, 256, 256),          # 5x5 at 256^2: the tensor-core weight-gradient ring exceeds shared memory -> generic kernel
    ('In_conv3', 96, 48, 3, 2, 0, 256, 256),         # stride 2 at 256^2: same fallback
    ('denselayer4', 144, 32, 3, 1, 0, 128, 128),
    ('LastTransUp.conv1', 176, 88, 1, 1, 0, 128, 128),
    ('LastTransUp.conv2', 88, 44, 3, 1, 1, 128, 128),
    ('LastTransUp.conv3', 44, 2, 5, 1, 0, 256, 256),
]


@pytest.mark.parametrize('name,cin,cout,k,s,up,H,W', C5_LAYERS)
def test_c5_layer_gradients_256_vs_oracle(name, cin, cout, k, s, up, H, W):
    """Weight gradient (auto-dispatched kernel AND the generic kernel, path = 1) and data gradient of every layer shape of the wide net
    at its 256x256 resolution, BN + ReLU prologue included, against fp64 autograd of the oracle's circular convolution.  Per layer, so
    that no ReLU-kink flip upstream can mask an indexing error."""
    torch.manual_seed(cin + cout)
    N = 2
    x = torch.randn(N, cin, H, W)
    Ho, Wo = (H << up) // s, (W << up) // s
    gy = torch.randn(N, cout, Ho, Wo)
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    # fp64 oracle: a = up2?(relu(sc*x+sh)); y = conv_circ(a, w)
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    a = torch.relu(x64 * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    if up:
        a = O._up2(a)
    O.conv_circ(a, w64, stride=s).backward(gy.double())
    xd, gd, scd, shd = x.to(DEV), gy.to(DEV), sc.to(DEV), sh.to(DEV)
    d64 = torch.empty((cout * cin * k * k,), device=DEV, dtype=torch.float64)
    n_scr = int(L.lib.arpde_conv_wgrad_tc_scratch(N, cin, H, W, cout, k, k, s, up))
    scr = torch.empty((max(n_scr, 1),), device=DEV)
    for path in (0, 1):
        dW = torch.full((cout, cin, k, k), 5.0, device=DEV)
        L.call('arpde_conv_wgrad', L.ptr(xd), cin * H * W, L.ptr(scd), L.ptr(shd), L.ptr(gd), cout * Ho * Wo, L.ptr(d64), L.ptr(dW),
               N, cin, H, W, cout, k, k, s, up, 1, L.ptr(scr) if n_scr else 0, n_scr, path, L.stream())
        torch.cuda.synchronize()
        assert rel_err(dW, w64.grad) < TOL, (name, path)
    # data gradient through the product's autograd path: a one-layer model [BN, ReLU, (up), conv] in eval mode reproduces sc/sh
    conv = D.CircConv2d(cin, cout, k, s).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(w)
    layers = [('norm', D.BatchNorm2d(cin)), ('relu', D.ReLU(True))] + ([('up', D.UpsamplingNearest2d(2))] if up else []) + [('conv', conv)]
    seq = D._Fused()
    for n_, l_ in layers:
        seq.add_module(n_, l_)
    seq = seq.to(DEV).eval()
    with torch.no_grad():
        seq.norm.weight.copy_(sc); seq.norm.bias.copy_(sh); seq.norm.running_mean.zero_(); seq.norm.running_var.fill_(1.0 - 1e-5)
    xg = xd.clone().requires_grad_(True)
    y = seq(xg)
    y.backward(gd)
    assert rel_err(xg.grad, x64.grad) < TOL, name
    assert rel_err(conv.weight.grad, w64.grad) < TOL, name
