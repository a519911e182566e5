"""Eval-mode stride-1 decoder convolutions on the fp16 two-term tensor-core kernel (csrc/conv_h.cu): BatchNorm affine -> ReLU -> [nearest x2]
-> circular conv (1x1 / 3x3) against the fp64 oracle convolution (oracle/arpde_oracle.py, following 2D-Burgers-SWAG/nn/denseEDcirc2d.py:103-108,
128-141,170-183), through the C ABI (arpde_conv_fwd_h) and as dispatched by arpde_rollout."""
import pytest
import torch

from _util import rel_err
from arpde_b200 import _lib as L
from oracle import arpde_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _conv_h(x, w, sc, sh, k, up, relu, n_per_group, y_ctot=None, y_coff=0, cin=None):
    N, ctot_in, H, W = x.shape
    cin = cin or ctot_in
    G = N // n_per_group
    cout = w.shape[-4]
    y_ctot = y_ctot or cout
    f = 2 if up else 1
    y = torch.full((N, y_ctot, H * f, W * f), 7.0, device=x.device)
    ns = int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, G))
    scratch = torch.empty(ns, device=x.device)
    grouped = w.dim() == 5
    L.call('arpde_conv_fwd_h', L.ptr(x), x.stride(0), L.ptr(w), w.stride(0) if grouped else 0, L.ptr(sc), L.ptr(sh),
           (sc.stride(0) if grouped else 0) if sc is not None else 0, L.ptr(y), N, cin, H, W, cout, k, up, relu, y_ctot, y_coff, n_per_group,
           L.ptr(scratch), scratch.numel(), L.stream())
    torch.cuda.synchronize()
    return y


def _oracle(x, w, sc, sh, up, relu):
    a = x.double().cpu()
    if sc is not None:
        a = a * sc.double().cpu().view(1, -1, 1, 1) + sh.double().cpu().view(1, -1, 1, 1)
    if relu:
        a = torch.relu(a)
    if up:
        a = O._up2(a)
    return O.conv_circ(a, w.double().cpu(), stride=1)


# the two layers of the default 2D network first (conv1x1 64 -> 32 at 32x32; upsampling conv3x3 32 -> 16), then odd heights (ragged last work
# item), channel counts that need padding, wide maps, a plain 3x3, tiny maps
@pytest.mark.parametrize('cin,cout,k,up,H,W,N', [(64, 32, 1, 0, 32, 32, 5), (32, 16, 3, 1, 32, 32, 5), (16, 24, 1, 0, 7, 12, 3), (48, 8, 3, 1, 9, 16, 2),
                                                (32, 32, 3, 1, 16, 64, 2), (48, 32, 3, 0, 32, 32, 2), (128, 64, 1, 0, 16, 16, 2), (16, 128, 3, 0, 5, 8, 3),
                                                (128, 16, 1, 0, 64, 64, 1), (32, 20, 3, 1, 6, 6, 4)])
def test_conv_h_matches_fp64_oracle(cin, cout, k, up, H, W, N):
    g = torch.Generator().manual_seed(cin * 100 + cout + H)
    x = torch.randn(N, cin, H, W, generator=g).to(DEV)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    sc = (0.5 + torch.rand(cin, generator=g)).to(DEV)
    sh = (0.3 * torch.randn(cin, generator=g)).to(DEV)
    y = _conv_h(x, w, sc, sh, k, up, 1, N)
    ref = _oracle(x, w, sc, sh, up, 1)
    # tolerance: the north_star's 1e-5 relative (max-norm over the tensor)
    assert rel_err(y.cpu(), ref) < 1e-5


def test_conv_h_grouped_weights_channel_window_input_view_no_affine():
    # SWAG layout (one weight / scale / shift set per group of samples, several groups per CTA), output into a channel window of a wider
    # buffer, input = the first channels of a wider dense-block buffer (batch stride != cin * H * W), and the no-prologue form
    g = torch.Generator().manual_seed(11)
    G, npg, ctot, cin, cout, H, W = 4, 3, 80, 64, 32, 32, 32
    xw = torch.randn(G * npg, ctot, H, W, generator=g).to(DEV)
    w = (torch.randn(G, cout, cin, 1, 1, generator=g) / 8).to(DEV)
    sc = (0.5 + torch.rand(G, cin, generator=g)).to(DEV)
    sh = (0.3 * torch.randn(G, cin, generator=g)).to(DEV)
    y = _conv_h(xw, w, sc, sh, 1, 0, 1, npg, y_ctot=40, y_coff=4, cin=cin)
    assert float((y[:, :4] - 7.0).abs().max()) == 0.0 and float((y[:, 36:] - 7.0).abs().max()) == 0.0
    for j in range(G):
        ref = _oracle(xw[j * npg:(j + 1) * npg, :cin], w[j], sc[j], sh[j], 0, 1)
        assert rel_err(y[j * npg:(j + 1) * npg, 4:36].cpu(), ref) < 1e-5
    w3 = (torch.randn(G, 16, 32, 3, 3, generator=g) / 17).to(DEV)
    x3 = (1e-3 * torch.randn(G * npg, 32, 16, 16, generator=g)).to(DEV)          # tiny values: the scaled fp16 residual
    y3 = _conv_h(x3, w3, None, None, 3, 1, 0, npg)
    for j in range(G):
        assert rel_err(y3[j * npg:(j + 1) * npg].cpu(), _oracle(x3[j * npg:(j + 1) * npg], w3[j], None, None, 1, 0)) < 1e-5


def test_conv_h_many_items_per_cta_and_unsupported_shapes():
    # more work items than CTAs x 2 patch buffers x 4 accumulator slots: every ring wraps several times
    g = torch.Generator().manual_seed(3)
    N, cin, cout, H, W = 700, 32, 16, 32, 32
    x = torch.randn(N, cin, H, W, generator=g).to(DEV)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / 17).to(DEV)
    sc = (0.5 + torch.rand(cin, generator=g)).to(DEV); sh = (0.3 * torch.randn(cin, generator=g)).to(DEV)
    y = _conv_h(x, w, sc, sh, 3, 1, 1, N)
    idx = [0, 1, 349, 698, 699]
    assert rel_err(y[idx].cpu(), _oracle(x[idx], w, sc, sh, 1, 1)) < 1e-5
    y2 = _conv_h(x, w, sc, sh, 3, 1, 1, N)
    assert torch.equal(y, y2)                                                     # deterministic
    scratch = torch.empty(1 << 20, device=DEV)
    for cin_, cout_, k_, up_ in [(24, 16, 1, 0), (32, 48, 3, 1), (32, 16, 5, 0), (32, 16, 1, 1)]:
        xb = torch.zeros(1, cin_, 8, 8, device=DEV); wb = torch.zeros(cout_, cin_, k_, k_, device=DEV); yb = torch.zeros(1, cout_, 16, 16, device=DEV)
        rc = L.lib.arpde_conv_fwd_h(L.ptr(xb), xb.stride(0), L.ptr(wb), 0, None, None, 0, L.ptr(yb), 1, cin_, 8, 8, cout_, k_, up_, 0, cout_, 0, 1,
                                    L.ptr(scratch), scratch.numel(), L.stream())
        assert rc == 3, (cin_, cout_, k_, up_, rc)                                # ARPDE_ERR_UNSUPPORTED


def test_rollout_uses_conv_h_and_equals_layerwise_rollout():
    import contextlib
    import io
    import _models as M
    from arpde_b200 import engine
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        net = M.CLS['burgers2d'](6, 2, [4], growth_rate=4, init_features=96).to(DEV)
    with torch.no_grad():
        for mod in net.modules():
            if hasattr(mod, 'running_var'):
                mod.running_mean.uniform_(-0.2, 0.2); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.2, 0.2)
    net.eval()
    ics = O.ic_burgers2d(64, range(3)).to(DEV)
    L.lib.arpde_launch_count(1)
    a = engine.rollout(net, ics, 4)
    n_fused = int(L.lib.arpde_launch_count(1))
    b = engine.rollout(net, ics, 4, layerwise=True)
    errs = [rel_err(a[:, t], b[:, t]) for t in range(5)]
    print('fused vs layerwise per-step rel err', ['%.1e' % e for e in errs], 'launches', n_fused)
    assert errs[0] == 0.0 and 0.0 < errs[1] < 1e-5 and max(errs) < 1e-4
