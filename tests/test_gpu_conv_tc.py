"""GPU parity of the tensor-core (tcgen05 / TMEM, 3xTF32) circular convolution against the fp64 oracle.

Same contract as arpde_conv_fwd (tests/test_gpu_parity.py::test_circular_conv_semantics); tolerance is the
north_star's 1e-5 relative (max|a-b|/max|b|): the TF32 split must deliver fp32-grade results."""
import pytest
import torch

from _util import rel_err
from oracle import arpde_oracle as O

from arpde_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def conv_tc(x4, x_bs, w, w_gs, sc, sh, ss_gs, y, stats, N, cin, H, W, cout, kh, kw, s, up, relu_in, act, y_ctot, y_coff, npg):
    groups = N // npg
    n = int(L.lib.arpde_conv_fwd_tc_scratch_floats(cin, cout, kh, kw, groups))
    scratch = torch.empty((max(n, 1),), device=DEV)
    L.call('arpde_conv_fwd_tc', L.ptr(x4), x_bs, L.ptr(w), w_gs, L.ptr(sc), L.ptr(sh), ss_gs, L.ptr(y), L.ptr(stats), N, cin, H, W,
           cout, kh, kw, s, up, relu_in, act, y_ctot, y_coff, npg, L.ptr(scratch), n, L.stream())
    torch.cuda.synchronize()


# the layers of the default / wide 2D networks and the generic-topology kernels
LAYERS_2D = [
    # cin, cout, k, s, H, W
    (6, 96, 5, 1, 64, 64),        # In_conv1
    (96, 48, 3, 2, 64, 64),       # In_conv3
    (48, 4, 3, 1, 32, 32),        # denselayer1
    (60, 4, 3, 1, 32, 32),        # denselayer4
    (64, 32, 1, 1, 32, 32),       # LastTransUp.conv1
    (16, 2, 5, 1, 64, 64),        # LastTransUp.conv3
    (144, 32, 3, 1, 16, 16),      # wide dense layer (C5 growth 32)
    (17, 5, 3, 1, 10, 24),        # odd channel counts, non-square
    (8, 16, 4, 2, 12, 20),        # TransDown k4/s2 (pad 1)
    (12, 24, 7, 2, 16, 16),       # k7/s2
    (5, 128, 3, 1, 8, 8),         # cout = 128 (N' = 256)
]


@pytest.mark.parametrize('cin,cout,k,s,H,W', LAYERS_2D)
def test_conv_tc_2d_vs_oracle(cin, cout, k, s, H, W):
    torch.manual_seed(cin * 7 + cout)
    N = 3
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, k, k) / (cin * k * k) ** 0.5
    ref = O.conv_circ(x.double(), w.double(), stride=s)
    xd = x.to(DEV)
    y = torch.full((N, cout, H // s, W // s), 3.0, device=DEV)
    conv_tc(xd, cin * H * W, w.to(DEV), 0, None, None, 0, y, None, N, cin, H, W, cout, k, k, s, 0, 0, 0, cout, 0, N)
    assert rel_err(y, ref) < TOL


@pytest.mark.parametrize('k,s,L_', [(3, 1, 512), (5, 1, 96), (7, 2, 96), (3, 2, 512), (1, 1, 256)])
def test_conv_tc_1d_vs_oracle(k, s, L_):
    torch.manual_seed(k + s + L_)
    N, cin, cout = 4, 34, 17
    x = torch.randn(N, cin, L_)
    w = torch.randn(cout, cin, k) / (cin * k) ** 0.5
    ref = O.conv_circ(x.double(), w.double(), stride=s)
    y = torch.empty((N, cout, 1, L_ // s), device=DEV)
    conv_tc(x.unsqueeze(2).to(DEV), cin * L_, w.to(DEV), 0, None, None, 0, y, None, N, cin, 1, L_, cout, 1, k, s, 0, 0, 0, cout, 0, N)
    assert rel_err(y.squeeze(2), ref) < TOL


@pytest.mark.parametrize('cin,cout,H,W', [(32, 16, 32, 32), (20, 8, 16, 24), (24, 32, 16, 16), (36, 5, 12, 16)])
def test_conv_tc_upsample_fold(cin, cout, H, W):
    """nearest x2 + 3x3 folded into a low-resolution conv with 4*cout parity channels: float2 epilogue (16 | cout), generic folded
    stores (cout = 8, 5), N' = 256 (cout = 32); BN+ReLU prologue, output written at a channel offset."""
    torch.manual_seed(cin + cout)
    N = 3
    x = torch.randn(N, cin, H, W)
    w = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    a = torch.relu(x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    ref = O.conv_circ(O._up2(a), w.double(), stride=1)
    xd, wd, scd, shd = x.to(DEV), w.to(DEV), sc.to(DEV), sh.to(DEV)
    ybuf = torch.full((N, cout + 4, 2 * H, 2 * W), 7.0, device=DEV)
    conv_tc(xd, cin * H * W, wd, 0, scd, shd, 0, ybuf, None, N, cin, H, W, cout, 3, 3, 1, 1, 1, 0, cout + 4, 4, N)
    assert float((ybuf[:, :4] - 7.0).abs().max()) == 0.0
    assert rel_err(ybuf[:, 4:], ref) < TOL


def test_conv_tc_prologue_upsample_concat_groups_stats():
    """BN+ReLU prologue, nearest x2, channel-offset output, per-group weights, tanh, batch statistics."""
    torch.manual_seed(0)
    N, cin, cout, H, W = 6, 32, 16, 16, 16
    x = torch.randn(N, cin + 3, H, W, device=DEV)[:, 3:]          # channel-slice view
    G = 3
    w = torch.randn(G, cout, cin, 3, 3, device=DEV) * 0.1
    sc = torch.rand(G, cin, device=DEV) + 0.5
    sh = torch.randn(G, cin, device=DEV) * 0.3
    ybuf = torch.full((N, cout + 5, 2 * H, 2 * W), 7.0, device=DEV)
    stats = torch.zeros(2 * (cout + 5), device=DEV, dtype=torch.float64)
    conv_tc(x, (cin + 3) * H * W, w, cout * cin * 9, sc, sh, cin, ybuf, stats, N, cin, H, W, cout, 3, 3, 1, 1, 1, 1, cout + 5, 5, N // G)
    xc = x.cpu().double()
    ref = []
    for n in range(N):
        gi = n // (N // G)
        a = torch.relu(xc[n:n + 1] * sc[gi].cpu().double().view(1, -1, 1, 1) + sh[gi].cpu().double().view(1, -1, 1, 1))
        ref.append(torch.tanh(O.conv_circ(O._up2(a), w[gi].cpu().double())))
    ref = torch.cat(ref, 0)
    assert rel_err(ybuf[:, 5:], ref) < TOL
    assert float((ybuf[:, :5] - 7.0).abs().max()) == 0.0           # channels outside the window untouched
    st = stats.view(-1, 2)[5:].cpu()
    assert rel_err(st[:, 0], ref.sum((0, 2, 3))) < 1e-5 and rel_err(st[:, 1], (ref ** 2).sum((0, 2, 3))) < 1e-5


def test_conv_tc_matches_ffma_kernel_on_large_batch():
    """Many CTAs per SM / several waves: the tensor-core kernel against the fp32 FFMA kernel of the same library."""
    torch.manual_seed(1)
    N, cin, cout, H, W = 40, 6, 96, 64, 64
    x = torch.randn(N, cin, H, W, device=DEV)
    w = torch.randn(cout, cin, 5, 5, device=DEV) / (cin * 25) ** 0.5
    y0 = torch.empty((N, cout, H, W), device=DEV)
    y1 = torch.empty_like(y0)
    L.call('arpde_conv_fwd', L.ptr(x), cin * H * W, L.ptr(w), 0, 0, 0, 0, L.ptr(y0), 0, N, cin, H, W, cout, 5, 5, 1, 0, 0, 0, cout, 0, N, L.stream())
    conv_tc(x, cin * H * W, w, 0, None, None, 0, y1, None, N, cin, H, W, cout, 5, 5, 1, 0, 0, 0, cout, 0, N)
    assert rel_err(y1, y0) < TOL


def test_conv_tc_errors():
    x = torch.zeros(1, 4, 8, 8, device=DEV)
    w = torch.zeros(200, 4, 3, 3, device=DEV)
    y = torch.zeros(1, 200, 8, 8, device=DEV)
    with pytest.raises(RuntimeError):      # cout > 128 is outside the tensor-core kernel
        conv_tc(x, 256, w, 0, None, None, 0, y, None, 1, 4, 8, 8, 200, 3, 3, 1, 0, 0, 0, 200, 0, 1)
    with pytest.raises(RuntimeError):      # scratch too small
        L.call('arpde_conv_fwd_tc', L.ptr(x), 256, L.ptr(w), 0, 0, 0, 0, L.ptr(y), 0, 1, 4, 8, 8, 16, 3, 3, 1, 0, 0, 0, 16, 0, 1,
               L.ptr(y), 4, L.stream())
