"""GPU parity of the tensor-core weight gradient (conv_wgrad_tc.cu: tcgen05 with an MN-major patch ring, 3xTF32)
against the fp64 autograd gradient of the oracle's circular convolution.  Tolerance: the north_star's 1e-5 relative
(max|a-b|/max|b|)."""
import pytest
import torch

from _util import rel_err
from oracle import arpde_oracle as O

from arpde_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
TOL = 1e-5


def wgrad_tc(x, x_bs, sc, sh, dO, dO_bs, N, cin, H, W, cout, kh, kw, s, up, relu_in):
    d64 = torch.empty((cout * cin * kh * kw,), device=DEV, dtype=torch.float64)
    dW = torch.full((cout, cin, kh, kw), 5.0, device=DEV)
    n = int(L.lib.arpde_conv_wgrad_tc_scratch(N, cin, H, W, cout, kh, kw, s, up))
    scratch = torch.full((max(n, 1),), float('nan'), device=DEV)          # the kernel must not depend on its initial contents
    L.call('arpde_conv_wgrad_tc', L.ptr(x), x_bs, L.ptr(sc), L.ptr(sh), L.ptr(dO), dO_bs, L.ptr(d64), L.ptr(dW), N, cin, H, W, cout,
           kh, kw, s, up, relu_in, L.ptr(scratch), n, L.stream())
    torch.cuda.synchronize()
    return dW


def oracle_wgrad(x, sc, sh, relu_in, up, gy, cout, kh, kw, s):
    a = x.double()
    if sc is not None:
        a = a * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1)
    if relu_in:
        a = torch.relu(a)
    if up:
        a = O._up2(a)
    w = torch.zeros((cout, x.shape[1], kh, kw), dtype=torch.float64, requires_grad=True)
    O.conv_circ(a, w, stride=s).backward(gy.double())
    return w.grad


LAYERS = [
    # cin, cout, k, s, up, H, W, N
    (6, 96, 5, 1, 0, 64, 64, 5),         # In_conv1: (ky, ci) packed columns, M = 128
    (96, 48, 3, 2, 0, 64, 64, 5),        # In_conv3: stride 2 (four parity planes), three channel chunks, M = 64
    (32, 16, 3, 1, 1, 32, 32, 5),        # LastTransUp conv3x3 on the nearest-upsampled input
    (48, 32, 3, 1, 0, 32, 32, 4),        # wide dense layer
    (144, 32, 3, 1, 0, 16, 16, 6),       # C5 growth-32 dense layer: five channel chunks, the last one partial
    (33, 17, 3, 1, 0, 16, 24, 3),        # odd channel counts, non-square
    (12, 24, 4, 2, 0, 32, 32, 3),        # TransDown k4/s2 (pad 1)
    (3, 20, 7, 1, 0, 32, 32, 3),         # k7, packed (21 columns)
    (20, 128, 3, 1, 0, 16, 16, 7),       # cout = 128
    (40, 64, 5, 1, 0, 16, 16, 2),        # 25 taps x 16 columns, M = 64
    (48, 4, 3, 1, 0, 32, 32, 5),         # denselayer1: swapped roles (channels on the accumulator rows, N = 8)
    (60, 4, 3, 1, 0, 32, 32, 5),         # denselayer4
    (150, 8, 3, 1, 0, 16, 16, 4),        # swapped, M = 128 (four channel blocks), two channel chunks
    (64, 5, 5, 1, 0, 16, 16, 3),         # swapped, 25 taps
    (16, 2, 5, 1, 0, 64, 64, 5),         # LastTransUp conv5x5: swapped and (ky, ci) packed, 80 rows in three channel blocks
    (64, 32, 1, 1, 0, 32, 32, 5),        # 1x1 conv: a single tap, two channel chunks
    (8, 4, 5, 1, 0, 32, 32, 3),          # swapped + packed with M = 64 (40 rows)
]


@pytest.mark.parametrize('cin,cout,k,s,up,H,W,N', LAYERS)
def test_wgrad_tc_vs_oracle(cin, cout, k, s, up, H, W, N):
    torch.manual_seed(cin * 11 + cout)
    x = torch.randn(N, cin, H, W)
    Ho, Wo = (H << up) // s, (W << up) // s
    gy = torch.randn(N, cout, Ho, Wo)
    ref = oracle_wgrad(x, None, None, 0, up, gy, cout, k, k, s)
    dW = wgrad_tc(x.to(DEV), cin * H * W, None, None, gy.to(DEV), cout * Ho * Wo, N, cin, H, W, cout, k, k, s, up, 0)
    assert rel_err(dW, ref) < TOL


def test_wgrad_tc_prologue_and_strided_views():
    """BN scale/shift + ReLU prologue; x and dO are channel-slice views of larger buffers (dense-block layout)."""
    torch.manual_seed(1)
    N, cin, cout, H, W = 6, 52, 16, 32, 32
    xbuf = torch.randn(N, cin + 8, H, W)
    gbuf = torch.randn(N, cout + 12, H, W)
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    x, gy = xbuf[:, :cin], gbuf[:, 12:]
    ref = oracle_wgrad(x, sc, sh, 1, 0, gy, cout, 3, 3, 1)
    xd, gd = xbuf.to(DEV), gbuf.to(DEV)
    dW = wgrad_tc(xd[:, :cin], (cin + 8) * H * W, sc.to(DEV), sh.to(DEV), gd[:, 12:], (cout + 12) * H * W, N, cin, H, W, cout, 3, 3, 1, 0, 1)
    assert rel_err(dW, ref) < TOL


def test_wgrad_tc_training_batch():
    """Training-sized batch (every CTA walks several row blocks, ring wraps many times) against an fp64 gradient computed on the GPU."""
    torch.manual_seed(2)
    N, cin, cout, H, W = 128, 96, 48, 64, 64
    x = torch.randn(N, cin, H, W, device=DEV)
    gy = torch.randn(N, cout, H // 2, W // 2, device=DEV)
    sc, sh = torch.rand(cin, device=DEV) + 0.5, torch.randn(cin, device=DEV) * 0.3
    dW = wgrad_tc(x, cin * H * W, sc, sh, gy, cout * (H // 2) * (W // 2), N, cin, H, W, cout, 3, 3, 2, 0, 1)
    a = torch.relu(x.double() * sc.double().view(1, -1, 1, 1) + sh.double().view(1, -1, 1, 1))
    ap = torch.nn.functional.pad(a, (1, 1, 1, 1), mode='circular')
    ref = torch.nn.grad.conv2d_weight(ap, (cout, cin, 3, 3), gy.double(), stride=2)
    assert rel_err(dW, ref) < TOL


def test_wgrad_tc_unsupported_layer_reports_error():
    x = torch.randn(2, 16, 32, 32, device=DEV)
    gy = torch.randn(2, 12, 32, 32, device=DEV)
    with pytest.raises(RuntimeError, match='not supported'):
        wgrad_tc(x, 16 * 32 * 32, None, None, gy, 12 * 32 * 32, 2, 16, 32, 32, 12, 3, 3, 1, 0, 0)


UP_SMALL = [(9, 4, 16, 16, 3), (9, 8, 16, 16, 3), (10, 5, 16, 16, 3), (9, 16, 16, 16, 3), (12, 4, 16, 16, 3), (24, 12, 16, 16, 2), (9, 4, 32, 32, 2)]


@pytest.mark.parametrize('cin,cout,H,W,N', UP_SMALL)
def test_wgrad_upsampled_small_channel_counts(cin, cout, H, W, N):
    """nearest-x2 + conv3x3 layers of small encoder-decoder topologies (LastTransUp / TransUp of a (2,3,2) net: 9 -> 4, 10 -> 5 ...):
    arpde_conv_wgrad's automatic choice (folded tensor-core kernel where it plans, else FFMA) and the generic kernel vs fp64 autograd.
    cin * 3 <= 32 would select the (ky, ci)-packed ring, which the fold does not support: regression test for that dispatch."""
    torch.manual_seed(cin * 7 + cout)
    x = torch.randn(N, cin, H, W)
    gy = torch.randn(N, cout, 2 * H, 2 * W)
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    ref = oracle_wgrad(x, sc, sh, 1, 1, gy, cout, 3, 3, 1)
    xd, gd, scd, shd = x.to(DEV), gy.to(DEV), sc.to(DEV), sh.to(DEV)
    d64 = torch.empty((cout * cin * 9,), device=DEV, dtype=torch.float64)
    n_scr = int(L.lib.arpde_conv_wgrad_tc_scratch(N, cin, H, W, cout, 3, 3, 1, 1))
    scr = torch.empty((max(n_scr, 1),), device=DEV)
    for path in (0, 1):
        dW = torch.full((cout, cin, 3, 3), 5.0, device=DEV)
        L.call('arpde_conv_wgrad', L.ptr(xd), cin * H * W, L.ptr(scd), L.ptr(shd), L.ptr(gd), cout * 4 * H * W, L.ptr(d64), L.ptr(dW),
               N, cin, H, W, cout, 3, 3, 1, 1, 1, L.ptr(scr) if n_scr else 0, n_scr, path, L.stream())
        torch.cuda.synchronize()
        assert rel_err(dW, ref) < TOL, path


NARROW = [  # cin, cout, k, H, W, N
    (16, 2, 5, 64, 64, 5),       # LastTransUp conv5x5
    (60, 4, 3, 32, 32, 5),       # denselayer4
    (7, 3, 3, 20, 24, 3),        # odd channel counts, image height not a multiple of the tile height
    (12, 1, 5, 16, 40, 2),
]


@pytest.mark.parametrize('cin,cout,k,H,W,N', NARROW)
def test_wgrad_narrow_layers_vs_oracle(cin, cout, k, H, W, N):
    """arpde_conv_wgrad picks the sliding-window kernel for cout <= 4 (path 0); path 1 = generic FFMA kernel as cross-check."""
    torch.manual_seed(cin + 100 * cout)
    xbuf = torch.randn(N, cin + 4, H, W)
    x = xbuf[:, 4:]                                               # channel-slice view (dense-block layout)
    gy = torch.randn(N, cout, H, W)
    sc, sh = torch.rand(cin) + 0.5, torch.randn(cin) * 0.3
    ref = oracle_wgrad(x, sc, sh, 1, 0, gy, cout, k, k, 1)
    xd, gd, scd, shd = xbuf.to(DEV), gy.to(DEV), sc.to(DEV), sh.to(DEV)
    d64 = torch.empty((cout * cin * k * k,), device=DEV, dtype=torch.float64)
    for path in (0, 1):
        dW = torch.full((cout, cin, k, k), 5.0, device=DEV)
        L.call('arpde_conv_wgrad', L.ptr(xd[:, 4:]), (cin + 4) * H * W, L.ptr(scd), L.ptr(shd), L.ptr(gd), cout * H * W, L.ptr(d64), L.ptr(dW),
               N, cin, H, W, cout, k, k, 1, 0, 1, 0, 0, path, L.stream())
        torch.cuda.synchronize()
        assert rel_err(dW, ref) < TOL, path
