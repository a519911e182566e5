"""CPU check of the tensor-core convolution's tiling: the planner (arpde_conv_tc_plan, host code of conv_tc.cu) is asked
for the real plan of a layer and the kernel's "shifted window" index arithmetic -- patch planes, tap offsets, M-tile ->
pixel map -- is replayed in numpy and compared with the oracle's circular convolution.  No GPU involved."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import arpde_oracle as O
from arpde_b200 import _lib as L

NAMES = ('pitch min_ry min_rx ntile ctas PP PHP nplanes cchunk ncc tchunk ntc ntaps coutP tmem nslot_p smem Ho Wo Hc Wc sy sx '
         'wanted resident nacc kp').split()


def plan(cin, H, W, cout, kh, kw, s, up):
    out = (C.c_int32 * 160)()
    L.call('arpde_conv_tc_plan', cin, H, W, cout, kh, kw, s, up, C.addressof(out))
    p = dict(zip(NAMES, out[:len(NAMES)]))
    p['upfold'], p['cout_fold'] = out[27], out[28]
    p['tap_plane'] = list(out[32:32 + kh * kw])
    p['tap_off'] = list(out[81:81 + kh * kw])
    return p


def replay(x, w, up, p):
    """x [cin,H,W], w [cout,cin,kh,kw] -> y [cout,Ho,Wo] following conv_tc_kernel's indexing."""
    cin, H, W = x.shape
    cout, ntaps = w.shape[0], w.shape[2] * w.shape[3]
    up_w, up_h = up, (0 if (H == 1 and w.shape[2] == 1) else up)
    wt = w.reshape(cout, cin, ntaps)
    y = np.full((cout, p['Ho'], p['Wo']), np.nan)
    for b in range(p['ctas']):
        q0 = b * p['ntile'] * 128
        r0 = q0 // p['pitch']
        ql0 = q0 - r0 * p['pitch']
        patch = np.full((p['nplanes'], p['PP'], cin), np.nan)
        for pl in range(p['nplanes']):
            py, px = pl >> 1, pl & 1
            for pos in range(p['PHP']):
                i, j = divmod(pos, p['pitch'])
                R = (r0 + p['min_ry'] + i) % p['Ho']
                Cc = (p['min_rx'] + j) % p['Wo']
                off_r, off_c = (R * p['sy']) >> up_h, (Cc * p['sx']) >> up_w      # plane (0,0) source coordinates
                patch[pl, pos] = x[:, off_r + py, off_c + px]
        for t in range(p['ntile']):
            for r in range(128):
                q = q0 + t * 128 + r
                oy, ox = divmod(q, p['pitch'])
                if oy >= p['Ho'] or ox >= p['Wo']:
                    continue
                acc = np.zeros(cout)
                for tap in range(ntaps):
                    pos = ql0 + t * 128 + r + p['tap_off'][tap]
                    assert pos < p['PP']
                    a = patch[p['tap_plane'][tap], pos]
                    assert not np.isnan(a).any(), (q, tap, pos)
                    acc += wt[:, :, tap] @ a
                y[:, oy, ox] = acc
    return y


CASES = [  # cin, cout, kh, kw, stride, up, H, W
    (6, 8, 5, 5, 1, 0, 16, 16), (12, 8, 3, 3, 2, 0, 16, 12), (5, 4, 3, 3, 1, 1, 6, 10), (8, 16, 4, 4, 2, 0, 12, 20),
    (3, 4, 7, 7, 2, 0, 16, 16), (4, 4, 1, 1, 1, 0, 8, 8), (7, 3, 1, 5, 1, 0, 1, 96), (7, 3, 1, 7, 2, 0, 1, 96),
    (7, 3, 1, 3, 1, 1, 1, 48), (20, 4, 3, 3, 1, 0, 32, 32), (16, 8, 1, 1, 1, 0, 32, 32)]


@pytest.mark.parametrize('cin,cout,kh,kw,s,up,H,W', CASES)
def test_shifted_window_index_arithmetic(cin, cout, kh, kw, s, up, H, W):
    p = plan(cin, H, W, cout, kh, kw, s, up)
    rng = np.random.default_rng(cin * 100 + kh)
    x, w = rng.standard_normal((cin, H, W)), rng.standard_normal((cout, cin, kh, kw))
    if p['upfold']:
        # nearest x2 + 3x3 folded into a 3x3 conv on the low-resolution grid with 4*cout parity-major channels (conv_tc.cu header):
        # low-res tap t of parity p collects the original taps k with (p + k + 1) >> 1 == t; channel (py, px, co) -> pixel (2y+py, 2x+px)
        assert p['cout_fold'] == cout and up == 1 and kh == 3 and kw == 3
        wf = np.zeros((4 * cout, cin, 3, 3))
        for py in range(2):
            for px in range(2):
                for ky in range(3):
                    for kx in range(3):
                        wf[(py * 2 + px) * cout:(py * 2 + px + 1) * cout, :, (py + ky + 1) >> 1, (px + kx + 1) >> 1] += w[:, :, ky, kx]
        yl = replay(x, wf, 0, p)
        y = np.zeros((cout, 2 * H, 2 * W))
        for py in range(2):
            for px in range(2):
                y[:, py::2, px::2] = yl[(py * 2 + px) * cout:(py * 2 + px + 1) * cout]
    else:
        y = replay(x, w, up, p)
    xt = torch.from_numpy(x)[None]
    if up:
        xt = O._up2(xt) if H > 1 else xt.repeat_interleave(2, -1)
    xin = xt if H > 1 else xt.squeeze(2)
    wt = torch.from_numpy(w) if H > 1 else torch.from_numpy(w).squeeze(2)
    ref = O.conv_circ(xin, wt, stride=s)[0].numpy().reshape(y.shape)
    assert not np.isnan(y).any()
    assert np.abs(y - ref).max() < 1e-10


def test_plans_of_the_default_network_fit_the_sm():
    for (cin, cout, k, s, up, H) in ((6, 96, 5, 1, 0, 64), (96, 48, 3, 2, 0, 64), (48, 4, 3, 1, 0, 32), (60, 4, 3, 1, 0, 32), (64, 32, 1, 1, 0, 32),
                                     (32, 16, 3, 1, 1, 32), (176, 88, 1, 1, 0, 128), (88, 44, 3, 1, 1, 128), (6, 96, 5, 1, 0, 256)):
        p = plan(cin, H, H, cout, k, k, s, up)
        assert p['smem'] <= 227 * 1024 and p['tmem'] <= 512 and p['nacc'] * p['ntile'] * 2 * p['coutP'] <= p['tmem']
        assert p['upfold'] == (1 if (up and k == 3 and 4 * cout <= 128) else 0)
        assert p['PHP'] <= 2048 and p['kp'] == -(-p['PHP'] // 256)
        # (the diagnostic plans a single sample; at rollout batch sizes the 4-channel dense layers go to the row-tiled FFMA kernel instead)
        assert p['wanted'] == 1


def test_unsupported_layers_are_reported():
    out = (C.c_int32 * 160)()
    with pytest.raises(RuntimeError):
        L.call('arpde_conv_tc_plan', 4, 8, 8, 200, 3, 3, 1, 0, C.addressof(out))       # cout > 128
