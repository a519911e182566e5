"""CPU check of the tensor-core weight-gradient tiling: the planner (arpde_conv_wgrad_tc_plan, host code of
conv_wgrad_tc.cu) is asked for the real plan of a layer and the kernel's index arithmetic -- patch ring with mirror rows,
parity planes, (kx, ci) column packing, chunked positions of the padded row pitch, tap offsets -- is replayed in numpy
with the converter running as far ahead of the MMAs as the barriers allow (and as little as they require), then compared
with the autograd weight gradient of the oracle's circular convolution.  No GPU involved."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import arpde_oracle as O
from arpde_b200 import _lib as L

NAMES = 'pitch min_ry min_rx RB RBH nch_u lead pcu NR npl NC ncc cchunk packed ntaps M mrows tmem slices smem Ho Wo sy sx swap nblk'.split()


def plan(cin, H, W, cout, kh, kw, s, up):
    out = (C.c_int32 * 400)()
    L.call('arpde_conv_wgrad_tc_plan', cin, H, W, cout, kh, kw, s, up, C.addressof(out))
    p = dict(zip(NAMES, out[:len(NAMES)]))
    p['fold'], p['cout_fold'] = out[26], out[27]
    p['tap_plane'] = list(out[32:32 + p['ntaps']])
    p['tap_off'] = list(out[81:81 + p['ntaps']])
    p['col_ci'] = list(out[130:258])
    p['col_dy'] = list(out[258:386])
    return p


def replay(x, gy, up, kh, kw, p, ahead):
    """x [N,cin,H,W], gy [N,cout,Ho,Wo] -> dW [cout,cin,kh,kw] following conv_wgrad_tc_kernel's indexing.
    ahead = True: the converter fills every ring chunk the empty barriers allow before each MMA chunk."""
    N, cin, H, W = x.shape
    cout = gy.shape[1]
    Ho, Wo, pitch, NR, lead, nch_u, pcu = p['Ho'], p['Wo'], p['pitch'], p['NR'], p['lead'], p['nch_u'], p['pcu']
    ring = NR * 32
    dW = np.zeros((cout, cin, kh, kw))
    for cc in range(p['ncc']):
        nblk, swap = p['nblk'], p['swap']
        climit = min(cin, (cc + 1) * p['cchunk'])
        D = np.zeros((p['M'] if swap else cout, p['ntaps'], p['NC']))
        patch = np.full((p['npl'] * nblk, ring + 8, 32), np.nan)       # slot = plane * nblk + 32-channel block
        units = [(n, rb) for n in range(N) for rb in range(p['RB'])]

        def produce(g):      # global patch chunk g -> ring slot
            lu, t = divmod(g, pcu)
            n, rb = units[lu]
            R0 = rb * p['RBH']
            s = g % NR
            for lane in range(32):
                pos = t * 32 + lane
                i, j = divmod(pos, pitch)
                for e in range(32 if swap else p['NC']):
                  for blk in range(nblk):
                    ci = p['col_ci'][32 * blk + e]
                    c = cc * p['cchunk'] + ci if ci >= 0 else -1
                    if c < 0 or c >= climit:
                        val = np.zeros(p['npl'])
                    else:
                        R = (R0 + p['min_ry'] + i + p['col_dy'][32 * blk + e]) % Ho
                        Cc = (p['min_rx'] + j) % Wo
                        sr, sc = (R * p['sy']) >> up, (Cc * p['sx']) >> up
                        val = np.array([x[n, c, sr + (pl >> 1), sc + (pl & 1)] for pl in range(p['npl'])])
                    rr = s * 32 + lane
                    sl = slice(blk, None, nblk) if nblk == 1 else slice(blk, blk + 1)
                    patch[sl, rr, e] = val
                    if s == 0 and lane < 8:
                        patch[sl, ring + lane, e] = val
        produced = 0
        total_chunks = len(units) * pcu
        released = 0          # chunks [0, released) have been released by the MMAs
        pg = 0
        for lu, (n, rb) in enumerate(units):
            R0 = rb * p['RBH']
            for j in range(nch_u):
                need = pg + j + lead
                limit = min(total_chunks, released + NR) if ahead else need + 1
                assert need < released + NR, 'deadlock: ring too small'
                while produced < limit:
                    produce(produced)
                    produced += 1
                base_row = ((pg + j) % NR) * 32
                for k8 in range(4):
                    q = j * 32 + k8 * 8 + np.arange(8)
                    oy, ox = q // pitch, q % pitch
                    g8 = np.where(ox < Wo, gy[n][:, np.minimum(R0 + oy, Ho - 1), np.minimum(ox, Wo - 1)], 0.0)     # [cout, 8]
                    for t in range(p['ntaps']):
                        rr = base_row + k8 * 8 + p['tap_off'][t]
                        if rr >= ring:
                            rr -= ring
                        assert rr + 8 <= ring + 8
                        if swap:      # ring = M operand: rows = channels of the nblk blocks, columns = cout (padded)
                            xa = np.concatenate([patch[p['tap_plane'][t] * nblk + blk, rr:rr + 8, :] for blk in range(nblk)], axis=1)   # [8, M]
                            assert not np.isnan(xa[:, [r for r in range(p['M']) if p['col_ci'][r] >= 0 and cc * p['cchunk'] + p['col_ci'][r] < climit]]).any()
                            D[:, t, :cout] += np.nan_to_num(xa).T @ g8.T
                            continue
                        b = patch[p['tap_plane'][t], rr:rr + 8, :p['NC']]                               # [8, NC]
                        assert not np.isnan(b).any(), (lu, j, k8, t)
                        D[:, t, :] += g8 @ b
                released = pg + j + 1
                if j == nch_u - 1:
                    released = pg + pcu
            pg += pcu
        if swap:
            for row in range(p['M']):
                ci = p['col_ci'][row]
                c = cc * p['cchunk'] + ci
                if ci < 0 or c >= climit:
                    continue
                for t in range(p['ntaps']):
                    tap = p['col_dy'][row] * kw + t if p['packed'] else t
                    dW[:, c, tap // kw, tap % kw] += D[row, t, :cout]
            continue
        for t in range(p['ntaps']):
            for e in range(p['NC']):
                ci = p['col_ci'][e]
                c = cc * p['cchunk'] + ci if ci >= 0 else -1
                if 0 <= c < cin:
                    tap = p['col_dy'][e] * kw + t if p['packed'] else t
                    dW[:, c, tap // kw, tap % kw] += D[:, t, e]
    return dW


CASES = [  # cin, cout, kh, kw, stride, up, H, W
    (6, 16, 5, 5, 1, 0, 16, 16), (40, 16, 3, 3, 2, 0, 16, 32), (12, 16, 3, 3, 1, 1, 8, 8), (3, 20, 7, 7, 1, 0, 16, 16),
    (12, 16, 4, 4, 2, 0, 32, 32), (33, 17, 3, 3, 1, 0, 8, 16), (48, 4, 3, 3, 1, 0, 8, 16), (150, 8, 3, 3, 1, 0, 8, 16), (16, 2, 5, 5, 1, 0, 16, 16), (64, 32, 1, 1, 1, 0, 8, 16)]


@pytest.mark.parametrize('ahead', [False, True])
@pytest.mark.parametrize('cin,cout,kh,kw,s,up,H,W', CASES)
def test_wgrad_ring_index_arithmetic(cin, cout, kh, kw, s, up, H, W, ahead):
    p = plan(cin, H, W, cout, kh, kw, s, up)
    rng = np.random.default_rng(cin * 100 + kh)
    N = 2
    x = rng.standard_normal((N, cin, H, W))
    gy = rng.standard_normal((N, cout, p['Ho'], p['Wo']))
    if p['fold']:
        # nearest x2 + 3x3 folded: gradient of the low-resolution conv with 4*cout parity-major channels, then every folded tap
        # (ty, tx) of parity (py, px) is added into the original taps with (py + ky + 1) >> 1 == ty and (px + kx + 1) >> 1 == tx
        assert up == 1 and kh == 3 and kw == 3 and p['cout_fold'] == cout
        gyf = np.concatenate([gy[:, :, py::2, px::2] for py in range(2) for px in range(2)], axis=1)       # [N, 4*cout, H, W]
        dWf = replay(x, gyf, 0, 3, 3, p, ahead)
        dW = np.zeros((cout, cin, 3, 3))
        for py in range(2):
            for px in range(2):
                for ky in range(3):
                    for kx in range(3):
                        dW[:, :, ky, kx] += dWf[(py * 2 + px) * cout:(py * 2 + px + 1) * cout, :, (py + ky + 1) >> 1, (px + kx + 1) >> 1]
    else:
        dW = replay(x, gy, up, kh, kw, p, ahead)
    xt = torch.from_numpy(x)
    if up:
        xt = O._up2(xt)
    w = torch.zeros((cout, cin, kh, kw), dtype=torch.float64, requires_grad=True)
    y = O.conv_circ(xt, w, stride=s)
    y.backward(torch.from_numpy(gy))
    ref = w.grad.numpy()
    assert np.abs(dW - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


def test_wgrad_plan_of_the_c4_layers():
    p1 = plan(6, 64, 64, 96, 5, 5, 1, 0)            # stem: (ky, ci) packed into the MMA's N, the 5 x-taps left (one-chunk halo)
    assert p1['packed'] == 1 and p1['NC'] == 32 and p1['ntaps'] == 5 and p1['M'] == 128 and p1['lead'] == 1 and p1['smem'] <= 227 * 1024
    p3 = plan(96, 64, 64, 48, 3, 3, 2, 0)           # stride 2: four parity planes, three channel chunks
    assert p3['npl'] == 4 and p3['ncc'] == 3 and p3['M'] == 64 and p3['ntaps'] * p3['NC'] <= 512 and p3['smem'] <= 227 * 1024
    assert p3['NR'] >= p3['lead'] + 2
    pd = plan(60, 32, 32, 4, 3, 3, 1, 0)            # dense layer: roles swapped, 64 accumulator rows = two channel blocks, N = 8
    assert pd['swap'] == 1 and pd['M'] == 64 and pd['nblk'] == 2 and pd['NC'] == 8 and pd['ncc'] == 1
    pl = plan(16, 64, 64, 2, 5, 5, 1, 0)            # LastTransUp 5x5 16 -> 2: swapped and packed, 80 (ky, ci) rows in three blocks
    assert pl['swap'] == 1 and pl['packed'] == 1 and pl['M'] == 128 and pl['nblk'] == 4 and pl['ntaps'] == 5 and pl['smem'] <= 227 * 1024


def test_wgrad_unsupported_layers_fall_back():
    out = (C.c_int32 * 400)()
    for args in [(16, 32, 32, 12, 3, 3, 1, 0), (4, 64, 64, 2, 3, 3, 1, 0), (96, 256, 256, 48, 3, 3, 2, 0), (16, 1, 512, 32, 1, 5, 1, 0)]:
        rc = L.lib.arpde_conv_wgrad_tc_plan(*args, C.addressof(out))
        assert rc != 0, args
