"""CPU: the index arithmetic of the bulk-copy ring stencil kernels (csrc/stencil.cu: bulk_rows, unit_map, lds_row6, the consumer loop of
burgers2d_fwd_bulk / burgers2d_bwd_bulk) replayed in numpy from the REAL plan (arpde_stencil_bulk_plan), for shapes the GPU tests do not
all cover: every output point is produced exactly once, every tile row holds the periodic field row the consumer assumes, the halo values a
lane takes from its neighbour lane (or, at a warp boundary of a wide row, from the tile) are the periodic neighbours of its four columns.
The arithmetic of the stencil itself is the GPU tests' business; here a 'field' is just its own linear index."""
import ctypes as C

import numpy as np
import pytest

from arpde_b200 import _lib as L

CONSUMER_WARPS = 8


def plan(N, H, W, backward):
    ty, ns, smem, items = C.c_int(0), C.c_int(0), C.c_int64(0), C.c_int64(0)
    ok = L.lib.arpde_stencil_bulk_plan(N, H, W, backward, 0, C.byref(ty), C.byref(ns), C.byref(smem), C.byref(items))
    return (ty.value, ns.value, items.value) if ok else None


def bulk_rows(field, ystart, nrows):
    """producer: rows ystart .. ystart + nrows - 1 (periodic) as runs of consecutive rows, split where the wrap falls"""
    H = field.shape[0]
    out, y, left = [], ystart % H, nrows
    while left > 0:
        nr = min(left, H - y)
        out.append(field[y:y + nr])          # ONE contiguous bulk copy
        left -= nr; y = 0
    return np.concatenate(out, 0)


@pytest.mark.parametrize('backward', [0, 1])
@pytest.mark.parametrize('N,H,W', [(3, 64, 64), (2, 256, 256), (1, 2, 4), (2, 3, 8), (5, 18, 64), (2, 70, 256), (1, 10, 384), (1, 8, 1024),
                                   (7, 16, 16), (2, 33, 128), (1, 5, 32), (400, 8, 8)])
def test_ring_index_arithmetic(N, H, W, backward):
    p = plan(N, H, W, backward)
    assert p is not None, 'shape expected on the bulk-copy path'
    TY, NS, items = p
    RB = 2 if backward else 4
    w4 = W // 4
    LX = min(w4, 32)
    edge = w4 > 32
    nstrips = (H + TY - 1) // TY
    assert items == N * nstrips
    nrb, nunits = TY // RB, (TY // RB) * w4
    field = np.arange(H * W, dtype=np.int64).reshape(H, W)           # a field whose value is its own index
    produced = np.zeros((H, W), dtype=np.int32)
    for strip in range(nstrips):
        y0 = strip * TY
        tile = bulk_rows(field, y0 - 1, TY + 2)                       # [TY + 2][W]: tile row r = field row (y0 - 1 + r) mod H
        for r in range(TY + 2):
            assert np.array_equal(tile[r], field[(y0 - 1 + r) % H])
        for i0 in range(0, nunits, 32 * CONSUMER_WARPS):              # one pass of the 8 consumer warps
            for warp in range(CONSUMER_WARPS):
                base = i0 + warp * 32
                if base >= nunits:
                    continue
                lanes = np.arange(32)
                i = base + lanes
                c = i % w4
                rb = i // w4
                valid = rb < nrb
                rb = np.where(valid, rb, nrb - 1)
                x0 = c * 4
                lxi, lbase = lanes & (LX - 1), lanes & ~(LX - 1)
                lsrc, rsrc = lbase | ((lxi - 1) & (LX - 1)), lbase | ((lxi + 1) & (LX - 1))
                for r in range(RB + 2):                               # every tile row the unit reads
                    trow = tile[rb * RB + r]                          # [32 lanes][W]
                    own = np.stack([trow[lanes, x0 + j] for j in range(4)], 1)
                    left, right = own[lsrc, 3].copy(), own[rsrc, 0].copy()          # the two shuffles
                    if edge:                                          # lanes 0 / 31 of a row wider than one warp: predicated tile loads
                        left[0] = trow[0, (x0[0] - 1) % W]
                        right[31] = trow[31, (x0[31] + 4) % W]
                    yy = (y0 - 1 + rb * RB + r) % H
                    sel = valid
                    assert np.array_equal(left[sel], field[yy[sel], (x0[sel] - 1) % W]), (strip, warp, r)
                    assert np.array_equal(right[sel], field[yy[sel], (x0[sel] + 4) % W]), (strip, warp, r)
                for r in range(RB):                                   # the rows the unit writes
                    y = y0 + rb * RB + r
                    ok = valid & (y < H)
                    for j in range(4):
                        np.add.at(produced, (y[ok], x0[ok] + j), 1)
    assert produced.min() == 1 and produced.max() == 1, 'every output point exactly once'
