"""CPU: the tiling of the bulk-copy ring stencil kernels (csrc/stencil.cu: bulk_plan) through its host-only hook -- the invariants the
kernels rely on (rows per item a multiple of the rows a consumer thread rolls over, at least two stages within the shared memory of an
SM, stage size within the mbarrier's transaction-count range) and the dispatch rule (which widths take the path)."""
import ctypes as C

import pytest

from arpde_b200 import _lib as L

SMEM_MAX = 227 * 1024


def plan(N, H, W, backward, mode=0):
    ty, ns, smem, items = C.c_int(0), C.c_int(0), C.c_int64(0), C.c_int64(0)
    ok = L.lib.arpde_stencil_bulk_plan(N, H, W, backward, mode, C.byref(ty), C.byref(ns), C.byref(smem), C.byref(items))
    return (ty.value, ns.value, smem.value, items.value) if ok else None


def eligible_width(W):
    w4 = W // 4
    return W % 4 == 0 and w4 >= 1 and ((w4 & (w4 - 1)) == 0 and w4 <= 32 or w4 % 32 == 0)


@pytest.mark.parametrize('backward', [0, 1])
@pytest.mark.parametrize('mode', [0, 1])
def test_bulk_plan_invariants(backward, mode):
    rb = 2 if backward else 4
    nf = (6 if mode == 0 else 4) if backward else 4
    seen = 0
    for N in (1, 2, 7, 128, 240, 1920, 5000):
        for H in (2, 3, 8, 10, 17, 18, 64, 70, 256):
            for W in (4, 8, 12, 16, 23, 24, 32, 64, 128, 256, 384, 1024, 4096):
                p = plan(N, H, W, backward, mode)
                if not eligible_width(W):
                    assert p is None, (N, H, W)
                    continue
                if p is None:                                   # eligible width but two stages of the smallest tile do not fit
                    assert 2 * nf * (rb + 2) * W * 4 + 256 > SMEM_MAX, (N, H, W)
                    continue
                ty, ns, smem, items = p
                seen += 1
                stage = nf * (ty + 2) * W * 4
                assert ty >= rb and ty % rb == 0, (N, H, W, p)
                assert ty <= (H + rb - 1) // rb * rb, (N, H, W, p)                  # never taller than the (rounded-up) field
                assert 2 <= ns <= 4 and smem == ns * stage + 16 * ns and smem + 256 <= SMEM_MAX, (N, H, W, p)
                assert stage < (1 << 20), 'mbarrier transaction count'
                assert items == N * ((H + ty - 1) // ty), (N, H, W, p)
    assert seen > 300


def test_bulk_plan_reference_shapes():
    # C4 rollout batch: whole-sample items, three stages; C5: 16-row strips; the training batch gets shorter strips so that every SM has work
    assert plan(1920, 64, 64, 0)[:2] == (64, 3)
    assert plan(256, 256, 256, 0)[:2] == (16, 3)
    assert plan(1920, 64, 64, 1)[:2] == (32, 4)
    assert plan(256, 256, 256, 1)[0] == 8
    ty_small = plan(128, 64, 64, 0)[0]
    assert ty_small < 64 and 128 * (64 // ty_small) >= 2 * 148
    assert plan(3, 17, 23, 0) is None and plan(2, 6, 12, 1) is None                # widths of the fallback kernels
    assert plan(0, 64, 64, 0) is None and plan(4, 64, 64, 0, mode=5) is None
