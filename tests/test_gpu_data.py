"""GPU parity of the on-device initial-condition synthesis (csrc/ic.cu, arpde_b200.data.InitPeriodicCond2d) against the oracle's
restatement of the reference data set and against the golden samples taken from the reference itself.  fp64 on both sides,
rounded to fp32 once: tolerance 1e-6 absolute on fields normalised to |u| <= 3."""
import os

import numpy as np
import pytest
import torch

from oracle import arpde_oracle as O
from arpde_b200.data import InitPeriodicCond2d

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.mark.parametrize('ncells,order', [(64, 4), (16, 4), (48, 2), (256, 4)])
def test_ic_batch_vs_oracle(ncells, order):
    idx = [0, 1, 5, 77, 4242]
    ds = InitPeriodicCond2d(ncells, 100000, order=order, device='cuda:0')
    got = ds.batch(idx).cpu()
    ref = O.ic_burgers2d(ncells, idx, order=order)
    assert got.shape == ref.shape and got.dtype == torch.float32
    assert float((got - ref).abs().max()) <= TOL
    assert torch.equal(ds[77].cpu(), got[3])          # __getitem__ == row of the batch


def test_ic_vs_reference_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ic2d.npz'))
    for ncells in (16, 64):
        ds = InitPeriodicCond2d(ncells, 100000, device='cuda:0')
        got = ds.batch([int(i) for i in g['ic/%d/index' % ncells]]).cpu().numpy()
        assert np.abs(got - g['ic/%d/u' % ncells]).max() <= TOL


def test_ic_needs_cuda_device():
    with pytest.raises(RuntimeError, match='no CPU path'):
        InitPeriodicCond2d(16, 10, device='cpu')


def test_error_metrics_vs_oracle():
    """arpde_error_metrics == testBayesianMSE's MSE / ESE curves (post/plotMSE.py:131-137), incl. test_every sub-sampling and strided rows."""
    from arpde_b200 import engine
    torch.manual_seed(5)
    N, T, nel, every = 3, 10, 32, 2
    buf = torch.randn(N, T + 1, 5, nel, nel)                       # prediction rows are channel-slice views of a larger buffer
    pred = buf[:, :, 1:3]
    target = pred[:, ::every][:, :5] + 0.05 * torch.randn(N, 5, 2, nel, nel)
    mse_ref, ese_ref = O.error_metrics(pred, target, nel, test_every=every)
    mse, ese = engine.error_metrics(buf.to('cuda:0')[:, :, 1:3], target.to('cuda:0'), test_every=every, points=nel * nel)
    assert mse.shape == (N, 5) and mse.dtype == torch.float64
    assert float((mse.cpu() - mse_ref).abs().max() / mse_ref.abs().max()) < 1e-6
    assert float((ese.cpu() - ese_ref.double()).abs().max() / ese_ref.abs().max()) < 1e-4      # the reference sums 2048 squares in fp32
