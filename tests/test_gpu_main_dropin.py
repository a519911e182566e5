"""T13: the reference's UNMODIFIED drivers -- 2D-Burgers-SWAG/main.py, 1D-Burger-SWAG/main.py,
1D-KS-SWAG/main.py -- run end to end on this repo's drop-in nn/ package (ar-pde-cnn_b200/dropin/<project>/nn) on the GPU:
argument parsing, data loaders, train() with torch.optim.Adam + ExponentialLR / ReduceLROnPlateau, test() rollout and
SwagNN.saveModel, two epochs.  The printed loss / MSE / noise of every epoch and the first saved steps of the test rollout are
compared with tests/golden/t13_main.json, recorded from the same command on the reference's own modules (CPU,
tests/golden/make_t13.py).

The driver scripts themselves are not part of this repository: they are read from the staged copy baseline/_ref (written by
__graft_entry__.build() in the build container, shipped to the GPU box next to the built library)."""
import json
import os
import sys

import pytest
import torch

from _util import GOLDEN, ROOT

sys.path.insert(0, os.path.join(GOLDEN))
import make_t13 as T13          # noqa: E402  (command lines + output parser shared with the fixture generator)
from oracle import ref_shims    # noqa: E402

pytestmark = pytest.mark.gpu

# Two epochs of Adam on an untrained net amplify fp32 reassociation differences (1e-6 per forward) through ~20-60 optimizer steps of
# BPTT; how much is CALIBRATED, not guessed: the same command is also run on the reference's own modules ON THIS GPU (cuDNN, TF32 off),
# and the spread between the reference's two backends (CUDA here vs the CPU fixture) is the yardstick -- the drop-in must stay within
# SPREAD_FACTOR x that spread (or FLOOR, whichever is larger) of the reference-on-CUDA run.
# The checksum of the test() rollout is far more sensitive than the training losses: in eval mode the barely-trained network runs on
# BatchNorm running statistics estimated from a handful of mini-batches, and even the reference on 1 vs 8 CPU threads differs by up to
# 7 % there (1D Burgers) while its losses agree to 7e-4 -- hence the separate floor for it.  Layer- and model-level parity at 1e-5 is
# the business of test_gpu_parity.py; this test proves the unmodified drivers run end to end on the drop-in and stay on the
# reference's trajectory to within the reference's own reproducibility.
SPREAD_FACTOR = 5.0
FLOOR = {'loss': 2e-3, 'mse': 2e-3, 'noise': 2e-3, 'pred': 1e-1}


def _close(ours, ref_cuda, ref_cpu, key):
    spread = abs(ref_cuda - ref_cpu)
    return abs(ours - ref_cuda) <= max(SPREAD_FACTOR * spread, FLOOR[key[-1]] * abs(ref_cuda)), (key, ours, ref_cuda, ref_cpu)


@pytest.mark.parametrize('system', ['burgers2d', 'burgers1d', 'ks1d'])
def test_reference_main_runs_on_dropin(system):
    if ref_shims.ref_root() is None:
        pytest.skip('reference drivers not staged (baseline/_ref is written by __graft_entry__.build() in the build container)')
    gold = json.load(open(os.path.join(GOLDEN, 't13_main.json')))[system]
    assert gold['argv'] == T13.ARGS[system]
    got = T13.run(system, 'dropin')
    ref = T13.run(system, 'ref')            # the reference's own nn/ modules on this GPU
    print(system, 'dropin   :', got['epochs'], got['preds'][:2])
    print(system, 'ref cuda :', ref['epochs'], ref['preds'][:2])
    print(system, 'ref cpu  :', gold['epochs'], gold['preds'][:2])
    assert [e['epoch'] for e in got['epochs']] == [e['epoch'] for e in ref['epochs']] == [e['epoch'] for e in gold['epochs']] == [1, 2]
    for a, b, c in zip(got['epochs'], ref['epochs'], gold['epochs']):
        for k in ('loss', 'mse', 'noise'):
            ok, info = _close(a[k], b[k], c[k], (system, a['epoch'], k))
            assert ok, info
    assert len(got['preds']) == len(ref['preds']) == len(gold['preds']) > 0
    for a, b, c in zip(got['preds'], ref['preds'], gold['preds']):
        assert a['shape'] == b['shape'] == c['shape']
        ok, info = _close(a['absmean_first4'], b['absmean_first4'], c['absmean_first4'], (system, 'pred'))
        assert ok, info
