"""GPU ground-truth FD solver (csrc/fd_solver.cu, arpde_b200.solver.fd_burgers1d) against the fixture recorded from the reference's
burgers1d() (1D-Burger-SWAG/solver/fd_burger1D.py:55-98, tests/golden/fd_burgers1d.npz) and against the numpy oracle on a larger grid."""
import numpy as np
import pytest
import torch

from _util import golden, rel_err
from oracle import arpde_oracle as O

from arpde_b200 import solver

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('go', [(1, 2), (2, 4), (1, 4)])
def test_fd_solver_vs_reference_fixture(go):
    g = golden('fd_burgers1d')
    n = int(g['ncell'])
    u = solver.fd_burgers1d(torch.from_numpy(g['u0']).to(DEV), float(g['dt']), int(g['nsteps']), nu=float(g['nu']), dx=1.0 / n, grad_order=go)
    assert u.shape == (3, int(g['nsteps']), n) and u.dtype == torch.float32
    # both integrate in fp64 and store fp32: agreement to fp32 rounding (bit-exactness is not defined: cuDNN/MKL vs our tap order)
    assert rel_err(u[:, g['rows']], g['u/g%dd%d' % go]) < 2e-7


def test_fd_solver_reference_grid_vs_oracle_and_save_every():
    """the reference's own configuration (512 cells, dt 1e-4, nu 0.0025) for 2000 steps, every 100th step saved; a 1D input"""
    n, dt, nsteps, nu = 512, 1e-4, 2000, 0.0025
    x = np.linspace(0, 1, n + 1)[:-1]
    rng = np.random.RandomState(0)
    u0 = np.stack([np.sin(2 * np.pi * x) + 0.5 * np.cos(6 * np.pi * x + 1.0), 2 * rng.rand() * np.sin(4 * np.pi * (x + 0.3)) - 0.7])
    rows = np.arange(99, nsteps, 100)
    ref = O.fd_burgers1d(u0, dt, nsteps, 1.0 / n, nu, rows=rows)
    u = solver.fd_burgers1d(torch.from_numpy(u0).to(DEV), dt, nsteps, nu=nu, save_every=100)
    assert u.shape == (2, 20, n)
    assert rel_err(u, ref) < 2e-7
    u1 = solver.fd_burgers1d(torch.from_numpy(u0[0]).float().to(DEV), dt, 300, nu=nu, save_every=100)
    assert u1.shape == (3, n) and rel_err(u1, ref[0, :3]) < 1e-6          # fp32 initial state
    # long line: several points per thread
    n2 = 2048
    x2 = np.linspace(0, 1, n2 + 1)[:-1]
    v0 = np.sin(2 * np.pi * x2)[None]
    ref2 = O.fd_burgers1d(v0, 2e-5, 50, 1.0 / n2, nu, grad_order=(2, 4), rows=[49])
    u2 = solver.fd_burgers1d(torch.from_numpy(v0).to(DEV), 2e-5, 50, nu=nu, grad_order=(2, 4), save_every=50)
    assert rel_err(u2, ref2) < 2e-7
    with pytest.raises(RuntimeError):
        solver.fd_burgers1d(torch.from_numpy(u0), dt, 10)              # CPU tensor: no CPU path
