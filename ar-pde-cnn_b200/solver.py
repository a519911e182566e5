"""Ground-truth finite-difference solver on the GPU.

`fd_burgers1d` is the batched, single-launch form of `burgers1d()` in the reference's data-generation script
1D-Burger-SWAG/solver/fd_burger1D.py:32-98 (RK4 + upwind gradient + central second derivative, fp64, periodic): it gives in-repo
reference trajectories for the 1D Burgers surrogate without the external FEniCS data sets."""
import torch

from . import _lib as L


@torch.no_grad()
def fd_burgers1d(u0, dt, nsteps, nu=0.0025, dx=None, grad_order=(1, 2), save_every=1):
    """u0: [N, L] (or [L]) initial states on the periodic grid x_i = i*dx (CUDA tensor, any float dtype; integrated in float64).
    Returns float32 [N, nsteps // save_every, L]: the states after save_every, 2*save_every, ... steps (the reference stores float32)."""
    L.require_cuda(u0)
    one = u0.dim() == 1
    u = (u0[None] if one else u0).to(torch.float64).contiguous()
    if u.dim() != 2:
        raise ValueError('u0 must be [N, L] or [L], got %s' % (tuple(u0.shape),))
    N, Ln = u.shape
    dx = 1.0 / Ln if dx is None else float(dx)
    nsave = int(nsteps) // int(save_every)
    out = torch.empty((N, nsave, Ln), device=u.device, dtype=torch.float32)
    with torch.cuda.device(u.device):
        L.call('arpde_fd_burgers1d', L.ptr(u), L.ptr(out), N, Ln, dx, float(nu), float(dt), int(nsteps), int(save_every),
               int(grad_order[0]), int(grad_order[1]), L.stream())
    return out[0] if one else out
