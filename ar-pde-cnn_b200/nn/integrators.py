"""Physics time integrators on libarpde -- mirrors of
  2D-Burgers-SWAG/nn/burger2DFiniteDifference.py (LapLaceFilter2d :6-54, SobelFilter2d :57-129, Burger2DIntegrate :131-211)
  1D-Burger-SWAG/nn/burgerFiniteDifference.py    (filters :8-86,:168-222, BurgerIntegrate :246-302)
  1D-KS-SWAG/nn/ksFiniteDifference.py            (filters :6-119, KSIntegrate :121-173)
crankNicolson / backwardEuler are single fused kernels, differentiable w.r.t. BOTH inputs (the loss target is
not detached in the reference, main.py:71-73).  The filter objects (.grad1/.grad2/.grad4) stay callable.

Variants no reference driver calls -- the 5x5 2D stencils (burger2DFiniteDifference.py:31-40,74-82), the kernel_size=2
"conditional upwind" gradient (burgerFiniteDifference.py:224-244, which returns the forward difference whatever the sign
of u: reproduced) and the WENO flux-limited integrator (:89-166,:304-322) -- run on a reference-equivalent path built from
torch.roll on the same CUDA tensors (differentiable through torch autograd), as SURVEY.md 8a row P4 prescribes."""
import ctypes as C

import torch

from .. import _lib as L
from .. import functional as F_

_GRAD = {3: ([-1., 0., 1.], 2.), 5: ([1., -8., 0., 8., -1.], 12.), 7: ([-1., 9., -45., 0., 45., -9., 1.], 60.)}
_LAP = {3: ([1., -2., 1.], 1.), 5: ([-1., 16., -30., 16., -1.], 12.)}
_D4 = {5: ([1., -4., 6., -4., 1.], 1.), 7: ([-1., 12., -39., 56., -39., 12., -1.], 6.)}


def _roll_filter(u, taps, dims):
    """periodic cross-correlation out[i] = sum_k w[k] u[i + k - p] along `dims` (1 or 2 trailing dims) with torch.roll; taps is a
    nested list.  Zero taps are skipped; the summation runs in tap order."""
    out = None
    if len(dims) == 1:
        p = (len(taps) - 1) // 2
        for k, w in enumerate(taps):
            if w != 0.0:
                t = torch.roll(u, shifts=p - k, dims=dims[0]) * w
                out = t if out is None else out + t
    else:
        p = (len(taps) - 1) // 2
        for ky, row in enumerate(taps):
            for kx, w in enumerate(row):
                if w != 0.0:
                    t = torch.roll(u, shifts=(p - ky, p - kx), dims=dims) * w
                    out = t if out is None else out + t
    return out


class _Filter1D(object):
    table, order, what = None, 1, ''

    def __init__(self, dx, kernel_size=3, device='cpu'):
        if kernel_size not in self.table:
            raise NotImplementedError('%s: kernel_size %s not supported (choices %s)' % (self.what, kernel_size, sorted(self.table)))
        self.dx = dx
        self.kernel_size = kernel_size
        taps, den = self.table[kernel_size]
        self.weight = (torch.FloatTensor([[taps]]) / den).to(device)          # same fp32 construction as the reference
        r = (kernel_size - 1) // 2
        w = [0.0] * 7
        for i, t in enumerate(self.weight.flatten().tolist()):
            w[3 - r + i] = t / (dx ** self.order)
        self._taps = (C.c_float * 7)(*w)

    def __call__(self, u):
        L.require_cuda(u)
        if torch.is_grad_enabled() and u.requires_grad:
            raise NotImplementedError('stand-alone filters are forward-only; differentiate through the integrators')
        x = u.detach().contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            L.call('arpde_filter1d', L.ptr(x), L.ptr(y), x.numel() // x.shape[-1], x.shape[-1], self._taps, L.stream())
        return y


class GradFilter1D(_Filter1D):
    table, order, what = _GRAD, 1, 'GradFilter1D'

    def __init__(self, dx, kernel_size=3, device='cpu'):
        if kernel_size == 2:
            # "conditional upwind" (burgerFiniteDifference.py:181-182,224-244): weight [-1, 1, 0]; the method computes the backward
            # and the forward difference, selects by the sign of u -- and then returns the forward difference (u2).  Reproduced.
            self.dx, self.kernel_size = dx, 2
            self.weight = torch.FloatTensor([[[-1, 1, 0]]]).to(device)
            return
        super().__init__(dx, kernel_size, device)

    def __call__(self, u):
        if self.kernel_size == 2:
            return self.conditionalUpwind(u)
        return super().__call__(u)

    def conditionalUpwind(self, u):
        L.require_cuda(u)
        return (torch.roll(u, -1, -1) - u) / self.dx


class FluxWENOFilter1D(object):
    """3rd-order WENO flux derivative of u^2/2 (burgerFiniteDifference.py:89-166), restated with torch.roll:
    edge flux e[i] = w1 (-f[i-2]/2 + 3 f[i-1]/2) + w2 (f[i] + f[i+1])/2 with the smoothness weights of :150-160, its mirror image
    for the other wind direction, and the upwind selection by the sign of u."""

    def __init__(self, dx, limiter=1, device='cpu'):
        self.dx = dx

    @staticmethod
    def _edge(f):
        fm2, fm1, fp1 = torch.roll(f, 2, -1), torch.roll(f, 1, -1), torch.roll(f, -1, -1)
        e1 = -0.5 * fm2 + 1.5 * fm1
        e2 = 0.5 * f + 0.5 * fp1
        b1 = torch.pow(fm1 - fm2, 2)
        b2 = torch.pow(fp1 - f, 2)
        eps = 1e-6
        w1 = 1. / (3 * (eps + b1) ** 2)
        w2 = 2. / (3 * (eps + b2) ** 2)
        tot = w1 + w2
        return (w1 / tot) * e1 + (w2 / tot) * e2          # periodic: edge[L] == edge[0]

    def __call__(self, u):
        L.require_cuda(u)
        flux = u ** 2 / 2.0
        e = self._edge(flux)
        er = torch.roll(torch.flip(self._edge(torch.flip(flux, dims=[-1])), dims=[-1]), 1, -1)      # E[i] = e'[(L - i) mod L]
        g = (torch.roll(e, -1, -1) - e) / self.dx
        gr = (torch.roll(er, -1, -1) - er) / self.dx
        return torch.where(u < 0, g, gr)


class LaplaceFilter1D(_Filter1D):
    table, order, what = _LAP, 2, 'LaplaceFilter1D'


class FourthOrderFilter1D(_Filter1D):
    table, order, what = _D4, 4, 'FourthOrderFilter1D'


class _Filter2D(object):
    def _apply(self, u, w9):
        L.require_cuda(u)
        if torch.is_grad_enabled() and u.requires_grad:
            raise NotImplementedError('stand-alone filters are forward-only; differentiate through the integrators')
        x = u.detach().contiguous()
        y = torch.empty_like(x)
        H, W = x.shape[-2:]
        with torch.cuda.device(x.device):
            L.call('arpde_filter2d', L.ptr(x), L.ptr(y), x.numel() // (H * W), H, W, w9, L.stream())
        return y


class LapLaceFilter2d(_Filter2D):
    def __init__(self, dx, kernel_size=3, device='cpu'):
        if kernel_size not in (3, 5):
            raise NotImplementedError('LapLaceFilter2d: kernel_size %s not supported (choices [3, 5])' % kernel_size)
        self.dx = dx
        self.kernel_size = kernel_size
        w = torch.FloatTensor([[[[1, 2, 1], [-2, -4, -2], [1, 2, 1]]]]).to(device) / 4.
        w3 = w + torch.transpose(w, -2, -1)
        print(w3)                                                 # the reference prints it (:29)
        if kernel_size == 3:
            self.weight = w3
            self._w9 = (C.c_float * 9)(*[t / (dx ** 2) for t in self.weight.flatten().tolist()])
        else:                                                     # 4th-order cross (:31-35)
            self.weight = torch.FloatTensor([[[[0, 0, -1, 0, 0], [0, 0, 16, 0, 0], [-1, 16, -60, 16, -1], [0, 0, 16, 0, 0],
                                               [0, 0, -1, 0, 0]]]]).to(device) / 12.
        self._taps = self.weight[0, 0].tolist()

    def torch_call(self, u):
        """reference-equivalent roll path (any kernel size, differentiable)"""
        return _roll_filter(u, self._taps, (-2, -1)) / (self.dx ** 2)

    def __call__(self, u):
        if self.kernel_size == 3:
            return self._apply(u, self._w9)
        L.require_cuda(u)
        return self.torch_call(u)


class SobelFilter2d(_Filter2D):
    def __init__(self, dx, kernel_size=3, device='cpu'):
        if kernel_size not in (3, 5):
            raise NotImplementedError('SobelFilter2d: kernel_size %s not supported (choices [3, 5])' % kernel_size)
        self.dx = dx
        self.kernel_size = kernel_size
        if kernel_size == 3:
            self.weight_h = torch.FloatTensor([[[[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]]]).to(device) / 8.
            self.weight_v = self.weight_h.transpose(-1, -2)
            self.weight = torch.cat((self.weight_h, self.weight_v), 0)
            self._wh = (C.c_float * 9)(*[t / dx for t in self.weight_h.flatten().tolist()])
            self._wv = (C.c_float * 9)(*[t / dx for t in self.weight_v.contiguous().flatten().tolist()])
        else:                                                     # smoothed 4th-order difference (:74-78); no .weight, as in the reference
            self.weight_h = torch.FloatTensor([[[[1, -8, 0, 8, -1], [2, -16, 0, 16, -2], [3, -24, 0, 24, -3], [2, -16, 0, 16, -2],
                                                 [1, -8, 0, 8, -1]]]]).to(device) / (9 * 12.)
            self.weight_v = self.weight_h.transpose(-1, -2)
        self._th = self.weight_h[0, 0].tolist()
        self._tv = self.weight_v[0, 0].tolist()

    def torch_grad_h(self, u):
        return _roll_filter(u, self._th, (-2, -1)) / self.dx

    def torch_grad_v(self, u):
        return _roll_filter(u, self._tv, (-2, -1)) / self.dx

    def grad_h(self, u):
        if self.kernel_size == 3:
            return self._apply(u, self._wh)
        L.require_cuda(u)
        return self.torch_grad_h(u)

    def grad_v(self, u):
        if self.kernel_size == 3:
            return self._apply(u, self._wv)
        L.require_cuda(u)
        return self.torch_grad_v(u)

    def __call__(self, u):
        """(B, C, 2, H, W): [:, :, 0] = grad_h, [:, :, 1] = grad_v"""
        return torch.stack([self.grad_h(u), self.grad_v(u)], 2)


class Burger2DIntegrate(object):
    def __init__(self, dx, nu=1.0, grad_kernels=[3, 3], device='cpu'):
        self.nu = nu
        self.dx = dx
        self.grad1 = SobelFilter2d(dx, kernel_size=grad_kernels[0], device=device)
        self.grad2 = LapLaceFilter2d(dx, kernel_size=grad_kernels[1], device=device)

        self._fused = tuple(grad_kernels[:2]) == (3, 3)

    def _rhs_torch(self, uv):
        """R(u, v) of burger2DFiniteDifference.py:185-203 on the roll path (5x5 stencils)."""
        u, v = uv[:, :1], uv[:, 1:]
        g1, g2 = self.grad1, self.grad2
        ru = -g1.torch_grad_h(0.5 * u ** 2) - v * g1.torch_grad_v(u) + self.nu * g2.torch_call(u)
        rv = -u * g1.torch_grad_h(v) - g1.torch_grad_v(0.5 * v ** 2) + self.nu * g2.torch_call(v)
        return torch.cat([ru, rv], dim=1)

    def backwardEuler(self, uPred, uPred0, dt):
        if self._fused:
            return F_.integrate_burgers2d(uPred, uPred0, dt, self.dx, self.nu, F_.BE)
        L.require_cuda(uPred, uPred0)
        return uPred0 + dt * self._rhs_torch(uPred)

    def crankNicolson(self, uPred, uPred0, dt):
        if self._fused:
            return F_.integrate_burgers2d(uPred, uPred0, dt, self.dx, self.nu, F_.CN)
        L.require_cuda(uPred, uPred0)
        return uPred0 + 0.5 * dt * (self._rhs_torch(uPred) + self._rhs_torch(uPred0))


class BurgerIntegrate(object):
    def __init__(self, dx, nu=1.0, grad_kernels=[3, 3], device='cpu'):
        self.nu = nu
        self.dx = dx
        self.k = (grad_kernels[0], grad_kernels[1])
        self.grad1 = GradFilter1D(dx, kernel_size=grad_kernels[0], device=device)
        self.grad2 = LaplaceFilter1D(dx, kernel_size=grad_kernels[1], device=device)
        self.fluxgrad = FluxWENOFilter1D(dx, limiter=1, device=device)

    def _lap_torch(self, u):
        taps, den = _LAP[self.k[1]]
        w = (torch.FloatTensor(taps) / den).tolist()
        return _roll_filter(u, w, (-1,)) / (self.dx ** 2)

    def backwardEuler(self, uPred, uPred0, dt):
        if self.k[0] == 2:        # upwind gradient: roll path (the fused kernel's adjoint assumes antisymmetric gradient taps)
            L.require_cuda(uPred, uPred0)
            return uPred0 + dt * (-0.5 * (0.5 * self.grad1(torch.pow(uPred, 2))) + self.nu * self._lap_torch(uPred))
        return F_.integrate_1d(F_.SYS_BURGERS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], 5, F_.BE)

    def crankNicolson(self, uPred, uPred0, dt):
        if self.k[0] == 2:
            L.require_cuda(uPred, uPred0)
            return uPred0 - 0.5 * dt * ((uPred * self.grad1(uPred) + uPred0 * self.grad1(uPred0))
                                        - self.nu * (self._lap_torch(uPred) + self._lap_torch(uPred0)))
        return F_.integrate_1d(F_.SYS_BURGERS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], 5, F_.CN)

    def crankNicolsonLimiter(self, uPred, uPred0, dt):
        """Crank-Nicolson with the WENO flux derivative (burgerFiniteDifference.py:304-322); roll path."""
        L.require_cuda(uPred, uPred0)
        return uPred0 - 0.5 * dt * ((self.fluxgrad(uPred) + self.fluxgrad(uPred0)) - self.nu * (self._lap_torch(uPred) + self._lap_torch(uPred0)))


class KSIntegrate(object):
    def __init__(self, dx, nu=1.0, grad_kernels=[3, 3, 5], device='cpu'):
        self.nu = nu
        self.dx = dx
        self.k = tuple(grad_kernels)
        self.grad1 = GradFilter1D(dx, kernel_size=grad_kernels[0], device=device)
        self.grad2 = LaplaceFilter1D(dx, kernel_size=grad_kernels[1], device=device)
        self.grad4 = FourthOrderFilter1D(dx, kernel_size=grad_kernels[2], device=device)

    def backwardEuler(self, uPred, uPred0, dt):
        return F_.integrate_1d(F_.SYS_KS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], self.k[2], F_.BE)

    def crankNicolson(self, uPred, uPred0, dt):
        return F_.integrate_1d(F_.SYS_KS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], self.k[2], F_.CN)
