"""Physics time integrators on libarpde -- mirrors of
  2D-Burgers-SWAG/nn/burger2DFiniteDifference.py (LapLaceFilter2d :6-54, SobelFilter2d :57-129, Burger2DIntegrate :131-211)
  1D-Burger-SWAG/nn/burgerFiniteDifference.py    (filters :8-86,:168-222, BurgerIntegrate :246-302)
  1D-KS-SWAG/nn/ksFiniteDifference.py            (filters :6-119, KSIntegrate :121-173)
crankNicolson / backwardEuler are single fused kernels, differentiable w.r.t. BOTH inputs (the loss target is
not detached in the reference, main.py:71-73).  The filter objects (.grad1/.grad2/.grad4) stay callable."""
import ctypes as C

import torch

from .. import _lib as L
from .. import functional as F_

_GRAD = {3: ([-1., 0., 1.], 2.), 5: ([1., -8., 0., 8., -1.], 12.), 7: ([-1., 9., -45., 0., 45., -9., 1.], 60.)}
_LAP = {3: ([1., -2., 1.], 1.), 5: ([-1., 16., -30., 16., -1.], 12.)}
_D4 = {5: ([1., -4., 6., -4., 1.], 1.), 7: ([-1., 12., -39., 56., -39., 12., -1.], 6.)}


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
        if kernel_size != 3:
            raise NotImplementedError('LapLaceFilter2d: only the 3x3 stencil used by main.py is supported')
        self.dx = dx
        w = torch.FloatTensor([[[[1, 2, 1], [-2, -4, -2], [1, 2, 1]]]]).to(device) / 4.
        self.weight = w + torch.transpose(w, -2, -1)
        print(self.weight)                                        # the reference prints it (:29)
        self._w9 = (C.c_float * 9)(*[t / (dx ** 2) for t in self.weight.flatten().tolist()])

    def __call__(self, u):
        return self._apply(u, self._w9)


class SobelFilter2d(_Filter2D):
    def __init__(self, dx, kernel_size=3, device='cpu'):
        if kernel_size != 3:
            raise NotImplementedError('SobelFilter2d: only the 3x3 stencil used by main.py is supported')
        self.dx = dx
        self.weight_h = torch.FloatTensor([[[[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]]]).to(device) / 8.
        self.weight_v = self.weight_h.transpose(-1, -2)
        self.weight = torch.cat((self.weight_h, self.weight_v), 0)
        self._wh = (C.c_float * 9)(*[t / dx for t in self.weight_h.flatten().tolist()])
        self._wv = (C.c_float * 9)(*[t / dx for t in self.weight_v.contiguous().flatten().tolist()])

    def grad_h(self, u):
        return self._apply(u, self._wh)

    def grad_v(self, u):
        return self._apply(u, self._wv)

    def __call__(self, u):
        """(B, C, 2, H, W): [:, :, 0] = grad_h, [:, :, 1] = grad_v"""
        return torch.stack([self.grad_h(u), self.grad_v(u)], 2)


class Burger2DIntegrate(object):
    def __init__(self, dx, nu=1.0, grad_kernels=[3, 3], device='cpu'):
        self.nu = nu
        self.dx = dx
        self.grad1 = SobelFilter2d(dx, kernel_size=grad_kernels[0], device=device)
        self.grad2 = LapLaceFilter2d(dx, kernel_size=grad_kernels[1], device=device)

    def backwardEuler(self, uPred, uPred0, dt):
        return F_.integrate_burgers2d(uPred, uPred0, dt, self.dx, self.nu, F_.BE)

    def crankNicolson(self, uPred, uPred0, dt):
        return F_.integrate_burgers2d(uPred, uPred0, dt, self.dx, self.nu, F_.CN)


class BurgerIntegrate(object):
    def __init__(self, dx, nu=1.0, grad_kernels=[3, 3], device='cpu'):
        self.nu = nu
        self.dx = dx
        self.k = (grad_kernels[0], grad_kernels[1])
        if grad_kernels[0] == 2:
            raise NotImplementedError('the conditional-upwind gradient (kernel_size=2) is not used by main.py and not supported')
        self.grad1 = GradFilter1D(dx, kernel_size=grad_kernels[0], device=device)
        self.grad2 = LaplaceFilter1D(dx, kernel_size=grad_kernels[1], device=device)

    def backwardEuler(self, uPred, uPred0, dt):
        return F_.integrate_1d(F_.SYS_BURGERS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], 5, F_.BE)

    def crankNicolson(self, uPred, uPred0, dt):
        return F_.integrate_1d(F_.SYS_BURGERS1D, uPred, uPred0, dt, self.dx, self.nu, self.k[0], self.k[1], 5, F_.CN)

    def crankNicolsonLimiter(self, uPred, uPred0, dt):
        raise NotImplementedError('the WENO flux-limited integrator is not called by any reference driver and is out of scope')


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
