"""One-launch Adam for the reference's training loops.

`Adam` is a drop-in for `torch.optim.Adam` as the reference builds it (2D-Burgers-SWAG/main.py:219-224: two parameter groups with
their own learning rates, `weight_decay=0.0`, stepped once per BPTT window, `ExponentialLR` on top): it *is* a
`torch.optim.Adam` -- parameter groups, `state_dict()` / `load_state_dict()` (nn/swag.py:207-208 checkpoints the optimizer), LR
schedulers and `zero_grad()` are inherited -- and only `step()` is replaced by `arpde_adam_step`, one kernel over all tensors
instead of torch's ~14 multi-tensor launches.  Same fp32 update arithmetic, bias corrections in Python doubles like torch."""
import ctypes as C
import math

import torch

from . import _lib as L


class Adam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError('arpde_b200.optim.Adam: amsgrad is not used by the reference and not implemented')
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, foreach=False, fused=False)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        by_cfg = {}
        # first pass: validate everything, so that a failure leaves every step counter and moment untouched
        for group in self.param_groups:
            if group.get('maximize') or group.get('amsgrad'):
                raise NotImplementedError('arpde_b200.optim.Adam: maximize / amsgrad are not implemented')
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError('arpde_b200.optim.Adam updates dense fp32 CUDA parameters only; there is no CPU path')
                st = self.state.get(p, {})
                if not p.is_contiguous() or any(k in st and not st[k].is_contiguous() for k in ('exp_avg', 'exp_avg_sq')):
                    raise RuntimeError('arpde_b200.optim.Adam needs contiguous parameters')
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            lr = float(group['lr'])
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:                      # same lazy state as torch.optim.Adam (non-capturable: step is a CPU scalar)
                    state['step'] = torch.tensor(0.0, dtype=torch.float32)
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state['step'] += 1
                t = float(state['step'])
                bc1, bc2 = 1 - beta1 ** t, 1 - beta2 ** t
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                key = (p.device, beta1, beta2, float(group['eps']))
                by_cfg.setdefault(key, []).append((p, g, state['exp_avg'], state['exp_avg_sq'], lr / bc1, math.sqrt(bc2), float(group['weight_decay'])))
        for (dev, beta1, beta2, eps), items in by_cfg.items():
            n = len(items)
            f64 = lambda k: (C.c_double * n)(*[it[k] for it in items])
            with torch.cuda.device(dev):
                L.call('arpde_adam_step', L.ptr_array([it[0].data_ptr() for it in items]), L.ptr_array([it[1].data_ptr() for it in items]),
                       L.ptr_array([it[2].data_ptr() for it in items]), L.ptr_array([it[3].data_ptr() for it in items]),
                       L.i64_array([it[0].numel() for it in items]), f64(4), f64(5), f64(6), n, beta1, beta2, eps, L.stream())
        return loss
