"""torch.autograd glue over the C ABI: physics integrators and the negative log joint.

Every function takes/returns ordinary torch CUDA tensors; the arithmetic happens in libarpde.so on
the current CUDA stream.  CPU tensors raise (no fallback)."""
import torch

from . import _lib as L

CN, BE = 0, 1
SYS_BURGERS1D, SYS_KS1D = 1, 2


def _slice_view(t, nd):
    """(tensor, batch stride) such that channel/spatial dims are compact; copies only if needed."""
    if t.dtype != torch.float32:
        raise TypeError('arpde_b200 computes in float32, got %s' % t.dtype)
    shape, st = t.shape, t.stride()
    inner = 1
    ok = True
    for d in range(t.dim() - 1, 0, -1):
        if shape[d] != 1 and st[d] != inner:
            ok = False
            break
        inner *= shape[d]
    if ok and (shape[0] == 1 or st[0] >= inner):
        return t, (st[0] if shape[0] > 1 else inner)
    t = t.contiguous()
    return t, inner


class _Integrate2D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uPred, uPred0, dt, dx, nu, mode):
        L.require_cuda(uPred, uPred0)
        if uPred.dim() != 4 or uPred.shape[1] != 2 or uPred0.shape != uPred.shape:
            raise ValueError('expected uPred and uPred0 of shape [B,2,H,W], got %s and %s' % (tuple(uPred.shape), tuple(uPred0.shape)))
        a, abs_ = _slice_view(uPred.detach(), 4)
        b, bbs = _slice_view(uPred0.detach(), 4)
        N, _, H, W = a.shape
        out = torch.empty((N, 2, H, W), device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            L.call('arpde_cn_burgers2d_fwd', L.ptr(a), abs_, L.ptr(b), bbs, L.ptr(out), 0, N, H, W,
                   float(dx), float(nu), float(dt), mode, L.stream())
        ctx.save_for_backward(a, b)
        ctx.meta = (abs_, bbs, float(dt), float(dx), float(nu), mode)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        abs_, bbs, dt, dx, nu, mode = ctx.meta
        g = g.contiguous()
        N, _, H, W = a.shape
        ga = torch.empty((N, 2, H, W), device=a.device, dtype=torch.float32)
        gb = torch.empty_like(ga)
        with torch.cuda.device(a.device):
            L.call('arpde_cn_burgers2d_bwd', L.ptr(a), abs_, L.ptr(b), bbs, L.ptr(g), L.ptr(ga), L.ptr(gb), N, H, W,
                   dx, nu, dt, mode, L.stream())
        return ga, gb, None, None, None, None


class _Integrate1D(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uPred, uPred0, dt, dx, nu, mode, system, k1, k2, k4):
        L.require_cuda(uPred, uPred0)
        if uPred.shape != uPred0.shape:
            raise ValueError('uPred %s and uPred0 %s differ in shape' % (tuple(uPred.shape), tuple(uPred0.shape)))
        shape = uPred.shape
        # every channel is filtered independently (reference: u.view(-1, 1, L))
        a = uPred.detach().reshape(-1, 1, shape[-1]) if uPred.dim() != 3 or shape[1] != 1 else uPred.detach()
        b = uPred0.detach().reshape(-1, 1, shape[-1]) if uPred0.dim() != 3 or shape[1] != 1 else uPred0.detach()
        a, abs_ = _slice_view(a, 3)
        b, bbs = _slice_view(b, 3)
        N, Ln = a.shape[0], a.shape[-1]
        out = torch.empty((N, 1, Ln), device=a.device, dtype=torch.float32)
        with torch.cuda.device(a.device):
            L.call('arpde_cn_1d_fwd', system, L.ptr(a), abs_, L.ptr(b), bbs, L.ptr(out), 0, N, Ln,
                   float(dx), float(nu), float(dt), k1, k2, k4, mode, L.stream())
        ctx.save_for_backward(a, b)
        ctx.meta = (abs_, bbs, float(dt), float(dx), float(nu), mode, system, k1, k2, k4, shape)
        return out.view(shape)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        abs_, bbs, dt, dx, nu, mode, system, k1, k2, k4, shape = ctx.meta
        g = g.contiguous()
        N, Ln = a.shape[0], a.shape[-1]
        ga = torch.empty((N, 1, Ln), device=a.device, dtype=torch.float32)
        gb = torch.empty_like(ga)
        with torch.cuda.device(a.device):
            L.call('arpde_cn_1d_bwd', system, L.ptr(a), abs_, L.ptr(b), bbs, L.ptr(g), L.ptr(ga), L.ptr(gb), N, Ln,
                   dx, nu, dt, k1, k2, k4, mode, L.stream())
        return ga.view(shape), gb.view(shape), None, None, None, None, None, None, None, None


def integrate_burgers2d(uPred, uPred0, dt, dx, nu, mode=CN):
    return _Integrate2D.apply(uPred, uPred0, dt, dx, nu, mode)


def integrate_1d(system, uPred, uPred0, dt, dx, nu, k1, k2, k4=5, mode=CN):
    return _Integrate1D.apply(uPred, uPred0, dt, dx, nu, mode, system, k1, k2, k4)


def physics_sse(system, uPred, uPred0, dt, dx, nu, grad_kernels=(3, 3, 5), mode=CN):
    """Fused residual: sum((integrate(uPred,uPred0) - uPred)^2) without materialising ustar.
    Returns a float64 scalar tensor (no autograd; used by the rollout/bench diagnostics)."""
    L.require_cuda(uPred, uPred0)
    sse = torch.empty((), device=uPred.device, dtype=torch.float64)
    with torch.cuda.device(uPred.device):
        if system == 'burgers2d':
            a, abs_ = _slice_view(uPred.detach(), 4)
            b, bbs = _slice_view(uPred0.detach(), 4)
            N, _, H, W = a.shape
            L.call('arpde_cn_burgers2d_fwd', L.ptr(a), abs_, L.ptr(b), bbs, 0, L.ptr(sse), N, H, W,
                   float(dx), float(nu), float(dt), mode, L.stream())
        else:
            sysid = SYS_BURGERS1D if system == 'burgers1d' else SYS_KS1D
            a, abs_ = _slice_view(uPred.detach().reshape(-1, 1, uPred.shape[-1]), 3)
            b, bbs = _slice_view(uPred0.detach().reshape(-1, 1, uPred0.shape[-1]), 3)
            k = list(grad_kernels) + [5]
            L.call('arpde_cn_1d_fwd', sysid, L.ptr(a), abs_, L.ptr(b), bbs, 0, L.ptr(sse), a.shape[0], a.shape[-1],
                   float(dx), float(nu), float(dt), k[0], k[1], k[2], mode, L.stream())
    return sse


class _NegLogJoint(torch.autograd.Function):
    """BayesNN.calc_neg_log_joint: inputs (output, target, log_beta, *feature_params)."""

    @staticmethod
    def forward(ctx, consts, output, target, log_beta, *params):
        L.require_cuda(output, target, log_beta, *params)
        ntrain, w_shape, w_rate, b_shape, b_rate = consts
        o = output.detach().contiguous()
        t = target.detach().contiguous()
        ps = [p.detach().contiguous() for p in params]
        dev = o.device
        loss = torch.empty((1,), device=dev, dtype=torch.float32)
        scratch = torch.empty((2,), device=dev, dtype=torch.float64)
        pp = L.ptr_array([p.data_ptr() for p in ps])
        nn_ = L.i64_array([p.numel() for p in ps])
        with torch.cuda.device(dev):
            L.call('arpde_neg_log_joint_fwd', L.ptr(o), L.ptr(t), o.numel(), o.shape[0], L.ptr(log_beta.detach()),
                   pp, nn_, len(ps), float(ntrain), w_shape, w_rate, b_shape, b_rate, L.ptr(loss), L.ptr(scratch), L.stream())
        ctx.save_for_backward(o, t, log_beta.detach(), scratch, *ps)
        ctx.consts = consts
        ctx.shapes = [p.shape for p in params]
        ctx.param_refs, ctx.lb_ref = list(params), log_beta      # identities for the gradient sink (no extra memory: same tensors)
        return loss

    @staticmethod
    def backward(ctx, g):
        o, t, lb, scratch, *ps = ctx.saved_tensors
        ntrain, w_shape, w_rate, b_shape, b_rate = ctx.consts
        g = g.contiguous()
        dev = o.device
        need_o, need_t, need_lb = ctx.needs_input_grad[1:4]
        need_p = any(ctx.needs_input_grad[4:])
        go = torch.empty_like(o) if need_o else None
        gt = torch.empty_like(t) if need_t else None
        glb = torch.empty((1,), device=dev, dtype=torch.float32) if need_lb else None
        P = sum(p.numel() for p in ps)
        gp = torch.empty((P,), device=dev, dtype=torch.float32) if need_p and P else None
        pp = L.ptr_array([p.data_ptr() for p in ps])
        nn_ = L.i64_array([p.numel() for p in ps])
        with torch.cuda.device(dev):
            L.call('arpde_neg_log_joint_bwd', L.ptr(o), L.ptr(t), o.numel(), o.shape[0], L.ptr(lb), pp, nn_, len(ps),
                   float(ntrain), w_shape, w_rate, b_shape, b_rate, L.ptr(g), L.ptr(scratch), L.ptr(go), L.ptr(gt),
                   L.ptr(glb), L.ptr(gp), L.stream())
        from . import program as P_
        sink = P_.GRAD_SINK
        if sink is not None and gp is not None and all(ctx.needs_input_grad[4:]):
            # parameter gradients straight into the flat bucket (parallel.FlatGradBucket.capture()): one launch for all feature
            # tensors when they are contiguous in the bucket in this order, one for log_beta
            offs = [sink.offset_of(p) for p in ctx.param_refs]
            if None not in offs and all(offs[i] + ps[i].numel() == offs[i + 1] for i in range(len(ps) - 1)):
                sink.flat[offs[0]:offs[0] + P].add_(gp)
                lb_off = sink.offset_of(ctx.lb_ref)
                if glb is not None and lb_off is not None:
                    sink.flat[lb_off:lb_off + 1].add_(glb.view(-1))
                    glb = None
                elif glb is not None:
                    glb = glb.view(lb.shape)
                return (None, go, gt, glb, *[None for _ in ps])
        gps = []
        off = 0
        for p, need in zip(ps, ctx.needs_input_grad[4:]):
            gps.append(gp[off:off + p.numel()].view(p.shape) if (need and gp is not None) else None)
            off += p.numel()
        if glb is not None:
            glb = glb.view(lb.shape)
        return (None, go, gt, glb, *gps)


def neg_log_joint(output, target, ntrain, log_beta, feature_params, w_shape, w_rate, beta_shape, beta_rate):
    consts = (float(ntrain), float(w_shape), float(w_rate), float(beta_shape), float(beta_rate))
    return _NegLogJoint.apply(consts, output, target, log_beta, *feature_params)
