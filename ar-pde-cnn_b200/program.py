"""Compile a DenseED module tree into the flat conv-op program executed by libarpde, and run it.

The plan is plain Python data (deep-copy / pickle safe: SwagNN.sample deep-copies the model,
reference nn/swag.py:101-102); ctypes tables are rebuilt on every call from the *current*
parameter storage, so optimizer updates and `param.data = w` (swag.py:148) are always seen.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L

BN_EPS_DEFAULT = 1e-5
# data-parallel training hooks, set by the context managers of parallel.py (None = single process / plain autograd)
BN_SYNC = None       # parallel.BatchNormSync: all-reduce of the train-mode BatchNorm sums over the ranks
GRAD_SINK = None     # parallel.FlatGradBucket: parameter gradients are added straight into the flat bucket


class Plan:
    """ops: list of dicts; bufs: list of dicts {ctot, scale (log2 down-sampling vs input)}."""

    def __init__(self):
        self.ops = []
        self.bufs = []
        self.param_slots = []     # list of (module, attr) in table order
        self.bn_total = 0
        self.c_in = None
        self.c_out = None
        self.nd = None
        self.input_to_buf = False


def _is_conv(m):
    return isinstance(m, (nn.Conv1d, nn.Conv2d))


def _is_bn(m):
    return isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d))


def _flatten(module, out):
    """Depth-first list of elementary modules with dense-layer markers."""
    from .nn import densed as D
    if isinstance(module, D._DenseLayer):
        out.append(('dense_begin', module))
        for ch in module.children():
            _flatten(ch, out)
        out.append(('dense_end', module))
    elif isinstance(module, nn.Sequential):
        for ch in module.children():
            _flatten(ch, out)
    elif _is_conv(module):
        out.append(('conv', module))
    elif _is_bn(module):
        out.append(('bn', module))
    elif isinstance(module, D.ReLU):
        out.append(('relu', module))
    elif isinstance(module, D.UpsamplingNearest1d):
        out.append(('up', module))
    elif isinstance(module, D.Activation):
        out.append(('act', module))
    elif isinstance(module, nn.modules.dropout._DropoutNd):
        pass        # identity here: the fused program only runs when no dropout is active (eval mode or p == 0; nn.densed.needs_generic)
    else:
        raise NotImplementedError('module %s cannot be compiled into the fused program' % type(module).__name__)
    return out


def compile_module(module):
    """Build (and cache on the module, as plain data) the op program of an nn.Sequential sub-tree."""
    cached = module.__dict__.get('_arpde_plan')
    if cached is not None:
        return cached
    items = _flatten(module, [])
    plan = Plan()
    # current tensor: (buf id or -1, channels, log2 scale relative to the input [negative = upsampled])
    cur_buf, cur_c, cur_scale = -1, None, 0
    pend_bn, pend_relu, pend_up = None, False, False
    in_dense = None
    n = len(items)

    def lookahead_growth(idx):
        """channels that dense layers directly following item idx will append to its output buffer."""
        g, j = 0, idx + 1
        while j < n and items[j][0] == 'dense_begin':
            g += items[j][1].growth_rate
            while items[j][0] != 'dense_end':
                j += 1
            j += 1
        return g

    def add_param(mod, attr):
        plan.param_slots.append((mod, attr))
        return len(plan.param_slots) - 1

    for idx, (kind, m) in enumerate(items):
        if kind == 'dense_begin':
            if cur_buf < 0:
                # stand-alone dense layer/block (DenseED.forward_test): the input is first copied into the
                # concat buffer, then the layers append their channels in place
                c0 = m.norm1.num_features
                plan.bufs.append(dict(ctot=c0 + lookahead_growth(idx - 1), scale=0))
                cur_buf, cur_c, cur_scale = 0, c0, 0
                plan.c_in = c0
                plan.input_to_buf = True
            in_dense = m
        elif kind == 'dense_end':
            in_dense = None
        elif kind == 'bn':
            if pend_bn is not None or pend_relu or pend_up:
                raise NotImplementedError('unsupported layer pattern before BatchNorm')
            pend_bn = m
        elif kind == 'relu':
            if pend_bn is None or pend_relu:
                raise NotImplementedError('ReLU must directly follow a BatchNorm')
            pend_relu = True
        elif kind == 'up':
            if pend_up:
                raise NotImplementedError('two upsampling layers in a row')
            pend_up = True
        elif kind == 'act':
            if not plan.ops or pend_bn is not None or pend_up:
                raise NotImplementedError('activation must directly follow a convolution')
            if idx != n - 1:
                raise NotImplementedError('activation is only supported as the output activation')
            plan.ops[-1]['act'] = m.act_id
        elif kind == 'conv':
            if pend_bn is not None and not pend_relu:
                raise NotImplementedError('BatchNorm without ReLU in front of a convolution')
            k, s = m.kernel_size[-1], m.stride[-1]
            nd = 1 if isinstance(m, nn.Conv1d) else 2
            plan.nd = nd if plan.nd is None else plan.nd
            if cur_c is None:
                cur_c = m.in_channels
                plan.c_in = cur_c
            if m.in_channels != cur_c:
                raise ValueError('channel mismatch while compiling: conv expects %d, tensor has %d' % (m.in_channels, cur_c))
            op = dict(in_buf=cur_buf, cin=m.in_channels, cout=m.out_channels, kh=(k if nd == 2 else 1), kw=k, stride=s,
                      up=1 if pend_up else 0, act=0, w_param=add_param(m, 'weight'), bn_param=-1, bn_off=0)
            if pend_bn is not None:
                if pend_bn.num_features != m.in_channels:
                    raise ValueError('BatchNorm/conv channel mismatch')
                op['bn_param'] = add_param(pend_bn, 'weight')
                for attr in ('bias', 'running_mean', 'running_var', 'num_batches_tracked'):
                    add_param(pend_bn, attr)
                op['bn_off'] = plan.bn_total
                op['bn_eps'] = pend_bn.eps
                op['bn_momentum'] = pend_bn.momentum
                plan.bn_total += m.in_channels
            out_scale = cur_scale - (1 if pend_up else 0) + (1 if s == 2 else 0)
            if in_dense is not None:
                if s != 1 or pend_up:
                    raise NotImplementedError('dense layers are stride-1 convolutions')
                op['out_buf'] = cur_buf
                op['out_coff'] = cur_c
                cur_c = cur_c + m.out_channels
                if cur_c > plan.bufs[cur_buf]['ctot']:
                    raise RuntimeError('internal: dense buffer too small')
            else:
                g = lookahead_growth(idx)
                plan.bufs.append(dict(ctot=m.out_channels + g, scale=out_scale))
                op['out_buf'] = len(plan.bufs) - 1
                op['out_coff'] = 0
                cur_buf, cur_c, cur_scale = op['out_buf'], m.out_channels, out_scale
            plan.ops.append(op)
            pend_bn, pend_relu, pend_up = None, False, False
    if pend_bn is not None or pend_up:
        raise NotImplementedError('program must end with a convolution (or an output activation)')
    if not plan.ops:
        raise NotImplementedError('no convolution found')
    # the last conv writes the network output; a trailing dense block would need the concat buffer as output
    last = plan.ops[-1]
    if last['out_coff'] != 0 or plan.bufs[last['out_buf']]['ctot'] != last['cout']:
        plan.out_is_buf = True            # output = a full concat buffer (stand-alone dense layer / block)
        plan.c_out = cur_c
    else:
        plan.out_is_buf = False
        plan.bufs.pop(last['out_buf'])    # by construction it is the last buffer
        last['out_buf'] = -2
        plan.c_out = last['cout']
    plan.out_scale = cur_scale
    module.__dict__['_arpde_plan'] = plan
    return plan


def _spatial(plan, H, W, scale):
    if scale >= 0:
        if (plan.nd == 2 and H % (1 << scale)) or W % (1 << scale):
            raise ValueError('input size %dx%d is not divisible by %d' % (H, W, 1 << scale))
        return (H >> scale if plan.nd == 2 else 1), W >> scale
    return (H << -scale if plan.nd == 2 else 1), W << -scale


def layout(plan, N, H, W):
    """Buffer geometry for a batch: list of (offset, stats_off, ctot, h, w), total floats, total stat doubles."""
    bufs, off, soff = [], 0, 0
    for bi, b in enumerate(plan.bufs):
        h, w = _spatial(plan, H, W, b['scale'])
        bufs.append((off, soff, b['ctot'], h, w))
        per_sample = b['ctot'] * h * w
        # a buffer with one producer and one consumer may be handed over in the consumer's tensor-core operand format during rollouts
        # (csrc/conv_z.cu), which is a few percent larger than fp32 NCHW: reserve the room
        readers = [o for o in plan.ops if o['in_buf'] == bi]
        if (plan.nd == 2 and sum(o['out_buf'] == bi for o in plan.ops) == 1 and len(readers) == 1 and readers[0]['kw'] == 3 and
                readers[0]['stride'] == 2 and not readers[0]['up'] and readers[0]['bn_param'] >= 0):
            per_sample = max(per_sample, int(L.lib.arpde_conv_z_floats(b['ctot'], h, w)))
        off += N * per_sample
        off = (off + 3) & ~3          # keep every buffer 16-byte aligned
        soff += 2 * b['ctot']
    # BatchNorm directly on the network input (stand-alone _Transition / LastTransUp): tail slot for its batch statistics
    for o in plan.ops:
        if o['bn_param'] >= 0 and o['in_buf'] < 0:
            soff += 2 * o['cin']
            break
    return bufs, off, soff


def c_tables(plan, bufs):
    ops = (L.Op * len(plan.ops))()
    for i, o in enumerate(plan.ops):
        for f in ('in_buf', 'out_buf', 'cin', 'cout', 'kh', 'kw', 'stride', 'up', 'out_coff', 'act', 'w_param', 'bn_param', 'bn_off'):
            setattr(ops[i], f, o[f])
    cb = (L.Buf * max(1, len(bufs)))()
    for i, (off, soff, ctot, h, w) in enumerate(bufs):
        cb[i].offset, cb[i].stats_off, cb[i].ctot, cb[i].h, cb[i].w = off, soff, ctot, h, w
    return ops, cb


def param_tensors(plan):
    return [getattr(mod, attr) for mod, attr in plan.param_slots]


def bn_hyper(plan):
    eps = {o.get('bn_eps', BN_EPS_DEFAULT) for o in plan.ops if o['bn_param'] >= 0}
    mom = {o.get('bn_momentum', 0.1) for o in plan.ops if o['bn_param'] >= 0}
    if len(eps) > 1 or len(mom) > 1:
        raise NotImplementedError('all BatchNorm layers must share eps and momentum')
    m = mom.pop() if mom else 0.1
    if m is None:
        raise NotImplementedError('BatchNorm momentum=None (cumulative average) is not supported')
    return (eps.pop() if eps else BN_EPS_DEFAULT), m


def _in_view(x):
    """input as (4D tensor whose channel/spatial dims are compact, batch stride)."""
    if x.dtype != torch.float32:
        raise TypeError('arpde_b200 computes in float32, got %s' % x.dtype)
    x4 = x.unsqueeze(2) if x.dim() == 3 else x
    if x4.dim() != 4:
        raise ValueError('expected [N,C,L] or [N,C,H,W] input, got %s' % (tuple(x.shape),))
    N, Cc, H, W = x4.shape
    st = x4.stride()
    compact = (W == 1 or st[3] == 1) and (H == 1 or st[2] == W) and (Cc == 1 or st[1] == H * W)
    if not compact or (N > 1 and st[0] < Cc * H * W):
        x4 = x4.contiguous()
        st = x4.stride()
    return x4, (st[0] if N > 1 else Cc * H * W)


class _Forward(torch.autograd.Function):
    """(x, *float parameters) -> y through arpde_densed_forward; backward through arpde_densed_backward."""

    @staticmethod
    def forward(ctx, plan, training, grad_mode, x, *params):
        L.require_cuda(x, *params)
        tensors = param_tensors(plan)
        x4, x_bs = _in_view(x.detach())
        N, Cin, H, W = x4.shape
        if Cin != plan.c_in:
            raise ValueError('model expects %d input channels, got %d' % (plan.c_in, Cin))
        dev = x.device
        bufs, work_len, stats_len = layout(plan, N, H, W)
        ops_c, bufs_c = c_tables(plan, bufs)
        ho, wo = _spatial(plan, H, W, plan.out_scale)
        work = torch.empty((max(work_len, 1),), device=dev, dtype=torch.float32)
        ss = torch.empty((max(2 * plan.bn_total, 1),), device=dev, dtype=torch.float32)
        eps, mom = bn_hyper(plan)
        stats = torch.empty((max(stats_len, 1),), device=dev, dtype=torch.float64) if training else None
        saved = torch.empty((max(2 * plan.bn_total, 1),), device=dev, dtype=torch.float32) if training else None
        pp = L.ptr_array([t.data_ptr() for t in tensors])
        gs = L.i64_array([0] * len(tensors))
        if plan.out_is_buf:
            y = None
            y_ptr, y_ctot = work.data_ptr(), 1      # never written: no op has out_buf == -2
        else:
            y = torch.empty((N, plan.c_out, ho, wo), device=dev, dtype=torch.float32)
            y_ptr, y_ctot = y.data_ptr(), plan.c_out
        if plan.input_to_buf:
            off, _, ctot, h, w = bufs[0]
            work[off:off + N * ctot * h * w].view(N, ctot, h, w)[:, :Cin].copy_(x4)
        tc_len = int(L.lib.arpde_densed_tc_scratch_floats(ops_c, len(plan.ops), 1))
        tc = torch.empty((max(tc_len, 1),), device=dev, dtype=torch.float32)
        sync = BN_SYNC if (training and BN_SYNC is not None and BN_SYNC.world_size() > 1) else None
        with torch.cuda.device(dev):
            if sync is None:
                L.call('arpde_densed_forward', ops_c, len(plan.ops), bufs_c, len(bufs), pp, gs, len(tensors), L.ptr(x4), x_bs, N, N,
                       H, W, y_ptr, y_ctot, 0, L.ptr(work), L.ptr(ss), plan.bn_total, L.ptr(stats), stats_len if training else 0,
                       L.ptr(saved), 1 if training else 0, float(mom), float(eps), L.ptr(tc), tc_len, L.stream())
            else:
                if sync.peer is not None:
                    sync.peer.bind_stream(L.stream())
                    cb, user = sync.peer.fn, sync.peer.user
                else:
                    cb, user = L.allreduce_callback(sync, stats), 0
                L.call('arpde_densed_forward_dp', ops_c, len(plan.ops), bufs_c, len(bufs), pp, gs, len(tensors), L.ptr(x4), x_bs, N, N,
                       H, W, y_ptr, y_ctot, 0, L.ptr(work), L.ptr(ss), plan.bn_total, L.ptr(stats), stats_len,
                       L.ptr(saved), 1, float(mom), float(eps), L.ptr(tc), tc_len, cb, user, sync.world_size(), L.stream())
        if plan.out_is_buf:
            off, _, ctot, h, w = bufs[plan.ops[-1]['out_buf']]
            y = work[off:off + N * ctot * h * w].view(N, ctot, h, w)
            if grad_mode and (ctx.needs_input_grad[3] or any(ctx.needs_input_grad[4:])):
                raise NotImplementedError('stand-alone dense layers/blocks are inference-only')
        ctx.plan, ctx.training, ctx.geom = plan, training, (N, H, W, x_bs, bufs, work_len, stats_len, eps)
        ctx.sync = sync
        has_act = any(o['act'] for o in plan.ops)
        ctx.save_for_backward(x4, work, ss, saved if saved is not None else ss, y if has_act else ss, *[t for t in tensors])
        ctx.x_shape = x.shape
        if x.dim() == 3:
            y = y.squeeze(2)
        return y

    @staticmethod
    def backward(ctx, gy):
        from . import backward as B
        return B.densed_backward(ctx, gy)


def run_module(module, x, training):
    """Forward of an nn.Sequential sub-tree (DenseED.features or a stand-alone block)."""
    plan = compile_module(module)
    tensors = param_tensors(plan)
    diff = [t for t in tensors if t.dtype.is_floating_point and isinstance(t, nn.Parameter)]
    return _Forward.apply(plan, bool(training), torch.is_grad_enabled(), x, *diff)
