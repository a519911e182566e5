"""Backward of the fused DenseED program (autograd glue for program._Forward)."""
import torch

from . import _lib as L
from . import program as P


def densed_backward(ctx, gy):
    plan, training = ctx.plan, ctx.training
    N, H, W, x_bs, bufs, work_len, stats_len, eps = ctx.geom
    x4, work, ss, saved, y_saved, *tensors = ctx.saved_tensors
    if plan.out_is_buf or plan.input_to_buf:
        raise NotImplementedError('stand-alone dense layers/blocks are inference-only')
    dev = x4.device
    gy4 = gy.unsqueeze(2) if gy.dim() == 3 else gy
    gy4 = gy4.contiguous()
    ops_c, bufs_c = P.c_tables(plan, bufs)
    # BatchNorm directly on the network input: its gamma/beta gradients come out of the input-gradient pass
    need_x = ctx.needs_input_grad[3] or (plan.ops[0]['bn_param'] >= 0 and plan.ops[0]['in_buf'] < 0)
    # differentiable parameters arrive in plan order: the apply() call passed exactly those
    slot_need = {}
    k = 4
    for i, t in enumerate(tensors):
        if t.dtype.is_floating_point and isinstance(getattr(*plan.param_slots[i]), torch.nn.Parameter):
            slot_need[i] = ctx.needs_input_grad[k]
            k += 1
    sizes = [tensors[i].numel() if slot_need.get(i) else 0 for i in range(len(tensors))]
    # Gradient sink (parallel.FlatGradBucket.capture()): the per-tensor gradients are laid out in a scratch that mirrors the
    # bucket's slice of these parameters, added into the bucket with ONE launch, and autograd gets None for them.
    sink, sink_lo, sink_hi, sink_offs = P.GRAD_SINK, None, None, {}
    if sink is not None:
        for i, s_ in enumerate(sizes):
            if s_:
                o_ = sink.offset_of(getattr(*plan.param_slots[i]))
                if o_ is None:
                    sink = None
                    break
                sink_offs[i] = o_
        if sink is not None and sink_offs:
            sink_lo = min(sink_offs.values())
            sink_hi = max(sink_offs[i] + sizes[i] for i in sink_offs)
        else:
            sink = None
    if sink is not None:
        flat = torch.zeros((sink_hi - sink_lo,), device=dev, dtype=torch.float32)
        gptrs = [flat.data_ptr() + 4 * (sink_offs[i] - sink_lo) if sizes[i] else 0 for i in range(len(sizes))]
    else:
        flat = torch.empty((max(sum(sizes), 1),), device=dev, dtype=torch.float32)
        gptrs, off = [], 0
        for s in sizes:
            gptrs.append(flat.data_ptr() + 4 * off if s else 0)
            off += s
    if not training:
        # eval-mode backward (rare): the BN "saved" statistics are the running ones
        saved = torch.empty((max(2 * plan.bn_total, 1),), device=dev, dtype=torch.float32)
        for o in plan.ops:
            if o['bn_param'] >= 0:
                rm, rv = tensors[o['bn_param'] + 2], tensors[o['bn_param'] + 3]
                saved[o['bn_off']:o['bn_off'] + o['cin']] = rm
                saved[plan.bn_total + o['bn_off']:plan.bn_total + o['bn_off'] + o['cin']] = torch.rsqrt(rv + eps)
    # scratch sizes
    t1 = t2 = wmax = cmax = 1
    h, w = H, W
    for o in plan.ops:
        if o['in_buf'] < 0:
            hin, win = H, W
        else:
            _, _, _, hin, win = bufs[o['in_buf']]
        one_d = hin == 1 and o['kh'] == 1
        hc, wc = (hin if one_d else hin << o['up']), win << o['up']
        t1 = max(t1, N * o['cin'] * hc * wc)
        t2 = max(t2, N * o['cin'] * hin * win)
        wmax = max(wmax, o['cin'] * o['cout'] * o['kh'] * o['kw'])
        cmax = max(cmax, o['cin'])
    t2 = max(t2, gy4.numel())
    T1 = torch.empty((t1,), device=dev, dtype=torch.float32)
    T2 = torch.empty((t2,), device=dev, dtype=torch.float32)
    wT = torch.empty((wmax,), device=dev, dtype=torch.float32)
    d64 = torch.empty((max(wmax, 2 * cmax),), device=dev, dtype=torch.float64)
    gwork = torch.empty((max(work_len, 1),), device=dev, dtype=torch.float32)
    gx = torch.empty((N, plan.c_in, H, W), device=dev, dtype=torch.float32) if need_x else None
    pp = L.ptr_array([t.data_ptr() for t in tensors])
    # the data-gradient conv swaps cin/cout: size the tensor-core weight scratch for the largest transposed layer
    # (the tensor-core weight gradient keeps its per-CTA partial sums in the same scratch, earlier on the stream)
    tc_len = max([int(L.lib.arpde_conv_fwd_tc_scratch_floats(o['cout'], o['cin'], o['kh'], o['kw'], 1)) for o in plan.ops] + [1])
    for o in plan.ops:
        hin, win = (H, W) if o['in_buf'] < 0 else bufs[o['in_buf']][3:5]
        tc_len = max(tc_len, int(L.lib.arpde_conv_wgrad_tc_scratch(N, o['cin'], hin, win, o['cout'], o['kh'], o['kw'], o['stride'], o['up'])))
    tc = torch.empty((tc_len,), device=dev, dtype=torch.float32)
    sync = getattr(ctx, 'sync', None)
    with torch.cuda.device(dev):
        if sync is None:
            L.call('arpde_densed_backward', ops_c, len(plan.ops), bufs_c, len(bufs), pp, len(tensors), L.ptr(x4), x_bs, N, H, W,
                   L.ptr(gy4), L.ptr(y_saved), L.ptr(work), L.ptr(ss), plan.bn_total, L.ptr(saved), L.ptr(gwork), L.ptr(gx),
                   L.ptr_array(gptrs), L.ptr(T1), L.ptr(T2), L.ptr(wT), L.ptr(d64), 1 if training else 0, L.ptr(tc), tc_len, L.stream())
        else:
            if sync.peer is not None:
                sync.peer.bind_stream(L.stream())
                cb, user = sync.peer.fn, sync.peer.user
            else:
                cb, user = L.allreduce_callback(sync, d64), 0
            L.call('arpde_densed_backward_dp', ops_c, len(plan.ops), bufs_c, len(bufs), pp, len(tensors), L.ptr(x4), x_bs, N, H, W,
                   L.ptr(gy4), L.ptr(y_saved), L.ptr(work), L.ptr(ss), plan.bn_total, L.ptr(saved), L.ptr(gwork), L.ptr(gx),
                   L.ptr_array(gptrs), L.ptr(T1), L.ptr(T2), L.ptr(wT), L.ptr(d64), 1 if training else 0, L.ptr(tc), tc_len,
                   cb, user, sync.world_size(), L.stream())
    grads, off = [], 0
    if sink is not None:
        sink.flat[sink_lo:sink_hi].add_(flat)
        grads = [None for i in range(len(tensors)) if i in slot_need]
    else:
        for i, t in enumerate(tensors):
            if i in slot_need:
                if slot_need[i]:
                    grads.append(flat[off:off + t.numel()].view(t.shape))
                    off += t.numel()
                else:
                    grads.append(None)
    if not ctx.needs_input_grad[3]:
        gx = None
    if gx is not None and len(ctx.x_shape) == 3:
        gx = gx.squeeze(2)
    return (None, None, None, gx, *grads)
