"""Auto-regressive rollout engine: S weight samples x N initial conditions x T steps in one C call.

Replaces the Python loops of the reference drivers and post scripts
(2D-Burgers-SWAG/main.py:112-135 test(), :160-179 testSample(); post/plotBARContour.py:31-73;
1D main.py :106-121, :148-163): `model(input[:, -c_in:])`, `cat([input[:, c_out:], uPred])` and the
per-step device->host copy become one trajectory tensor written in place by the last convolution.
"""
import torch

from . import _lib as L
from . import program as P


def _features_of(model):
    """DenseED.features of a DenseED / BayesNN / SwagNN."""
    m = model
    if hasattr(m, 'base'):
        m = m.base                      # SwagNN -> BayesNN
    if not hasattr(m, 'features') and hasattr(m, 'model'):
        m = m.model                     # BayesNN -> DenseED
    if not hasattr(m, 'features'):
        raise TypeError('cannot find DenseED.features in %s' % type(model).__name__)
    return m.features


def _param_table(plan, named_params, flat):
    """(ptr array, gstride array, keepalive) -- sampled tensors point into `flat` [S,P], the rest at the module."""
    tensors = P.param_tensors(plan)
    offs, off = {}, 0
    if flat is not None:
        for _, p in named_params:
            offs[id(p)] = off
            off += p.numel()
        if off != flat.shape[1]:
            raise ValueError('flat weight table has %d columns, model has %d parameters' % (flat.shape[1], off))
    ptrs, gs = [], []
    for t in tensors:
        if flat is not None and id(t) in offs:
            ptrs.append(flat.data_ptr() + 4 * offs[id(t)])
            gs.append(flat.shape[1])
        else:
            ptrs.append(t.data_ptr())
            gs.append(0)
    return L.ptr_array(ptrs), L.i64_array(gs), tensors


_SIDE_STREAMS = {}


def _side_stream(dev, k):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), k)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


@torch.no_grad()
def rollout(model, ic, tsteps, flat_weights=None, named_params=None, out=None, layerwise=False, t_start=0, total_steps=None, n_streams=1, _skip_fill=False):
    """Eval-mode AR rollout.

    model: DenseED / BayesNN / SwagNN (running BN statistics and, without flat_weights, the weights come from it)
    ic: [N, c_out, H, W] or [N, c_out, L] initial states (the history is the IC repeated, as the loaders do)
    flat_weights: optional [S, P] table from SwagNN.sample_flat (P = named_params order) -> S*N rollouts,
                  sample-major: row s*N + n uses weight sample s on initial condition n
    layerwise: force the layer-by-layer path (the KS network otherwise runs as one persistent launch; cross-check hook)
    t_start / total_steps: roll out in pieces -- `out` is the buffer of a total_steps rollout; the call advances steps
                  [t_start, t_start + tsteps) from the states already in the buffer (t_start = 0 fills in the initial history), so
                  the caller can reduce / ship the finished steps of one piece while the next one runs
    n_streams: > 1 splits the S weight samples into that many contiguous groups and rolls each group out on its own CUDA stream
                  (the rows of a group are a contiguous slice of the trajectory tensor).  The layers of one step depend on each other,
                  so a single stream leaves SMs idle at the tail of every launch; with small batches (a strong-scaled shard: 30 x 8
                  trajectories) the tail is 15-35 % of each launch and the other stream's launches fill it.
    returns traj [S*N, T+1, c_out, (H,) W] (index 0 = initial state); with total_steps the view covers all total_steps + 1 states
    """
    feats = _features_of(model)
    plan = P.compile_module(feats)
    if plan.out_is_buf or plan.input_to_buf:
        raise NotImplementedError('rollout needs a full DenseED')
    L.require_cuda(ic)
    one_d = ic.dim() == 3
    ic4 = (ic.unsqueeze(2) if one_d else ic).contiguous().float()
    N, c_out, H, W = ic4.shape
    if c_out != plan.c_out or plan.c_in % c_out:
        raise ValueError('initial condition has %d channels; model maps %d -> %d' % (c_out, plan.c_in, plan.c_out))
    if P._spatial(plan, H, W, plan.out_scale) != (H, W):
        raise ValueError('rollout needs a model whose output grid equals its input grid')
    S = 1 if flat_weights is None else flat_weights.shape[0]
    if flat_weights is not None:
        L.require_cuda(flat_weights)
        flat_weights = flat_weights.contiguous()
        if named_params is None:
            raise ValueError('named_params (the module order flat_weights was built in) is required')
    NT = S * N
    c_hist = plan.c_in
    total = int(tsteps) if total_steps is None else int(total_steps)
    if t_start < 0 or t_start + tsteps > total:
        raise ValueError('steps [%d, %d) outside the %d-step rollout' % (t_start, t_start + tsteps, total))
    if t_start > 0 and out is None:
        raise ValueError('t_start > 0 continues a rollout: pass its buffer as `out`')
    ctot = c_hist + total * c_out
    dev = ic.device
    if out is None:
        out = torch.empty((NT, ctot, H, W), device=dev, dtype=torch.float32)
    elif out.shape != (NT, ctot, H, W) or not out.is_contiguous():
        raise ValueError('out must be a contiguous [%d,%d,%d,%d] tensor' % (NT, ctot, H, W))
    if t_start == 0 and not _skip_fill:
        out.view(S, N, ctot, H, W)[:, :, :c_hist] = ic4.repeat(1, c_hist // c_out, 1, 1).unsqueeze(0)
    if n_streams > 1 and flat_weights is not None and S >= n_streams:
        cur = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event(); ready.record(cur)
        for k in range(n_streams):
            lo, hi = k * S // n_streams, (k + 1) * S // n_streams
            st = _side_stream(dev, k)
            st.wait_event(ready)
            with torch.cuda.stream(st):
                # continue (t_start semantics) on the group's rows: the history has been filled above for all rows
                rollout(model, ic, tsteps, flat_weights=flat_weights[lo:hi], named_params=named_params, out=out[lo * N:hi * N], layerwise=layerwise,
                        t_start=max(t_start, 0), total_steps=total, n_streams=1, _skip_fill=True)
            done = torch.cuda.Event(); done.record(st)
            cur.wait_event(done)
        traj = out[:, c_hist - c_out:].view(NT, total + 1, c_out, H, W)
        return traj.squeeze(3) if one_d else traj
    bufs, work_len, _ = P.layout(plan, NT, H, W)
    ops_c, bufs_c = P.c_tables(plan, bufs)
    work = torch.empty((max(work_len, 1),), device=dev, dtype=torch.float32)
    ss = torch.empty((max(2 * S * plan.bn_total, 1),), device=dev, dtype=torch.float32)
    eps, _ = P.bn_hyper(plan)
    pp, gs, keep = _param_table(plan, named_params, flat_weights)
    tc_len = int(L.lib.arpde_densed_tc_scratch_floats(ops_c, len(plan.ops), S))
    tc = torch.empty((max(tc_len, 1),), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        # a continued piece: the same buffer seen from t_start * c_out channels further on (the batch stride is unchanged)
        L.call('arpde_rollout_layerwise' if layerwise else 'arpde_rollout', ops_c, len(plan.ops), bufs_c, len(bufs), pp, gs, len(keep),
               L.ptr(out) + 4 * t_start * c_out * H * W, ctot, plan.c_in, c_hist,
               c_out, int(tsteps), NT, N, H, W, L.ptr(work), L.ptr(ss), plan.bn_total, float(eps), L.ptr(tc), tc_len, L.stream())
    traj = out[:, c_hist - c_out:].view(NT, total + 1, c_out, H, W)
    return traj.squeeze(3) if one_d else traj


@torch.no_grad()
def predictive_moments(traj, n_samples, betas=None, formula='2d', noise=None, out_ptrs=None):
    """Sample mean and predictive variance over the S weight samples of a sample-major trajectory.

    formula '2d': var = 1/mean(beta) + E[y^2] - E[y]^2      (2D-Burgers-SWAG/post/plotBARContour.py:85-104)
    formula '1d': var = mean_s(1/beta_s + y_s^2) - mean^2   (1D-Burger-SWAG/post/plotBARContour.py:123,136-139)
    out_ptrs: (mean address, variance address) of caller-owned device memory for the [n_ic, ...] results instead of fresh tensors.
    betas None -> pure sample variance (noise term 0).  noise: the noise term as a Python float (skips the device->host read of
    betas, e.g. when the moments of several pieces of one rollout are taken with the same weight samples).
    """
    L.require_cuda(traj)
    t = traj
    inner = t[0].numel()
    # rows may be a channel-slice view of the rollout buffer (batch stride > row length): no copy
    row_compact = t[0].is_contiguous()
    if not row_compact or (t.shape[0] > 1 and t.stride(0) < inner):
        t = t.contiguous()
    bstride = t.stride(0) if t.shape[0] > 1 else inner
    NT = t.shape[0]
    if NT % n_samples:
        raise ValueError('%d rollouts is not a multiple of %d samples' % (NT, n_samples))
    n_ic = NT // n_samples
    if noise is not None:
        noise = float(noise)
    elif betas is None:
        noise = 0.0
    elif formula == '2d':
        noise = float(1.0 / betas.double().mean())
    else:
        noise = float((1.0 / betas.double()).mean())
    if out_ptrs is not None:
        # the caller owns the destination (e.g. rank 0's result buffer mapped through NVLink peer memory, parallel.PeerMoments): the
        # kernel's stores ARE the transfer
        with torch.cuda.device(t.device):
            L.call('arpde_predictive_moments', L.ptr(t), bstride, inner, n_ic, n_samples, noise, int(out_ptrs[0]), int(out_ptrs[1]), L.stream())
        return None, None
    mean = torch.empty((n_ic,) + tuple(t.shape[1:]), device=t.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    with torch.cuda.device(t.device):
        L.call('arpde_predictive_moments', L.ptr(t), bstride, inner, n_ic, n_samples, noise, L.ptr(mean), L.ptr(var), L.stream())
    return mean, var


@torch.no_grad()
def swag_rollout(swag_nn, ic, tsteps, n_samples, diagCov=False, scale=1.0, generator=None, noise=None, formula='2d',
                 sample_range=None, add_noise=True, out=None):
    """testSample() of the reference in one go: draw S posterior weight samples, roll every (sample, IC)
    pair out for tsteps, return (traj [S*N, T+1, ...], mean, var, betas).

    sample_range=(lo, hi): all n_samples weight vectors are drawn (so every rank of a sample-sharded job sees the same
    draws) but only samples [lo, hi) are rolled out and reduced; betas always covers all n_samples.
    add_noise=False: var is the plain (biased) sample variance -- the caller adds the noise term (parallel.py)."""
    flat, names, numels = swag_nn.sample_flat(n_samples, scale=scale, diagCov=diagCov, generator=generator, noise=noise)
    named = list(swag_nn.base.named_parameters())
    lb_col = [i for i, (n, _) in enumerate(named) if n.endswith('log_beta')]
    betas = None
    if lb_col:
        off = sum(numels[:lb_col[0]])
        betas = flat[:, off].exp()
    lo, hi = (0, n_samples) if sample_range is None else sample_range
    if not (0 <= lo < hi <= n_samples):
        raise ValueError('sample_range %s outside [0, %d)' % ((lo, hi), n_samples))
    traj = rollout(swag_nn, ic, tsteps, flat_weights=flat[lo:hi], named_params=named, out=out)
    mean, var = predictive_moments(traj, hi - lo, betas if add_noise else None, formula)
    return traj, mean, var, betas


@torch.no_grad()
def error_metrics(pred, target, test_every=1, points=None):
    """MSE and energy-squared error of a (sample-mean) trajectory against a target, per case and saved step.

    pred [N, T+1, ...state...] (e.g. the `mean` of `swag_rollout`), target [N, Tp, ...state...] with Tp <= T // test_every + 1;
    step t of the target is compared with step t * test_every of the prediction (testMSE / testBayesianMSE of
    2D-Burgers-SWAG/post/plotMSE.py:78-85,133-137).  points = grid values per component (nel**2; default: state size / channels).
    Returns (mse, ese) as float64 CUDA tensors [N, Tp]."""
    L.require_cuda(pred); L.require_cuda(target)
    if pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise TypeError('error_metrics: float32 tensors expected')
    N, Tp = target.shape[0], target.shape[1]
    F = target[0, 0].numel()
    if pred.shape[0] != N or pred[0, 0].numel() != F or (Tp - 1) * test_every >= pred.shape[1]:
        raise ValueError('error_metrics: prediction %s does not cover target %s with test_every=%d' % (tuple(pred.shape), tuple(target.shape), test_every))
    if points is None:
        points = F // target.shape[2] if target.dim() > 3 else F
    p = pred if pred[0, 0].is_contiguous() else pred.contiguous()
    q = target if target[0, 0].is_contiguous() else target.contiguous()
    mse = torch.empty((N, Tp), device=pred.device, dtype=torch.float64)
    ese = torch.empty_like(mse)
    with torch.cuda.device(pred.device):
        L.call('arpde_error_metrics', L.ptr(p), p.stride(0), p.stride(1) * test_every, L.ptr(q), q.stride(0), q.stride(1), N, Tp, F, points,
               L.ptr(mse), L.ptr(ese), L.stream())
    return mse, ese


class GraphedStep:
    """A fixed-shape piece of work -- typically the BPTT window of a training step: forward (batch-statistics BatchNorm), time integrator,
    negative log joint and the backward pass into a `parallel.FlatGradBucket` -- captured once in a CUDA graph and replayed.

    The reference's training loop (2D-Burgers-SWAG/main.py:61-91) launches ~200 kernels per optimizer step, every one through Python; the
    captured window replays them back to back (one GPU: 6.67 -> 6.28 ms at B = 128, gradients bit-identical: the launches are the same).
    Everything `fn` reads must live in tensors that stay in place: new inputs are copied INTO the static tensors passed as `static_inputs`
    (`step.copy_inputs(...)`), the results are the tensors `fn` returned at capture time (`step.outputs`).  `fn` must not synchronise with
    the host (`.item()`, prints of device values) and must not change shapes; the optimizer step stays outside (its bias corrections are
    host scalars).  Single process only: the data-parallel BatchNorm callback is not captured.
    """

    def __init__(self, fn, static_inputs=(), warmup=3):
        self.static_inputs = tuple(static_inputs)
        dev = self.static_inputs[0].device if self.static_inputs else torch.device('cuda', torch.cuda.current_device())
        L.require_cuda(*self.static_inputs) if self.static_inputs else None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                      # warm-up off the capture stream: lazy initialisations, allocator pools
            for _ in range(max(1, warmup)):
                fn(*self.static_inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.static_inputs)

    def copy_inputs(self, *new_inputs):
        if len(new_inputs) != len(self.static_inputs):
            raise ValueError('GraphedStep: %d inputs given, %d captured' % (len(new_inputs), len(self.static_inputs)))
        for dst, src in zip(self.static_inputs, new_inputs):
            if dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError('GraphedStep: input shape / dtype differs from the captured one')
            dst.copy_(src, non_blocking=True)

    def __call__(self, *new_inputs):
        if new_inputs:
            self.copy_inputs(*new_inputs)
        self.graph.replay()
        return self.outputs
