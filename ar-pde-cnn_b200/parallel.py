"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards only where it splits naturally:
  * SWAG inference: initial conditions are sharded across ranks, every rank runs ALL weight samples for
    its own ICs (so the sample moments stay local), then one all-gather of the predictive mean/variance.
    With fewer ICs than ranks the weight SAMPLES are sharded instead and the partial moments are combined
    with two all-reduces (Chan's pairwise update of mean and M2, i.e. sum y, sum y^2 without the cancellation);
  * training: the mini-batch of initial conditions is sharded; gradients live in ONE flat fp32 bucket that the
    backward kernels add into and a single all-reduce averages per optimizer step (72 k floats for the default
    2D net); train-mode BatchNorm uses the statistics of the GLOBAL mini-batch (per-layer all-reduce of
    (sum x, sum x^2) forward and (sum dz, sum dz*xhat) backward), so N ranks x B/N samples reproduce the
    single-GPU step at batch B -- the reference's semantics (2D-Burgers-SWAG/main.py:47-61).
A single AR step / the spatial grid is never split (64^2..256^2 fields are far too small).
"""
import contextlib

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous [lo, hi) slice of n_items owned by `rank`; the first n_items % world ranks get one extra."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local, n_items, dim=0):
    """All-gather shards that were split along `dim` by shard_range into the full tensor, on every rank.
    A rank whose shard is empty contributes a zero-length slice (padded for the collective)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(n_items, r, ws) for r in range(ws)]
    maxn = max(hi - lo for lo, hi in sizes)
    loc = local.movedim(dim, 0).contiguous()
    if loc.shape[0] < maxn:        # pad the short shards so one collective suffices
        pad = torch.zeros((maxn - loc.shape[0],) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device)
        loc = torch.cat([loc, pad], 0)
    out = torch.empty((ws,) + tuple(loc.shape), dtype=loc.dtype, device=loc.device)
    if loc.is_cuda:
        dist.all_gather_into_tensor(out.view(-1), loc.view(-1))
    else:
        dist.all_gather(list(out.unbind(0)), loc)
    parts = [out[r, :hi - lo] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, 0).movedim(0, dim)


def predictive_noise(betas, formula='2d'):
    """Noise term of the predictive variance from the precisions of ALL weight samples:
    '2d': 1/mean(beta) (2D-Burgers-SWAG/post/plotBARContour.py:85-104), '1d': mean(1/beta) (1D post scripts :136-139)."""
    if betas is None:
        return 0.0
    b = betas.double()
    return float(1.0 / b.mean()) if formula == '2d' else float((1.0 / b).mean())


def combine_sample_shards(mean_loc, var_loc, n_loc, n_total):
    """Merge per-rank sample moments (mean and biased variance over n_loc local weight samples) into the moments over all
    n_total samples with two all-reduces: mean = sum n_r mean_r / S, then M2 = sum (n_r var_r + n_r (mean_r - mean)^2)."""
    rank, ws = world()
    if ws == 1:
        return mean_loc, var_loc
    mean = mean_loc * (float(n_loc) / n_total)
    dist.all_reduce(mean, op=dist.ReduceOp.SUM)
    m2 = (var_loc + (mean_loc - mean) ** 2) * (float(n_loc) / n_total)
    dist.all_reduce(m2, op=dist.ReduceOp.SUM)
    return mean, m2


def sharded_swag_rollout(run_local, ics, tsteps, n_samples, mode='auto', formula='2d', gather=True, **kw):
    """SWAG rollout of n_samples weight samples x all initial conditions, split across the ranks.

    run_local(ic_shard, tsteps, (s_lo, s_hi), **kw) -> (traj, mean, var0, betas): the local rollout of weight samples
    [s_lo, s_hi) on ic_shard (engine.swag_rollout with `sample_range`, generator seeded identically on every rank so all
    ranks draw the same n_samples weight vectors); mean / var0 are the sample mean and the biased sample variance WITHOUT
    the noise term over the local samples, betas the precisions of ALL n_samples draws.
    mode 'ic'    : ICs sharded (a rank may own none), every rank runs all samples, mean/var all-gathered when `gather`
                   (gather=False returns the local shard only -- e.g. when the caller copies its shard to the host itself);
    mode 'sample': samples sharded, every rank runs all ICs, partial moments combined by all-reduce;
    mode 'auto'  : 'ic' when there are at least as many ICs as ranks, else 'sample'.
    Returns (traj_local, mean, var, betas) with var = noise(betas) + sample variance."""
    rank, ws = world()
    n_ic = ics.shape[0]
    if mode == 'auto':
        mode = 'ic' if n_ic >= ws else 'sample'
    if mode == 'ic':
        lo, hi = shard_range(n_ic, rank, ws)
        if hi > lo:
            traj, mean, var0, betas = run_local(ics[lo:hi], tsteps, (0, n_samples), **kw)
        else:
            # empty shard: no local rollout, but the rank still takes part in the gather with a zero-length slice
            traj, betas = None, None
            state = tuple(ics.shape[1:])
            mean = torch.zeros((0, tsteps + 1) + state, dtype=torch.float32, device=ics.device)
            var0 = torch.zeros_like(mean)
        if ws > 1 and n_ic < ws:
            # some shards are empty; every rank draws the same weight samples, so rank 0's precisions serve all
            b = betas.contiguous().float() if betas is not None else torch.zeros((n_samples,), dtype=torch.float32, device=ics.device)
            dist.broadcast(b, src=0)
            betas = b
        var = var0 + predictive_noise(betas, formula)
        if gather:
            mean, var = gather_shards(mean, n_ic), gather_shards(var, n_ic)
        return traj, mean, var, betas
    if mode != 'sample':
        raise ValueError("mode must be 'auto', 'ic' or 'sample'")
    lo, hi = shard_range(n_samples, rank, ws)
    if hi > lo:
        traj, mean, var0, betas = run_local(ics, tsteps, (lo, hi), **kw)
    else:
        traj, betas = None, None
        mean = torch.zeros((n_ic, tsteps + 1) + tuple(ics.shape[1:]), dtype=torch.float32, device=ics.device)
        var0 = torch.zeros_like(mean)
    if betas is None:
        betas = torch.zeros((n_samples,), dtype=torch.float32, device=ics.device)
    if ws > 1:
        src = betas.contiguous()
        dist.broadcast(src, src=0)
        betas = src
    mean, var0 = combine_sample_shards(mean, var0, hi - lo, n_samples)
    return traj, mean, var0 + predictive_noise(betas, formula), betas


def swag_rollout(swag_nn, ics, tsteps, n_samples, seed=0, diagCov=False, scale=1.0, formula='2d', mode='auto', gather=True):
    """Sharded SWAG inference on the CUDA path: engine.swag_rollout per rank behind sharded_swag_rollout.  Every rank seeds its
    generator with `seed`, so all ranks draw the same n_samples weight vectors (same device type => same Philox stream)."""
    from . import engine

    def run_local(ic_shard, T, srange):
        gen = torch.Generator(device=ic_shard.device)
        gen.manual_seed(int(seed))
        return engine.swag_rollout(swag_nn, ic_shard, T, n_samples, diagCov=diagCov, scale=scale, generator=gen, formula=formula,
                                   sample_range=srange, add_noise=False)
    return sharded_swag_rollout(run_local, ics, tsteps, n_samples, mode=mode, formula=formula, gather=gather)


class PeerMoments:
    """Predictive moments reduced STRAIGHT into rank `dst`'s result buffer over NVLink peer memory (one process per GPU, one node).

    The compute step (arpde_predictive_moments: mean / variance over the weight samples of this rank's initial conditions) is followed
    by a collective (the gather of all shards on rank dst).  Instead of reducing locally and then calling NCCL, the result buffers are
    allocated in symmetric memory (torch.distributed._symmetric_memory: every rank maps every peer's buffer), and each rank's moments
    kernel is handed the address of ITS slice of rank dst's buffer: the kernel's coalesced stores are the transfer, there is no second
    pass over the data and no staging copy.  `finish()` is the symmetric-memory barrier (device side, on the stream) after which rank
    dst may read `mean` / `var`.  `available()` is False where symmetric memory cannot be set up (then use gather_shards / NCCL)."""

    def __init__(self, n_items, item_shape, device, dst=0, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.ws = world(group)
        self.dst, self.n_items = dst, n_items
        self.inner = 1
        for d in item_shape:
            self.inner *= int(d)
        self.half = n_items * self.inner
        self.buf = symm_mem.empty(2 * self.half, dtype=torch.float32, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group=self.group)
        self.mean = self.buf[:self.half].view((n_items,) + tuple(item_shape))          # meaningful on rank dst after finish()
        self.var = self.buf[self.half:].view((n_items,) + tuple(item_shape))

    @staticmethod
    def available():
        try:
            import torch.distributed._symmetric_memory as symm_mem     # noqa: F401
            return dist.is_available() and dist.is_initialized() and dist.get_backend() == 'nccl' and torch.cuda.is_available()
        except Exception:
            return False

    def reduce_into(self, traj, n_samples, first_item, noise):
        """moments of `traj` (this rank's items [first_item, first_item + n_local)) written into rank dst's buffers"""
        from . import engine
        base = int(self.hdl.buffer_ptrs[self.dst])
        engine.predictive_moments(traj, n_samples, noise=noise, out_ptrs=(base + 4 * first_item * self.inner, base + 4 * (self.half + first_item * self.inner)))

    def finish(self):
        self.hdl.barrier()


class FlatGradBucket:
    """One flat fp32 buffer that IS the gradient storage of a parameter list (the reference has no distributed training;
    this is the natural data-parallel split of its mini-batch of initial conditions).

    `p.grad` of every parameter is a view into `flat`, so
      * autograd (or, inside `capture()`, the backward kernels themselves) accumulates straight into the bucket,
      * ONE all-reduce averages all gradients, with no pack / unpack copies,
      * `zero_()` is one memset, and the optimizer reads the views.
    Gradients are AVERAGED: every rank's loss already holds the full weight / beta priors and a likelihood scaled by
    ntrain / B_local (bayesNN.py:113-124), so the mean over ranks is the single-GPU loss at the global batch."""

    def __init__(self, params, attach=True):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.offsets = {}
        off = 0
        for p in self.params:
            self.offsets[id(p)] = off
            off += p.numel()
        self.flat = None
        self.attached = False
        if attach and self.params:
            self.attach()

    def attach(self):
        """Make p.grad a view of the bucket (values of existing gradients are kept)."""
        p0 = self.params[0]
        if self.flat is None:
            self.flat = torch.zeros((self.numel,), dtype=torch.float32, device=p0.device)
        for p in self.params:
            off = self.offsets[id(p)]
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
        self.attached = True

    def zero_(self):
        """Replacement for optimizer.zero_grad(): one memset, the views stay attached."""
        if self.flat is not None:
            self.flat.zero_()

    def offset_of(self, p):
        return self.offsets.get(id(p))

    @contextlib.contextmanager
    def capture(self):
        """While active, the backward passes of the fused DenseED program and of the negative log joint add their parameter
        gradients into the bucket with one launch each and hand `None` to autograd (instead of 25 per-tensor gradients that
        autograd then sums with one `add` kernel per tensor and use)."""
        from . import program as P
        if not self.attached:
            self.attach()
        prev = P.GRAD_SINK
        P.GRAD_SINK = self
        try:
            yield self
        finally:
            P.GRAD_SINK = prev

    def all_reduce_(self, average=True, group=None):
        rank, ws = world(group)
        if ws == 1 or not self.params:
            return
        if self.attached and all(p.grad is not None and p.grad.data_ptr() == self.flat.data_ptr() + 4 * self.offsets[id(p)] for p in self.params):
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat /= ws
            return
        # detached gradients (plain autograd, optimizer.zero_grad(set_to_none=True)): pack once, reduce, scatter views back.
        # A trailing mask counts the ranks that hold a gradient per tensor, so a parameter no rank differentiated keeps
        # grad = None and torch.optim.Adam skips it as it would on one GPU.
        dev = self.params[0].device
        pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in self.params]
        pieces.append(torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params], dtype=torch.float32, device=dev))
        flat = torch.cat(pieces)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        mask = flat[self.numel:].tolist()
        if average:
            flat /= ws
        self.flat, self.attached = flat[:self.numel], False
        for p, m in zip(self.params, mask):
            off = self.offsets[id(p)]
            p.grad = flat[off:off + p.numel()].view_as(p) if m > 0 else None


class PeerAllReduce:
    """NVLink peer-memory all-reduce of small fp64 vectors (csrc/peer.cu) for the BatchNorm sums of data-parallel training: exchange
    slots and flags live in symmetric memory, the context (peer pointers, sequence number) is owned here and handed to
    arpde_peer_allreduce_f64, which the library calls directly as its arpde_allreduce_fn -- no Python, no NCCL call per layer."""
    NSLOTS, SLOT_DOUBLES = 4, 512

    def __init__(self, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.ws = world(group)
        if self.ws > L.PEER_MAX:
            raise RuntimeError('peer all-reduce supports up to %d ranks' % L.PEER_MAX)
        nd = self.NSLOTS * self.SLOT_DOUBLES
        nflag = self.NSLOTS * self.ws
        self.buf = symm_mem.empty(nd + (nflag + 1) // 2 + 8, dtype=torch.float64, device=device)      # [slots | flags (uint32)]
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group=self.group)
        self.err = torch.zeros((1,), dtype=torch.int32, device=device)
        self.ctx = L.PeerCtx()
        for r in range(self.ws):
            self.ctx.slots[r] = int(self.hdl.buffer_ptrs[r])
            self.ctx.flags[r] = int(self.hdl.buffer_ptrs[r]) + 8 * nd
        self.ctx.world, self.ctx.rank, self.ctx.nslots, self.ctx.slot_doubles = self.ws, self.rank, self.NSLOTS, self.SLOT_DOUBLES
        self.ctx.seq = 0
        self.ctx.error_flag = self.err.data_ptr()
        torch.cuda.synchronize(device)
        self.hdl.barrier()                      # every rank's flags are zero before anyone signals
        torch.cuda.synchronize(device)
        self.fn = C.cast(L.lib.arpde_peer_allreduce_f64, L.ALLREDUCE_FN)
        self.user = C.addressof(self.ctx)

    @staticmethod
    def available(group=None):
        try:
            import torch.distributed._symmetric_memory as symm_mem     # noqa: F401
            rank, ws = world(group)
            return (dist.is_initialized() and dist.get_backend() == 'nccl' and torch.cuda.is_available() and ws > 1 and
                    torch.cuda.device_count() >= ws)          # one GPU per rank: the kernel waits for its peers' kernels
        except Exception:
            return False

    def bind_stream(self, stream_handle):
        self.ctx.stream = stream_handle

    def failed(self):
        return bool(self.err.item())


class BatchNormSync:
    """All-reduce hook for train-mode BatchNorm under data parallelism.

    `with parallel.sync_batchnorm():` makes every DenseED forward / backward inside it call back between the per-layer
    reduction kernels and the kernels that consume the sums; the callback all-reduces the fp64 sums in place (forward:
    (sum x, sum x^2) of the channels new to a BN layer; backward: (sum dz, sum dz*xhat)), so the normalisation, the
    running statistics and the BN backward are those of the global mini-batch.  Every rank must hold the same number of
    samples (the count is scaled by the world size)."""

    def __init__(self, group=None, peer=None):
        self.group = group
        self.calls = 0
        self.doubles = 0
        self.peer = peer          # PeerAllReduce: the library then calls the NVLink kernel itself instead of coming back to Python

    def world_size(self):
        return world(self.group)[1]

    def all_reduce(self, tensor):
        self.calls += 1
        self.doubles += tensor.numel()
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)


_PEER_CACHE = {}


def _peer_allreduce(group):
    """one PeerAllReduce per (device, group), created on first use; None where peer memory cannot be used"""
    if not PeerAllReduce.available(group):
        return None
    key = (torch.cuda.current_device(), id(group))
    if key not in _PEER_CACHE:
        try:
            _PEER_CACHE[key] = PeerAllReduce(torch.device('cuda', torch.cuda.current_device()), group)
        except Exception:
            _PEER_CACHE[key] = None
    return _PEER_CACHE[key]


@contextlib.contextmanager
def sync_batchnorm(group=None, peer='auto'):
    """peer: 'auto' = the NVLink peer-memory kernel when every rank has its own GPU and symmetric memory works, else
    torch.distributed.all_reduce from the Python callback; False = always the callback."""
    from . import program as P
    rank, ws = world(group)
    prev = P.BN_SYNC
    P.BN_SYNC = BatchNormSync(group, _peer_allreduce(group) if (peer and ws > 1) else None) if ws > 1 else None
    try:
        yield P.BN_SYNC
    finally:
        P.BN_SYNC = prev


def broadcast_module_(module, src=0):
    """Parameters and buffers of rank `src` to every rank (start of data-parallel training)."""
    rank, ws = world()
    if ws == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)
