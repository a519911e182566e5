"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the CPU tests).

The path shards only where it splits naturally:
  * SWAG inference: initial conditions are sharded across ranks, every rank runs ALL weight samples for
    its own ICs (so the sample moments stay local), then one all-gather of the predictive mean/variance;
  * training: the mini-batch of initial conditions is sharded, gradients are summed with ONE flat-bucket
    all-reduce per optimizer step (72 k floats for the default 2D net).
A single AR step / the spatial grid is never split (64^2..256^2 fields are far too small).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous [lo, hi) slice of n_items owned by `rank`; the first n_items % world ranks get one extra."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_shards(local, n_items, dim=0):
    """All-gather shards that were split along `dim` by shard_range into the full tensor, on every rank."""
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


def sharded_swag_rollout(run_local, ics, tsteps, n_samples, **kw):
    """ICs sharded across ranks; run_local(ic_shard, tsteps, n_samples, **kw) -> (traj, mean, var, betas) on the local
    shard (engine.swag_rollout with a generator seeded identically on every rank, so all ranks draw the same weight
    samples); returns the local trajectory and the gathered (mean, var) for ALL initial conditions."""
    rank, ws = world()
    lo, hi = shard_range(ics.shape[0], rank, ws)
    traj, mean, var, betas = run_local(ics[lo:hi], tsteps, n_samples, **kw)
    return traj, gather_shards(mean, ics.shape[0]), gather_shards(var, ics.shape[0]), betas


class FlatGradBucket:
    """One flat fp32 bucket over a parameter list: gradients are packed, summed across ranks with a single
    all-reduce and scattered back as views (the reference has no distributed training; this is the natural
    data-parallel split of its mini-batch of initial conditions)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)

    def all_reduce_(self, average=False):
        rank, ws = world()
        if ws == 1 or not self.params:
            return
        dev = self.params[0].device
        # one concatenation kernel instead of one copy per tensor (25 tensors for the default 2D net)
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in self.params])
        assert flat.device == dev and flat.numel() == self.numel
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= ws
        off = 0
        for p in self.params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
