"""T12: data-parallel training equivalence.  Two ranks x B/2 initial conditions -- global-batch BatchNorm
through the all-reduce hooks of arpde_densed_forward_dp / _backward_dp, gradients accumulated in the flat bucket by the backward kernels,
one averaging all-reduce, one-launch Adam -- must reproduce ONE rank x B (plain autograd): loss, every gradient, BatchNorm running
statistics and the updated weights to 1e-5 relative.  With two GPUs the collectives run over NCCL; on a one-GPU box both ranks share
the device and torch.distributed's gloo backend carries the (CUDA) tensors, which exercises the same library path."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from _util import ROOT, rel_err
import _dp_worker as W

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _run(ws, backend, target=None):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + ws + (13 if target else 0)) % 2000
    procs = [ctx.Process(target=target or W.worker, args=(r, ws, port, backend, ROOT, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    return dict(res)


def test_two_rank_step_equals_one_rank_step_at_global_batch():
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    res = _run(2, backend)
    for rank, errs in res.items():
        assert isinstance(errs, dict), errs
        # tolerance: the north_star's 1e-5 relative; gradients 2e-5 (two-step BPTT)
        assert errs['loss'] < 1e-5, (backend, rank, errs)
        assert errs['stats'] < 1e-5, (backend, rank, errs)
        assert errs['grads'] < 2e-5, (backend, rank, errs)
        assert errs['weights'] < 1e-5, (backend, rank, errs)
        assert errs['rank_consistency'] == 0.0, (backend, rank, errs)
        assert errs['peer_used'] == (torch.cuda.device_count() >= 2), errs
        if errs['peer_used']:
            assert errs['peer_vs_callback'] == 0.0 and not errs['peer_failed'], (rank, errs)


def test_sharded_swag_inference_equals_single_process():
    """ICs sharded across ranks (every rank all weight samples, all-gather of mean / variance; a rank may own no IC) and --
    fewer ICs than ranks -- weight SAMPLES sharded with the partial moments merged by all-reduce, both against one process doing the
    whole job.  Same seed on every rank => the same weight samples."""
    backend = 'nccl' if torch.cuda.device_count() >= 2 else 'gloo'
    res = _run(2, backend, target=W.infer_worker)
    for rank, errs in res.items():
        assert isinstance(errs, dict), errs
        for tag, (em, ev, eb, ntraj) in errs.items():
            assert em < 1e-6 and ev < 1e-5 and eb < 1e-6, (backend, rank, tag, em, ev, eb)
        assert errs['ic_shard'][3] == 5 * (3 if rank == 0 else 2)            # 5 ICs -> 3 + 2, all 5 samples each
        assert errs['sample_shard'][3] == (3 if rank == 0 else 2)            # 5 samples -> 3 + 2 on the single IC
        assert (errs['ic_shard_one_ic'][3] is None) == (rank == 1)           # the rank without an IC runs nothing and does not hang


def test_gradient_sink_equals_plain_autograd():
    """bucket.capture(): parameter gradients written by the backward kernels straight into the flat bucket (one add per backward call)
    == the gradients autograd accumulates tensor by tensor."""
    a = W.build(torch.device(DEV))
    b = W.build(torch.device(DEV))
    from oracle import arpde_oracle as O
    ics = O.ic_burgers2d(32, range(4)).to(DEV)
    la, ga, _, wa = W.train_step(*a, ics, dp=False, capture=True)
    lb, gb, _, wb = W.train_step(*b, ics, dp=False, capture=False)
    assert rel_err(la, lb) < 1e-6
    assert rel_err(ga, gb) < 1e-6
    assert rel_err(wa, wb) < 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='peer memory needs one GPU per rank (two processes spinning on one GPU must not wait on each other)')
def test_peer_memory_moments_equal_nccl_gather():
    """parallel.PeerMoments: each rank's moments kernel writes its shard into rank 0's symmetric-memory buffer over NVLink -- bit-identical to
    reducing locally and gathering with NCCL."""
    res = _run(2, 'nccl', target=W.peer_worker)
    for rank, errs in res.items():
        assert isinstance(errs, dict), errs
    assert res[0] == {'mean': 0.0, 'var': 0.0}, res
