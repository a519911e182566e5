#!/usr/bin/env python
"""bench.py -- headline benchmark of the ar-pde-cnn hot path on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): 2D coupled Burgers' SWAG
Bayesian inference, 30 posterior weight samples x 64 random-Fourier initial conditions x 100 auto-regressive
steps on a 64x64 periodic grid, default DenseED (blocks=[4], growth 4, init_features 96).  One "step" of this
benchmark = one complete job: draw the 30 weight samples, roll all 1920 (sample, IC) pairs out for 100 steps,
reduce the predictive mean/variance.  metric = AR steps x samples / s = S*N_ic*T / time.

    python bench.py --gpus N --steps K --warmup W          # our CUDA path (N>1: launched under torchrun)
    python bench.py --impl reference ...                   # the reference's own CPU path (its modules + testSample())

Multi-GPU (strong scaling): the FIXED 64-IC job is sharded -- rank r runs all 30 weight samples for its 64/N initial
conditions (the sample moments stay local, no data-path collective), then ONE NCCL gather of the predictive mean and
variance to rank 0.  The weak-scaling number (64 ICs per GPU) is reported next to it as `weak`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_SAMPLES, N_IC, T_STEPS, NEL = 30, 64, 100, 64
E2E_CHUNKS = 4                # pieces of the end-to-end rollout (device->host copy of piece k overlaps the rollout of piece k + 1)
MODEL_MFLOP = 267.32          # forward MFLOP per sample-step of the default 2D net
C5_MFLOP = 12233.7            # wide net (growth 32) at 256x256
KS_MFLOP = 0.51
STENCIL_BYTES_PT = {'fwd': 24, 'fwd_sse': 16, 'bwd': 40}      # per grid point, 2D


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained'), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (recipe in B200_PROFILING.md)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_ics(n, first=0, nel=NEL):
    """InitPeriodicCond2d of the reference (utils/burgerLoader2D.py:80-116), seeded per index (the benchmark's input spec)."""
    import numpy as np
    import torch
    order = 4
    x = np.linspace(0, 1, nel + 1)[:-1]
    xx, yy = np.meshgrid(x, x)
    aa, bb = np.meshgrid(np.arange(-order, order + 1), np.arange(-order, order + 1))
    k = np.stack((aa.flatten(), bb.flatten()), 1)
    kxly = (np.outer(k[:, 0], xx.flatten()) + np.outer(k[:, 1], yy.flatten())) * 2 * np.pi
    out = []
    for index in range(first, first + n):
        np.random.seed(index + 100000)
        lam = np.random.randn(2, 2, (2 * order + 1) ** 2)
        c = -1 + np.random.rand(2) * 2
        f = np.dot(lam[0], np.cos(kxly)) + np.dot(lam[1], np.sin(kxly))
        f = 2 * f / np.amax(np.abs(f), axis=1, keepdims=True) + c[:, None]
        out.append(torch.FloatTensor(f).reshape(-1, nel, nel))
    return torch.stack(out, 0)


def make_ks_ics(n, nel=96, seed=12345):
    """KS initial states: truncated Fourier series of 1D-KS-SWAG/utils/ksLoader.py:40-84 (amplitude U(2.5,3.5), three sine modes
    around the most unstable wave number), np.random.seed(12345) as SURVEY.md 8d prescribes.  Returns [n, 1, nel]."""
    import numpy as np
    import torch
    np.random.seed(seed)
    x0, x1 = 0.0, 22 * np.pi
    x = np.linspace(x0, x1, nel + 1)
    out = []
    for _ in range(n):
        a = np.random.rand() + 2.5
        k = np.linspace(1, 3, 3)
        mu_n = int((x1 - x0) / (2 * np.pi * np.sqrt(2)))
        nmode = np.round(np.abs(0.5 * np.random.randn() + mu_n))
        wl = x1 / (2 * nmode)
        c = 2 * np.pi * np.random.rand()
        lam = np.ones(3); lam[1] = 2 * np.random.randn(); lam = lam / k
        f = np.sum(lam[:, None] * np.sin(np.outer(k, x) * np.pi / wl + c), axis=0)
        f = 2 * a * (f - f.min()) / (f.max() - f.min()) - a
        out.append(torch.Tensor(f[:-1]).view(1, 1, -1))
    return torch.cat(out, 0)


def _quiet():
    import contextlib, io
    return contextlib.redirect_stdout(io.StringIO())


def build_swag(device, seed=12345, classes=None):
    """Default 2D model with torch.manual_seed init (no 2D checkpoint is shipped: reference .MISSING_LARGE_BLOBS)
    and a SWAG posterior built by 30 collect() calls on perturbed weights.  classes: (DenseED, BayesNN, SwagNN)
    -- this repo's modules by default, the reference's own for --impl reference."""
    import types
    import numpy as np
    import torch
    if classes is None:
        from arpde_b200.nn.densed import DenseED2d
        from arpde_b200.nn.bayesnn import BayesNN
        from arpde_b200.nn.swag import SwagNN
        classes = (DenseED2d, BayesNN, SwagNN)
    DenseED, BayesNN, SwagNN = classes
    torch.manual_seed(seed); np.random.seed(seed)
    args = types.SimpleNamespace(dt=0.005, device=torch.device(device), nic=3, nel=NEL)
    with _quiet():
        model = DenseED(6, 2, [4], growth_rate=4, init_features=96).to(device)
        bayes = BayesNN(args, model)
        swag = SwagNN(args, bayes, full_cov=True, max_models=30)
    g = torch.Generator(device='cpu'); g.manual_seed(seed)
    with torch.no_grad():
        # damp the random-init network so 100-step rollouts stay bounded (a trained model is contractive)
        model.features.LastTransUp.conv3.weight.mul_(0.05)
        base = [p.detach().clone() for p in bayes.parameters()]
        for it in range(30):
            for p, b in zip(bayes.parameters(), base):
                p.copy_(b + 0.002 * torch.randn(b.shape, generator=g).to(device) * (b.abs().mean() + 1e-3))
            swag.collect()
        for p, b in zip(bayes.parameters(), base):
            p.copy_(b)
    swag.eval()
    return swag, args


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own modules and its own testSample() loop, CPU
# ------------------------------------------------------------------------------------------------------------------
_REF = {}


def reference_rate(S, N, T, threads=None):
    """sample-steps/s of the reference's CPU path on a bounded sample of the workload.

    With the reference staged (baseline/_ref, or /root/reference in the build container): its own DenseED / BayesNN / SwagNN
    modules (padding shim of oracle/ref_shims.py) driven by its own main.testSample() -- sample a weight vector, roll the test
    loader out with torch.cat history updates -- followed by the sample mean / variance of post/plotBARContour.py:85-104 in numpy
    (kind 'reference').  Without it: the oracle port of the same loops (kind 'port')."""
    import numpy as np
    import torch
    torch.set_num_threads(threads or os.cpu_count())
    from oracle import ref_shims
    if ref_shims.ref_root() is not None:
        if 'swag' not in _REF:
            ref_shims.import_reference('burgers2d')
            from nn.denseEDcirc2d import DenseED
            from nn.bayesNN import BayesNN
            from nn.swag import SwagNN
            import importlib
            _REF['main'] = importlib.import_module('main')
            _REF['swag'], _REF['args'] = build_swag('cpu', classes=(DenseED, BayesNN, SwagNN))
        main, swag, args = _REF['main'], _REF['swag'], _REF['args']
        ic = make_ics(N)
        target = torch.zeros(N, T + 1, 2, NEL, NEL)
        loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(ic, target), batch_size=N, shuffle=False)
        t0 = time.perf_counter()
        with torch.no_grad(), _quiet():
            u_out, _ = main.testSample(args, swag, loader, tstep=T, n_samples=S, test_every=1)      # [N, S, T+1, 2, H, W]
        pred = u_out.numpy()
        mean = np.mean(pred, axis=1)
        var = 1.0 / swag.base.beta() + np.mean(pred ** 2, axis=1) - mean ** 2
        dt = time.perf_counter() - t0
        assert np.isfinite(var).all()
        return S * N * T / dt, dt, 'reference'
    from oracle import arpde_oracle as O
    if 'sd' not in _REF:
        _REF['sd'], _ = build_swag_port()
    sd = _REF['sd']
    ic = make_ics(N)
    t0 = time.perf_counter()
    with torch.no_grad():
        for s in range(S):        # one weight sample at a time, as testSample() does
            step = lambda inp: O.densed_forward('burgers2d', sd, inp, [4], training=False)
            O.rollout('burgers2d', step, ic, 3, T)
    dt = time.perf_counter() - t0
    return S * N * T / dt, dt, 'port'


def build_swag_port():
    """state dict of the default architecture for the oracle port (plain torch, no GPU needed)"""
    import torch
    torch.manual_seed(12345)
    sd = {}
    def conv(name, co, ci, k): sd[name + '.weight'] = torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5
    def bn(name, c):
        sd[name + '.weight'] = torch.ones(c); sd[name + '.bias'] = torch.zeros(c)
        sd[name + '.running_mean'] = torch.zeros(c); sd[name + '.running_var'] = torch.ones(c)
    f = 'features.'
    conv(f + 'In_conv1', 96, 6, 5); bn(f + 'In_conv_norm1', 96); conv(f + 'In_conv3', 48, 96, 3)
    for j in range(4):
        bn(f + 'DecBlock1.denselayer%d.norm1' % (j + 1), 48 + 4 * j); conv(f + 'DecBlock1.denselayer%d.conv1' % (j + 1), 4, 48 + 4 * j, 3)
    bn(f + 'LastTransUp.norm1', 64); conv(f + 'LastTransUp.conv1', 32, 64, 1); bn(f + 'LastTransUp.norm2', 32)
    conv(f + 'LastTransUp.conv2', 16, 32, 3); bn(f + 'LastTransUp.norm3', 16); conv(f + 'LastTransUp.conv3', 2, 16, 5)
    sd[f + 'LastTransUp.conv3.weight'] *= 0.05
    return sd, None


REF_SAMPLE = (4, 64, 16)     # S, N_ic, T of one reference-arm step (of 30 x 64 x 100): ~5 s of CPU work per step on 16 cores


def run_reference(args):
    """--impl reference: rank 0 times the CPU path; other ranks exit."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    S, N, T = REF_SAMPLE
    cores = os.cpu_count()
    kind = 'port'
    for _ in range(args.warmup):
        reference_rate(1, 8, 1)
    tot = 0.0
    for _ in range(args.steps):
        rate, dt, kind = reference_rate(S, N, T)
        tot += dt
    value = args.steps * S * N * T / tot
    how = ("the reference's own modules (nn/denseEDcirc2d.py, bayesNN.py, swag.py) driven by its main.testSample(), then numpy sample mean/variance"
           if kind == 'reference' else 'oracle port of the reference PyTorch CPU loops (reference not staged on this box)')
    sample = 'S=%d weight samples x %d ICs x T=%d steps per step (of 30x64x100), %s, torch CPU with %d threads' % (S, N, T, how, cores)
    line = {'impl': 'reference', 'metric': 'ar_steps_x_samples_per_s', 'value': value, 'unit': 'sample-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot / args.steps, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'sample-steps/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'e2e': {'value': value, 'unit': 'sample-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# training step (BASELINE config C3)
# ------------------------------------------------------------------------------------------------------------------
def train_step_probe(dev, B_total=128, weak=False, tback=2, nwarm=3, nrep=10, rank=0, world=1):
    """Training-step timing (SURVEY.md 8 rows M5 + P1 + L1; BASELINE config C3: mini-batch of 128 random Fourier ICs data-parallel over
    the GPUs).  Strong form: the B = 128 mini-batch is split, B/N per rank; weak form: 128 per rank.  Every rank runs a tback-step
    BPTT window (DenseED forward with GLOBAL-batch BatchNorm statistics, Crank-Nicolson target, negative log joint), backward with the
    parameter gradients accumulated straight into the flat bucket, ONE gradient all-reduce over NCCL and the one-launch Adam step
    (2D-Burgers-SWAG/main.py:61-91)."""
    import types
    import numpy as np
    import torch
    import torch.distributed as dist
    from arpde_b200.nn.densed import DenseED2d
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    from arpde_b200.nn.integrators import Burger2DIntegrate
    from arpde_b200 import parallel as PL
    from arpde_b200 import _lib as L
    torch.manual_seed(0); np.random.seed(0)
    a = types.SimpleNamespace(dt=0.005, device=dev, nic=3)
    with _quiet():
        model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(dev)
        bayes = BayesNN(a, model); swag = SwagNN(a, bayes, full_cov=True, max_models=30)
        integ = Burger2DIntegrate(1.0 / NEL, nu=0.005, grad_kernels=[3, 3], device=dev)
    from arpde_b200.optim import Adam
    opt = Adam([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
    from arpde_b200.data import InitPeriodicCond2d
    B = B_total if weak else B_total // world
    ic = InitPeriodicCond2d(NEL, 1 << 30, device=dev).batch(range(rank * B, rank * B + B))      # synthesised on the device (csrc/ic.cu)
    swag.train()
    bucket = PL.FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))
    ntrain = 10

    def step():
        inp = ic.repeat(1, 3, 1, 1)
        loss = 0
        bucket.zero_()
        with PL.sync_batchnorm():
            for _ in range(tback):
                u_pred = swag(inp[:, -6:])
                ustar = integ.crankNicolson(u_pred, inp[:, -2:], 0.003)
                loss = loss + swag.calc_neg_log_joint(u_pred, ustar, ntrain)
                inp = torch.cat([inp, u_pred], 1)
            with bucket.capture():
                loss.backward()
        bucket.all_reduce_()
        opt.step()
    for _ in range(nwarm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    L.lib.arpde_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        step()
    e1.record(); torch.cuda.synchronize()
    launches = int(L.lib.arpde_launch_count(1)) // nrep
    ms = e0.elapsed_time(e1) / nrep
    graph_ms = None
    if world == 1:
        # the same step with the BPTT window (forward + integrator + loss + backward into the bucket) replayed from a CUDA graph
        # (engine.GraphedStep); the optimizer step stays eager
        from arpde_b200 import engine

        def window(ic_):
            inp = ic_.repeat(1, 3, 1, 1)
            loss = 0
            bucket.zero_()
            for _ in range(tback):
                u_pred = swag(inp[:, -6:])
                ustar = integ.crankNicolson(u_pred, inp[:, -2:], 0.003)
                loss = loss + swag.calc_neg_log_joint(u_pred, ustar, ntrain)
                inp = torch.cat([inp, u_pred], 1)
            with bucket.capture():
                loss.backward()
            return loss.detach()
        try:
            gstep = engine.GraphedStep(window, (ic.clone(),))

            def step_g(i):
                gstep()
                opt.step()
            graph_ms = time_cuda(step_g, nrep)
        except Exception as e:                     # capture refused (e.g. an older driver): report the eager number only
            print('train_step: CUDA-graph capture unavailable (%s)' % repr(e)[:200], file=sys.stderr)
    if world > 1:                                   # slowest rank
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {'ms': ms, 'graph_ms': graph_ms, 'value': world * tback * B / ms * 1e3, 'unit': 'sample-steps/s', 'batch_total': world * B, 'batch_per_gpu': B, 'n_gpus': world,
            'tback': tback, 'scaling': 'weak' if weak else 'strong', 'library_launches_per_step': launches,
            'parallelism': 'mini-batch sharded x%d: global-batch BatchNorm (per-layer all-reduce of the fp64 sums: %s), gradients accumulated in one flat '
                           'bucket, one all-reduce (NCCL) per optimizer step' % (world, 'one NVLink peer-memory kernel each, csrc/peer.cu' if (
                               world > 1 and PL._peer_allreduce(None) is not None) else 'torch.distributed from the library callback'),
            'what': 'forward (batch-stat BN) + Crank-Nicolson + neg log joint + backward (tensor-core data and weight gradients) + one-launch Adam, '
                    'mean of %d steps after %d warm-up (CUDA events, max over ranks)' % (nrep, nwarm)}


def workload_config(n_gpus):
    per = N_IC // n_gpus if N_IC % n_gpus == 0 else None
    return {'workload': 'C4: 2D coupled Burgers SWAG inference, %d weight samples x %d ICs (fixed job, %s ICs per GPU) x %d AR steps, %dx%d grid, '
                        'DenseED blocks=[4] growth=4 init_features=96' % (S_SAMPLES, N_IC, per, T_STEPS, NEL, NEL),
            'samples': S_SAMPLES, 'ics_total': N_IC, 'ics_per_gpu': per, 'ar_steps': T_STEPS, 'grid': [NEL, NEL],
            'parallelism': "ic-shard x%d (strong scaling of the fixed 64-IC job) + final gather of mean/var on rank 0: every rank's moments kernel stores its shard into "
                           "rank 0's buffer over NVLink peer memory (symmetric memory; NCCL gather as fallback, see roofline.job_breakdown_ms.gather)" % n_gpus,
            'l2': 'inputs larger than L2 (per-step activations %.1f GB per GPU; stencil timings rotate 4 buffer sets, each > L2 / 4)' % (4.3 / n_gpus)}


def time_cuda(fn, nrep, nwarm=3):
    import torch
    for _ in range(nwarm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(nrep):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / nrep


def stencil_rooflines(dev, peaks):
    """K2 (fused periodic stencil + Crank-Nicolson) against the measured HBM peak with FOUR rotating buffer sets per case (SURVEY.md H5:
    no tensor of launch i+1 can be L2-resident from launch i), at the C4 rollout batch (1920 x 64x64) and at C5 (256 x 256x256):
    forward API form (24 B/pt), fused-loss form (16 B/pt, no ustar written) and the adjoint (40 B/pt)."""
    import torch
    import arpde_b200._lib as L
    out = {}
    NSETS = 4
    for tag, (B, n) in (('c4_64', (S_SAMPLES * N_IC, 64)), ('c5_256', (256, 256))):
        sets = []
        for _ in range(NSETS):
            sets.append([torch.randn((B, 2, n, n), device=dev) for _ in range(3)] + [torch.empty((B, 2, n, n), device=dev) for _ in range(3)])
        sse = torch.zeros((NSETS,), device=dev, dtype=torch.float64)
        pts = B * n * n
        bs = 2 * n * n

        def fwd(i):
            s = sets[i % NSETS]
            L.call('arpde_cn_burgers2d_fwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[3]), 0, B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())

        def fwd_sse(i):
            s = sets[i % NSETS]
            L.call('arpde_cn_burgers2d_fwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, 0, sse.data_ptr() + 8 * (i % NSETS), B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())

        def bwd(i):
            s = sets[i % NSETS]
            L.call('arpde_cn_burgers2d_bwd', L.ptr(s[0]), bs, L.ptr(s[1]), bs, L.ptr(s[2]), L.ptr(s[4]), L.ptr(s[5]), B, n, n, 1.0 / n, 0.005, 0.005, 0, L.stream())

        for name, fn in (('fwd', fwd), ('fwd_sse', fwd_sse), ('bwd', bwd)):
            ms = time_cuda(fn, 20)
            byts = pts * STENCIL_BYTES_PT[name]
            gbs = byts / (ms * 1e-3) / 1e9
            out['%s_%s' % (tag, name)] = {'ms': ms, 'bytes_per_launch': byts, 'achieved': gbs, 'frac': gbs / peaks['hbm_gbs'],
                                          'states': B, 'grid': [n, n], 'bytes_per_point': STENCIL_BYTES_PT[name], 'rotating_sets': NSETS}
        del sets
        torch.cuda.empty_cache()
    return out


def ks_probe(dev):
    """BASELINE config C2: 1D Kuramoto-Sivashinsky DenseED with SWAG, 30 posterior samples x 64 ICs x 200 AR steps on one GPU; weights and
    SWAG posterior = the reference's shipped epoch-200 checkpoint (tests/golden/ks1d_ckpt200.npz holds its state dict)."""
    import types
    import numpy as np
    import torch
    from arpde_b200 import engine
    from arpde_b200 import _lib as L
    from arpde_b200.nn.densed import DenseED1dKS
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'ks1d_ckpt200.npz'))
    K = int(z['max_models'])
    a = types.SimpleNamespace(dt=0.1, device=dev, nic=2)
    np.random.seed(0)
    with _quiet():
        model = DenseED1dKS(2, 1, [4], growth_rate=4, init_features=32).to(dev)
        swag = SwagNN(a, BayesNN(a, model), full_cov=True, max_models=K)
    swag.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('sd/')}, strict=True)
    swag.eval()
    S, N, T = 30, 64, 200
    ic = make_ks_ics(N).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(7)
    res = {}

    def job(_):
        res['out'] = engine.swag_rollout(swag, ic, T, S, generator=gen, formula='1d')
    L.lib.arpde_launch_count(1)
    job(0)
    launches = int(L.lib.arpde_launch_count(1))
    ms = time_cuda(job, 3, nwarm=1)
    traj, mean, var, betas = res['out']
    return {'value': S * N * T / ms * 1e3, 'unit': 'sample-steps/s', 'ms_per_job': ms, 'us_per_ar_step': 1e3 * ms / T, 'samples': S, 'ics': N,
            'ar_steps': T, 'grid': 96, 'tflops': S * N * T * KS_MFLOP * 1e6 / (ms * 1e-3) / 1e12, 'library_launches_per_job': launches,
            'finite': bool(torch.isfinite(mean).all().item()),
            'weights': "reference's shipped 1D-KS-SWAG/post/networks/torchSwagModel_epoch200.pth (via tests/golden/ks1d_ckpt200.npz), full-covariance samples (K = %d)" % K}


def c1_probe(dev, B=256, L_=512, tback=2, nwarm=3, nrep=10):
    """BASELINE config C1: the 1D viscous Burgers DenseED (1D-Burger-SWAG/args.py defaults: nic 5, blocks [1], growth 4, 64 initial features,
    512-point grid, mini-batch 256) -- one optimizer step of the reference's training loop (1D-Burger-SWAG/main.py:47-87): a tback-step BPTT
    window of DenseED forward (batch-statistics BatchNorm) + Crank-Nicolson target (dx = 1/513, nu = 0.0025) + negative log joint, backward,
    one-launch Adam.  ICs of the form burgerLoader.py:51-84 draws (order-3 Fourier series, numpy seed 12345)."""
    import types
    import numpy as np
    import torch
    from arpde_b200.nn.densed import DenseED1dBurgers
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    from arpde_b200.nn.integrators import BurgerIntegrate
    from arpde_b200.optim import Adam
    from arpde_b200 import parallel as PL
    from arpde_b200 import _lib as L
    torch.manual_seed(0); np.random.seed(12345)
    nic = 5
    a = types.SimpleNamespace(dt=0.005, device=dev, nic=nic)
    with _quiet():
        model = DenseED1dBurgers(nic, 1, [1], growth_rate=4, init_features=64).to(dev)
        bayes = BayesNN(a, model); swag = SwagNN(a, bayes, full_cov=True, max_models=30)
        integ = BurgerIntegrate(1.0 / (L_ + 1), nu=0.0025, grad_kernels=[3, 3], device=dev)
    opt = Adam([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
    x = np.linspace(0, 1, L_ + 1)[:-1]
    order = 3
    lam = np.random.randn(B, 2, 2 * order + 1)
    c = np.random.rand(B) - 0.5
    k = np.arange(-order, order + 1)
    kx = 2 * np.pi * k[None, :, None] * x[None, None, :]
    f = (lam[:, 0, :, None] * np.cos(kx) + lam[:, 1, :, None] * np.sin(kx)).sum(1)
    f = 2 * f / np.abs(f).max(1, keepdims=True) + 2 * c[:, None]
    ic = torch.from_numpy(f).float().to(dev).unsqueeze(1)
    swag.train()
    bucket = PL.FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))

    def step():
        inp = ic.repeat(1, nic, 1)
        loss = 0
        bucket.zero_()
        for _ in range(tback):
            u_pred = swag(inp[:, -nic:])
            ustar = integ.crankNicolson(u_pred, inp[:, -1:], 0.003)
            loss = loss + swag.calc_neg_log_joint(u_pred, ustar, 10)
            inp = torch.cat([inp, u_pred[:, :1]], 1)
        with bucket.capture():
            loss.backward()
        opt.step()
        return loss
    for _ in range(nwarm):
        loss = step()
    torch.cuda.synchronize()
    L.lib.arpde_launch_count(1)
    ms = time_cuda(lambda i: step(), nrep, nwarm=0)
    launches = int(L.lib.arpde_launch_count(1)) // nrep
    return {'ms': ms, 'value': tback * B / ms * 1e3, 'unit': 'sample-steps/s', 'batch': B, 'grid': L_, 'tback': tback, 'library_launches_per_step': launches,
            'finite': bool(torch.isfinite(loss.detach()).all()),
            'model': 'DenseED 1D Burgers (nic 5, blocks [1], growth 4, init_features 64: 13 441 parameters, 8.57 MFLOP per sample-step forward)',
            'what': 'one optimizer step = %d-step BPTT window: forward (batch-stat BN) + Crank-Nicolson + neg log joint + backward + one-launch Adam; mean of '
                    '%d steps after %d warm-up (CUDA events)' % (tback, nrep, nwarm)}


def c5_probe(dev, peaks):
    """BASELINE config C5: scaled synthetic 256x256 grid, wide DenseED (growth 32, init_features 96): AR rollout of 32 initial conditions,
    every layer but the output conv on the tensor-core kernel."""
    import torch
    from arpde_b200 import engine
    from arpde_b200.nn.densed import DenseED2d
    torch.manual_seed(5)
    with _quiet():
        m = DenseED2d(6, 2, [4], growth_rate=32, init_features=96).to(dev)
    with torch.no_grad():
        m.features.LastTransUp.conv3.weight.mul_(0.05)
    m.eval()
    N, T = 32, 4
    ic = make_ics(N, nel=256).to(dev)
    res = {}

    def job(_):
        res['traj'] = engine.rollout(m, ic, T)
    ms = time_cuda(job, 3, nwarm=1)
    ok = bool(torch.isfinite(res['traj']).all().item())
    del res
    torch.cuda.empty_cache()
    tf = N * T * C5_MFLOP * 1e6 / (ms * 1e-3) / 1e12
    return {'value': N * T / ms * 1e3, 'unit': 'sample-steps/s', 'ms_per_ar_step': ms / T, 'ics': N, 'ar_steps': T, 'grid': [256, 256], 'tflops': tf,
            'frac_of_bf16_peak': tf / peaks['bf16_tflops'], 'frac_of_3xtf32_ceiling': tf / (peaks['bf16_tflops'] / 6), 'finite': ok,
            'model': 'DenseED blocks=[4] growth=32 init_features=96 (220 576 parameters, 12.23 GFLOP per sample-step)'}


def main():
    global T_STEPS
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-probe', action='store_true', help='skip the (untimed) training-step measurement')
    ap.add_argument('--nccl-gather', action='store_true', help='gather the moments with NCCL instead of peer-memory stores from the moments kernel')
    ap.add_argument('--no-extra', action='store_true', help='skip the C2 / C5 / weak-scaling / stencil side measurements')
    ap.add_argument('--tsteps', type=int, default=T_STEPS, help='debug: shorter rollouts (the JSON line then names the reduced config)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import arpde_b200
    from arpde_b200 import engine
    from arpde_b200 import parallel as PL
    import arpde_b200._lib as L
    T_STEPS = args.tsteps

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'        # (round-1 behaviour, kept: NCCL prints its version line at WARN as well; the JSON line stays the LAST line of stdout)
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()

    # the gather of the predictive moments: fused into the moments kernel through NVLink peer memory when symmetric memory can be set up
    # (N > 1), NCCL gather otherwise
    USE_PEER = False
    if world > 1 and not args.nccl_gather and PL.PeerMoments.available():
        try:
            probe = PL.PeerMoments(2, (4,), dev)
            probe.finish()
            USE_PEER = True
        except Exception as e:                        # symmetric memory not usable on this box: NCCL gather
            print('peer-memory gather unavailable (%s): NCCL gather' % repr(e)[:200], file=sys.stderr)
    swag, _ = build_swag(dev)
    named = list(swag.base.named_parameters())
    gen = torch.Generator(device=dev); gen.manual_seed(1234)       # every rank draws the SAME weight samples
    lo, hi = PL.shard_range(N_IC, rank, world)
    n_loc = hi - lo
    n_max = max(b - a for a, b in (PL.shard_range(N_IC, r, world) for r in range(world)))
    ic_all_host = make_ics(N_IC)
    state = (T_STEPS + 1, 2, NEL, NEL)

    class Job:
        """S weight samples x `n_loc` local ICs x T steps; predictive moments; gather of the shards to rank 0."""

        def __init__(self, ic_host_shard, gathered):
            self.n = ic_host_shard.shape[0]
            self.ic_host = ic_host_shard.pin_memory()
            self.ic_dev = self.ic_host.to(dev)
            self.traj_buf = torch.empty((S_SAMPLES * self.n, 6 + 2 * T_STEPS, NEL, NEL), device=dev)
            # small shards (strong scaling: 30 x 8 trajectories per GPU at N = 8): two groups of weight samples on two streams, so that
            # the tail of every launch (15-35 % at this batch) is filled by the other group's launches (measured 0.976 -> 0.926 ms per step)
            # two weight-sample groups on two streams fill each other's launch tails at mid-size shards (measured with the PDL chain: 960 trajectories
            # 1.79 -> 1.73 ms per step, 480: 0.934 -> 0.905); at 240 one stream is faster (0.507 vs 0.550), at 1920 it makes no difference
            self.n_streams = 2 if 400 <= S_SAMPLES * self.n <= 1200 else 1
            self.gathered = gathered
            self.peer = None
            if gathered and world > 1:
                # rank 0 receives every shard into one [world, n, ...] tensor per moment (equal shards; a ragged split pads to n_max)
                self.pad = n_max
                self.g_m = torch.empty((world, n_max) + state, device=dev) if rank == 0 else None
                self.g_v = torch.empty((world, n_max) + state, device=dev) if rank == 0 else None
                if USE_PEER:
                    # NVLink peer memory: every rank's moments kernel stores its shard straight into rank 0's buffer (parallel.PeerMoments)
                    self.peer = PL.PeerMoments(N_IC, state, dev)
            n_res = N_IC if gathered else self.n
            if rank == 0 or not gathered:
                self.mean_host = torch.empty((n_res,) + state, pin_memory=True)
                self.var_host = torch.empty((n_res,) + state, pin_memory=True)
            # pipelined end-to-end form: the rollout advances in E2E_CHUNKS pieces; the moments of a finished piece are gathered and copied
            # to the host on a side stream while the next piece runs (equal shards only)
            self.pipelined = (N_IC % world == 0) and T_STEPS % E2E_CHUNKS == 0 and T_STEPS >= 2 * E2E_CHUNKS
            if self.pipelined:
                cl = T_STEPS // E2E_CHUNKS
                self.spans = [(k * cl + (0 if k == 0 else 1), (k + 1) * cl + 1) for k in range(E2E_CHUNKS)]      # saved states [lo, hi) of piece k
                self.copy_stream = torch.cuda.Stream(device=dev)
                self.host_chunks, self.g_chunks = [], []
                for lo_, hi_ in self.spans:
                    shp = (hi_ - lo_, 2, NEL, NEL)
                    if rank == 0 or not gathered:
                        self.host_chunks.append((torch.empty((n_res,) + shp, pin_memory=True), torch.empty((n_res,) + shp, pin_memory=True)))
                    if gathered and world > 1 and USE_PEER:
                        self.g_chunks.append(PL.PeerMoments(N_IC, shp, dev))
                    elif gathered and world > 1 and rank == 0:
                        self.g_chunks.append((torch.empty((world, self.n) + shp, device=dev), torch.empty((world, self.n) + shp, device=dev)))

        def run(self, ic, marks=None):
            """marks: optional list that receives CUDA events after sampling / rollout / moments / gather (job breakdown)"""
            def mark():
                if marks is not None:
                    ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append(ev)
            mark()
            flat, names, numels = swag.sample_flat(S_SAMPLES, generator=gen)
            mark()
            traj = engine.rollout(swag, ic, T_STEPS, flat_weights=flat, named_params=named, out=self.traj_buf, n_streams=self.n_streams)
            mark()
            betas = flat[:, 0].exp()
            if self.peer is not None:
                self.peer.reduce_into(traj, S_SAMPLES, lo, noise=float(1.0 / betas.double().mean()))
                mark()
                self.peer.finish()
                mark()
                return self.peer.mean, self.peer.var
            mean, var = engine.predictive_moments(traj, S_SAMPLES, betas, '2d')
            mark()
            if self.gathered and world > 1:
                if self.n < self.pad:
                    padz = torch.zeros((self.pad - self.n,) + state, device=dev)
                    mean, var = torch.cat([mean, padz]), torch.cat([var, padz])
                dist.gather(mean, list(self.g_m.unbind(0)) if rank == 0 else None, dst=0)
                dist.gather(var, list(self.g_v.unbind(0)) if rank == 0 else None, dst=0)
                mark()
                if rank == 0:
                    return self.g_m, self.g_v
            return mean, var

        def run_e2e_pipelined(self):
            """host ICs in, host predictive mean/variance out (rank 0: all ICs, as a list of time pieces [N_ic, steps, 2, H, W]);
            the device->host copy of piece k overlaps the rollout of piece k + 1"""
            ic = self.ic_host.to(dev, non_blocking=True)
            flat, names, numels = swag.sample_flat(S_SAMPLES, generator=gen)
            noise = float(1.0 / flat[:, 0].exp().double().mean())         # one device->host read per job, before the rollout is queued
            cl = T_STEPS // E2E_CHUNKS
            main = torch.cuda.current_stream()
            for k, (lo_, hi_) in enumerate(self.spans):
                traj = engine.rollout(swag, ic, cl, flat_weights=flat, named_params=named, out=self.traj_buf, t_start=k * cl, total_steps=T_STEPS,
                                      n_streams=self.n_streams)
                if self.gathered and world > 1 and USE_PEER:
                    pm = self.g_chunks[k]
                    pm.reduce_into(traj[:, lo_:hi_], S_SAMPLES, lo, noise=noise)
                    pm.finish()
                    src = (pm.mean, pm.var) if rank == 0 else None
                    mean = var = None
                else:
                    mean, var = engine.predictive_moments(traj[:, lo_:hi_], S_SAMPLES, formula='2d', noise=noise)
                if mean is None:
                    pass
                elif self.gathered and world > 1:
                    gm, gv = self.g_chunks[k] if rank == 0 else (None, None)
                    dist.gather(mean, list(gm.unbind(0)) if rank == 0 else None, dst=0)
                    dist.gather(var, list(gv.unbind(0)) if rank == 0 else None, dst=0)
                    src = (gm.view((N_IC,) + tuple(gm.shape[2:])), gv.view((N_IC,) + tuple(gv.shape[2:]))) if rank == 0 else None
                else:
                    src = (mean, var)
                if src is not None and (rank == 0 or not self.gathered):
                    ev = torch.cuda.Event(); ev.record(main)
                    self.copy_stream.wait_event(ev)
                    with torch.cuda.stream(self.copy_stream):
                        self.host_chunks[k][0].copy_(src[0], non_blocking=True)
                        self.host_chunks[k][1].copy_(src[1], non_blocking=True)
                    src[0].record_stream(self.copy_stream); src[1].record_stream(self.copy_stream)
            self.copy_stream.synchronize()
            main.synchronize()

        def run_e2e(self):
            """the call a user makes: host ICs in, host predictive mean/variance (of ALL ICs, on rank 0) out"""
            if self.pipelined:
                return self.run_e2e_pipelined()
            ic = self.ic_host.to(dev, non_blocking=True)
            mean, var = self.run(ic)
            if self.gathered and world > 1:
                if rank == 0:
                    off = 0
                    for r in range(world):
                        a, b = PL.shard_range(N_IC, r, world)
                        self.mean_host[off:off + b - a].copy_(mean[r, :b - a], non_blocking=True)
                        self.var_host[off:off + b - a].copy_(var[r, :b - a], non_blocking=True)
                        off += b - a
            else:
                self.mean_host.copy_(mean, non_blocking=True)
                self.var_host.copy_(var, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- headline: the fixed 64-IC job, sharded (strong scaling) -------------------------------------------------------
    job = Job(ic_all_host[lo:hi], gathered=True)
    for _ in range(args.warmup):
        job.run(job.ic_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.lib.arpde_launch_count(1)
    ms = timed(lambda: job.run(job.ic_dev), args.steps)
    launches = int(L.lib.arpde_launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    units = S_SAMPLES * N_IC * T_STEPS * args.steps
    value = units / (ms * 1e-3)

    # where one job spends its time (rank 0, one extra untimed job): sampling | rollout | moments | gather
    marks = []
    job.run(job.ic_dev, marks)
    torch.cuda.synchronize()
    bd = [marks[i].elapsed_time(marks[i + 1]) for i in range(len(marks) - 1)]
    breakdown = dict(zip(('sample_ms', 'rollout_ms', 'moments_ms', 'gather_ms'), bd))
    breakdown['gather'] = ('peer-memory stores of the moments kernel into rank 0 (NVLink) + symmetric-memory barrier' if USE_PEER else 'NCCL gather') if world > 1 else 'none'
    rollout_ms = breakdown['rollout_ms']

    for _ in range(max(1, args.warmup // 3)):
        job.run_e2e()
    ms_e2e = timed(job.run_e2e, args.steps)
    e2e_value = units / (ms_e2e * 1e-3)
    del job
    torch.cuda.empty_cache()

    # ---- weak-scaling form (64 ICs per GPU, every rank copies its own shard to its host), N > 1 only -------------------
    weak = None
    if world > 1 and not args.no_extra:
        wjob = Job(make_ics(N_IC, first=rank * N_IC), gathered=False)
        wjob.run(wjob.ic_dev)
        wsteps = min(args.steps, 2)
        wms = timed(lambda: wjob.run(wjob.ic_dev), wsteps)
        wms_e2e = timed(wjob.run_e2e, wsteps)
        wunits = world * S_SAMPLES * N_IC * T_STEPS * wsteps
        weak = {'value': wunits / (wms * 1e-3), 'e2e_value': wunits / (wms_e2e * 1e-3), 'unit': 'sample-steps/s', 'ics_per_gpu': N_IC, 'steps': wsteps,
                'ms_per_step': wms / wsteps, 'scaling': 'weak', 'what': '64 ICs per GPU (the job grows with N), no gather: every rank keeps its own shard of the result'}
        del wjob
        torch.cuda.empty_cache()

    # ---- per-kernel rooflines, measured live with CUDA events on the launching stream ------------------
    # (1) one full AR step of the local batch = the rollout of the job (into its preallocated trajectory buffer) / T: includes the per-rollout
    #     set-up (weight split for the tensor-core layers, BN fold: ~0.1 ms per 100 steps)
    NT = S_SAMPLES * n_loc
    step_ms = rollout_ms / T_STEPS
    conv_tflops = NT * MODEL_MFLOP * 1e6 / (step_ms * 1e-3) / 1e12
    # (1b) the tensor-core launches of the step, each timed alone through the C ABI on the rollout batch with rotating inputs (every set > L2):
    #      In_conv1 (conv_stem_kernel, the dominant launch), In_conv3 (conv_z_kernel), conv1x1 and the upsampling conv3x3 (conv_h_kernel)
    bf16 = peaks['bf16_tflops']
    kern = {}
    xin = [torch.randn((NT, 6, NEL, NEL), device=dev) for _ in range(2)]
    w1 = torch.randn((S_SAMPLES, 96, 6, 5, 5), device=dev) / 150 ** 0.5
    w3 = torch.randn((S_SAMPLES, 48, 96, 3, 3), device=dev) / 864 ** 0.5
    sc1 = torch.rand((S_SAMPLES, 96), device=dev) + 0.5
    sh1 = torch.randn((S_SAMPLES, 96), device=dev) * 0.2
    zf = int(L.lib.arpde_conv_z_floats(96, NEL, NEL))
    z = torch.empty((NT * zf,), device=dev)
    y3 = torch.empty((NT, 64, NEL // 2, NEL // 2), device=dev)
    n1 = int(L.lib.arpde_conv_fwd_tc_scratch_floats(6, 96, 5, 5, S_SAMPLES))
    n3 = int(L.lib.arpde_conv_fwd_tc_scratch_floats(96, 48, 3, 3, S_SAMPLES))
    scr1, scr3 = torch.empty((n1,), device=dev), torch.empty((n3,), device=dev)

    def k_stem(i):
        L.call('arpde_stem_pair_producer', L.ptr(xin[i % 2]), 6 * NEL * NEL, L.ptr(w1), w1.stride(0), L.ptr(sc1), L.ptr(sh1), 96, NT, 6, NEL, NEL, 96, 5,
               n_loc, L.ptr(z), z.numel(), L.ptr(scr1), n1, L.stream())

    def k_z(i):
        L.call('arpde_stem_pair_consumer', L.ptr(z), L.ptr(w3), w3.stride(0), L.ptr(y3), NT, NEL, NEL, 96, 48, 64, 0, n_loc, L.ptr(scr3), n3, L.stream())
    k1_ms = time_cuda(k_stem, 10)
    k1_flops = 2.0 * NT * NEL ** 2 * 96 * 6 * 25
    k1_tflops = k1_flops / (k1_ms * 1e-3) / 1e12
    kz_ms = time_cuda(k_z, 10)
    kz_flops = 2.0 * NT * (NEL // 2) ** 2 * 48 * 96 * 9
    z_bytes = NT * zf * 4
    kern['In_conv1 conv_stem_kernel'] = {'ms': k1_ms, 'tflops': k1_tflops, 'frac_of_bf16_peak': k1_tflops / bf16, 'frac_of_fp16x3_ceiling': k1_tflops / (bf16 / 3),
                                         'hbm_gbs': (xin[0].numel() * 4 + z_bytes) / k1_ms / 1e6}
    kern['In_conv3 conv_z_kernel'] = {'ms': kz_ms, 'tflops': kz_flops / kz_ms / 1e9, 'frac_of_bf16_peak': kz_flops / kz_ms / 1e9 / bf16,
                                      'hbm_gbs': (z_bytes + NT * 48 * (NEL // 2) ** 2 * 4) / kz_ms / 1e6,
                                      'frac_of_hbm_peak': (z_bytes + NT * 48 * (NEL // 2) ** 2 * 4) / kz_ms / 1e6 / peaks['hbm_gbs']}
    del xin, z, y3
    torch.cuda.empty_cache()
    HL = NEL // 2
    for name, (cin, cout, k, up) in (('conv1x1 conv_h_kernel', (64, 32, 1, 0)), ('up-conv3x3 conv_h_kernel', (32, 16, 3, 1))):
        xs = [torch.randn((NT, cin, HL, HL), device=dev) for _ in range(3)]
        wh = torch.randn((S_SAMPLES, cout, cin, k, k), device=dev) / (cin * k * k) ** 0.5
        sch = torch.rand((S_SAMPLES, cin), device=dev) + 0.5
        shh = torch.randn((S_SAMPLES, cin), device=dev) * 0.2
        f = 2 if up else 1
        yh = torch.empty((NT, cout, HL * f, HL * f), device=dev)
        nh = int(L.lib.arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, S_SAMPLES))
        scrh = torch.empty((nh,), device=dev)

        def k_h(i):
            L.call('arpde_conv_fwd_h', L.ptr(xs[i % 3]), cin * HL * HL, L.ptr(wh), wh.stride(0), L.ptr(sch), L.ptr(shh), cin, L.ptr(yh), NT, cin, HL, HL, cout, k,
                   up, 1, cout, 0, n_loc, L.ptr(scrh), nh, L.stream())
        h_ms = time_cuda(k_h, 10)
        h_bytes = (xs[0].numel() + yh.numel()) * 4
        kern[name] = {'ms': h_ms, 'hbm_gbs': h_bytes / h_ms / 1e6, 'frac_of_hbm_peak': h_bytes / h_ms / 1e6 / peaks['hbm_gbs'],
                      'tflops': 2.0 * NT * (HL * f) ** 2 * cout * cin * k * k / h_ms / 1e9}
        del xs, yh
        torch.cuda.empty_cache()
    kern['dense block (dense_block_stream_kernel) + output conv (conv_out5x5_kernel), FFMA'] = {
        'ms': step_ms - sum(v['ms'] for v in kern.values()), 'what': 'AR step minus the four launches above'}
    # (2) the stencil-loss kernel (K2), rotating buffers, C4 and C5 sizes, forward / fused-SSE / adjoint
    sten = stencil_rooflines(dev, peaks)
    k2 = sten['c4_64_fwd']

    train = train_weak = ks = c5 = c1 = None
    if not args.no_train_probe:
        train = train_step_probe(dev, rank=rank, world=world)
        if world > 1 and not args.no_extra:
            train_weak = train_step_probe(dev, weak=True, rank=rank, world=world)
    if rank == 0 and not args.no_extra:
        ks = ks_probe(dev)
        c5 = c5_probe(dev, peaks)
        c1 = c1_probe(dev)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        'metric': 'ar_steps_x_samples_per_s', 'value': value, 'unit': 'sample-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks,
        'gpu_launches': launches, 'gpu_launches_source': 'arpde_launch_count(): kernels rank 0 launched through libarpde.so inside the timed region (thread-local counter '
                                                         'bumped after every <<<>>> of the library; torch kernels -- noise draws, history fill -- are not counted)',
        'e2e': {'value': e2e_value, 'unit': 'sample-steps/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': ic_all_host.numel() * 4, 'd2h_bytes_per_step': 2 * N_IC * (T_STEPS + 1) * 2 * NEL * NEL * 4,
                'what': 'pinned host ICs -> device (every rank its shard); rollout in %d pieces: the predictive moments of a finished piece are gathered to rank 0 '
                        '(NCCL) and copied to pinned host memory on a side stream while the next piece runs; rank 0 ends with mean and variance of all 64 ICs' % E2E_CHUNKS},
        'roofline': {'kernel': 'conv_stem_kernel on In_conv1 (6->96, 5x5 circular, %d samples, 30 weight groups; epilogue = eval BatchNorm + ReLU + fp16 two-term operand '
                               'format for In_conv3): tcgen05 implicit GEMM, kind::f16 two-term split, two horizontal taps per K = 16 MMA (+ its weight-split launch)' % NT,
                     'bound': 'tensor', 'achieved': k1_tflops, 'peak': bf16, 'unit': 'TFLOP/s', 'frac': k1_tflops / bf16,
                     'traffic': 3.433e9 if world == 1 else None,
                     'traffic_source': 'ncu --set full, profiles/r02g_step_launches_ncu.md: dram read 0.252 GB + write 3.181 GB per launch at 1920 samples '
                                       '(algorithmic: 0.189 GB input + 3.211 GB operand-format output)',
                     'peak_source': peaks['source'] + ' bf16 dense (burst)',
                     'note': 'achieved = algorithmic fp32 FLOPs (2*N*H*W*cout*cin*k^2) / CUDA-event time; the 1e-5 parity contract needs the two-term fp16 split '
                             '(3 tensor MACs per algorithmic MAC at the fp16 = bf16 rate) and the 6-channel input fills 6 of 8 K lanes with 5 taps in 3 pairs, '
                             'so the attainable ceiling is peak/3 = %.0f TFLOP/s -> frac_of_fp16x3_ceiling %.3f' % (bf16 / 3, k1_tflops / (bf16 / 3)),
                     'flops_per_launch': k1_flops, 'ms': k1_ms,
                     'kernels': kern,
                     'ar_step': {'ms': step_ms, 'tflops': conv_tflops, 'flops': NT * MODEL_MFLOP * 1e6, 'samples': NT,
                                 'what': 'one AR step of the local batch = 6 launches: In_conv1 (conv_stem_kernel) -> In_conv3 (conv_z_kernel) as an fp16 operand-format '
                                         'pair, the four dense layers as ONE FFMA launch (dense_block_stream_kernel), conv1x1 and the folded upsampling conv3x3 on '
                                         'conv_h_kernel (fp16 two-term tensor core), the 16->2 output conv (conv_out5x5_kernel, FFMA): rollout time of the job / %d steps' % T_STEPS},
                     'job_breakdown_ms': breakdown},
        'roofline_stencil': {'kernel': 'burgers2d_fwd_bulk<RB=4, 8 consumer warps> (fused periodic stencil + Crank-Nicolson, API form: read uPred, uPred0, '
                                       'write ustar; persistent CTAs, cp.async.bulk + mbarrier ring of whole-sample tiles), C4 rollout batch, 4 rotating buffer sets',
                             'bound': 'hbm', 'achieved': k2['achieved'], 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': k2['frac'],
                             'traffic': 160.4e6, 'traffic_source': 'ncu --set full, profiles/r02c_stencil_ncu_full.md: dram read 125.9 MB + write 34.5 MB in-kernel '
                                                                   '(algorithmic 125.8 + 62.9 MB; the rest of the written ustar is still in L2 when the kernel ends)', 'bytes_per_launch': k2['bytes_per_launch'], 'ms': k2['ms'], 'peak_source': peaks['source'],
                             'cases': sten},
    }
    if weak is not None:
        line['weak'] = weak
    if train is not None:
        line['train_step'] = train
    if train_weak is not None:
        line['train_step_weak'] = train_weak
    if ks is not None:
        line['c2_ks'] = ks
    if c5 is not None:
        line['c5_wide'] = c5
    if c1 is not None:
        line['c1_burgers1d_train'] = c1
    if not args.no_cpu_baseline:
        S, N, T = REF_SAMPLE
        reference_rate(1, 8, 1)
        rate, dt, kind = reference_rate(S, N, T)
        line['cpu_baseline'] = {'value': rate, 'unit': 'sample-steps/s', 'cores': os.cpu_count(), 'kind': kind,
                                'sample': 'S=%d x %d ICs x T=%d (%.1f s) of the 30x64x100 job; %s' % (
                                    S, N, T, dt, "the reference's own modules driven by its main.testSample() + numpy moments" if kind == 'reference'
                                    else 'oracle port of the reference PyTorch CPU loops')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
