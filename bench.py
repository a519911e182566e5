#!/usr/bin/env python
"""bench.py -- headline benchmark of the ar-pde-cnn hot path on B200.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): 2D coupled Burgers' SWAG
Bayesian inference, 30 posterior weight samples x 64 random-Fourier initial conditions x 100 auto-regressive
steps on a 64x64 periodic grid, default DenseED (blocks=[4], growth 4, init_features 96).  One "step" of this
benchmark = one complete job: draw the 30 weight samples, roll all 1920 (sample, IC) pairs out for 100 steps,
reduce the predictive mean/variance.  metric = AR steps x samples / s = S*N_ic*T / time.

    python bench.py --gpus N --steps K --warmup W          # our CUDA path (N>1: launched under torchrun)
    python bench.py --impl reference ...                   # the reference's CPU (PyTorch) path, oracle port

Multi-GPU: initial conditions are sharded across ranks (every rank runs all weight samples for its own 64 ICs;
weak scaling, no data-path collective) followed by one NCCL all-gather of the predictive mean and variance.
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
MODEL_MFLOP = 267.32          # forward MFLOP per sample-step of the default 2D net
STENCIL_BYTES_PT = 24         # K2 API form: read uPred + uPred0, write ustar, 2 comps x 4 B


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


def make_ics(n, first=0):
    """InitPeriodicCond2d of the reference (utils/burgerLoader2D.py:80-116), seeded per index."""
    import numpy as np
    import torch
    order = 4
    x = np.linspace(0, 1, NEL + 1)[:-1]
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
        out.append(torch.FloatTensor(f).reshape(-1, NEL, NEL))
    return torch.stack(out, 0)


def build_swag(device, seed=12345):
    """Default 2D model with torch.manual_seed init (no 2D checkpoint is shipped: reference .MISSING_LARGE_BLOBS)
    and a SWAG posterior built by 30 collect() calls on perturbed weights."""
    import types
    import numpy as np
    import torch
    from arpde_b200.nn.densed import DenseED2d
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    import contextlib, io
    torch.manual_seed(seed); np.random.seed(seed)
    args = types.SimpleNamespace(dt=0.005, device=torch.device(device), nic=3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(device)
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
    return swag


def cpu_reference_rate(S, N, T, threads=None):
    """The reference's CPU path (PyTorch ops, oracle port of the same loops) on a bounded sample of the workload."""
    import torch
    from oracle import arpde_oracle as O
    import types, contextlib, io
    import numpy as np
    torch.set_num_threads(threads or os.cpu_count())
    torch.manual_seed(12345)
    # weights: default init of the same architecture, built with plain torch (no GPU needed)
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
    ic = make_ics(N)
    t0 = time.perf_counter()
    with torch.no_grad():
        for s in range(S):        # one weight sample at a time, as testSample() does
            step = lambda inp: O.densed_forward('burgers2d', sd, inp, [4], training=False)
            O.rollout('burgers2d', step, ic, 3, T)
    dt = time.perf_counter() - t0
    return S * N * T / dt, dt


def run_reference(args):
    """--impl reference: rank 0 times the CPU path; other ranks exit."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import torch
    S, N, T = 2, 64, 4
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_reference_rate(1, 8, 1)
    t0 = time.perf_counter()
    tot = 0.0
    for _ in range(args.steps):
        rate, dt = cpu_reference_rate(S, N, T)
        tot += dt
    value = args.steps * S * N * T / tot
    sample = 'S=%d weight samples x %d ICs x T=%d steps per step (of 30x64x100), oracle port of the reference PyTorch CPU path' % (S, N, T)
    line = {'impl': 'reference', 'metric': 'ar_steps_x_samples_per_s', 'value': value, 'unit': 'sample-steps/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': 'sample-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': 'sample-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def train_step_probe(dev, B=128, tback=2, nwarm=3, nrep=10, rank=0, world=1):
    """Training-step timing (SURVEY.md 8 row M5 + P1 + L1; BASELINE config C2): every rank takes its own mini-batch of B random Fourier
    ICs (weak scaling), runs a tback-step BPTT window (DenseED forward with batch-statistics BatchNorm, Crank-Nicolson target, negative
    log joint), loss.backward(), ONE flat-bucket gradient all-reduce over NCCL when world > 1, and the Adam step
    (2D-Burgers-SWAG/main.py:61-91).  Reported next to the headline; not part of the timed rollout."""
    import contextlib, io, types
    import numpy as np
    import torch
    from arpde_b200.nn.densed import DenseED2d
    from arpde_b200.nn.bayesnn import BayesNN
    from arpde_b200.nn.swag import SwagNN
    from arpde_b200.nn.integrators import Burger2DIntegrate
    torch.manual_seed(0); np.random.seed(0)
    a = types.SimpleNamespace(dt=0.005, device=dev, nic=3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = DenseED2d(6, 2, [4], growth_rate=4, init_features=96).to(dev)
        bayes = BayesNN(a, model); swag = SwagNN(a, bayes, full_cov=True, max_models=30)
        integ = Burger2DIntegrate(1.0 / NEL, nu=0.005, grad_kernels=[3, 3], device=dev)
    from arpde_b200.optim import Adam
    opt = Adam([{'params': [bayes.model.log_beta], 'lr': 1e-3}, {'params': bayes.model.features.parameters()}], lr=1e-3)
    from arpde_b200.data import InitPeriodicCond2d
    ic = InitPeriodicCond2d(NEL, 1 << 30, device=dev).batch(range(rank * B, rank * B + B))      # synthesised on the device (csrc/ic.cu)
    swag.train()
    bucket = None
    if world > 1:
        import torch.distributed as dist
        from arpde_b200.parallel import FlatGradBucket
        bucket = FlatGradBucket([bayes.model.log_beta] + list(bayes.model.features.parameters()))

    def step():
        inp = ic.repeat(1, 3, 1, 1)
        loss = 0
        for _ in range(tback):
            u_pred = swag(inp[:, -6:])
            ustar = integ.crankNicolson(u_pred, inp[:, -2:], 0.003)
            loss = loss + swag.calc_neg_log_joint(u_pred, ustar, 10)
            inp = torch.cat([inp, u_pred], 1)
        loss.backward()
        if bucket is not None:
            bucket.all_reduce_(average=True)
        opt.step(); opt.zero_grad()
    for _ in range(nwarm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nrep
    if world > 1:                                   # slowest rank
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {'ms': ms, 'value': world * tback * B / ms * 1e3, 'unit': 'sample-steps/s', 'batch_per_gpu': B, 'n_gpus': world, 'tback': tback,
            'scaling': 'weak', 'parallelism': 'mini-batch sharded x%d, one flat-bucket gradient all-reduce (NCCL) per optimizer step' % world,
            'what': 'forward (batch-stat BN) + Crank-Nicolson + neg log joint + backward (tensor-core data and weight gradients) + one-launch Adam, '
                    'mean of %d steps after %d warm-up (CUDA events, max over ranks)' % (nrep, nwarm)}


def workload_config(n_gpus):
    return {'workload': 'C4: 2D coupled Burgers SWAG inference, %d weight samples x %d ICs/GPU x %d AR steps, %dx%d grid, DenseED blocks=[4] growth=4 init_features=96'
                        % (S_SAMPLES, N_IC, T_STEPS, NEL, NEL),
            'samples': S_SAMPLES, 'ics_per_gpu': N_IC, 'ar_steps': T_STEPS, 'grid': [NEL, NEL],
            'parallelism': 'ic-shard x%d + final all-gather of mean/var' % n_gpus,
            'l2': 'inputs larger than L2 (per-step activations 4.3 GB, stencil batch 189 MB)'}


def main():
    global T_STEPS
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-probe', action='store_true', help='skip the (untimed, rank-0) training-step measurement')
    ap.add_argument('--tsteps', type=int, default=T_STEPS, help='debug: shorter rollouts (the JSON line then names the reduced config)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import arpde_b200
    from arpde_b200 import engine, functional as F_
    T_STEPS = args.tsteps

    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'        # keep stdout to the single JSON line (the banner goes to stdout)
        dist.init_process_group('nccl', device_id=dev)
    peaks = load_peaks()

    swag = build_swag(dev)
    ic_host = make_ics(N_IC, first=rank * N_IC).pin_memory()        # distinct ICs per rank
    ic_dev = ic_host.to(dev)
    named = list(swag.base.named_parameters())
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + 0)     # every rank draws the SAME weight samples
    res_shape = (N_IC, T_STEPS + 1, 2, NEL, NEL)
    mean_host = torch.empty(res_shape, pin_memory=True)
    var_host = torch.empty(res_shape, pin_memory=True)
    gather_m = torch.empty((world,) + res_shape, device=dev) if world > 1 else None
    gather_v = torch.empty((world,) + res_shape, device=dev) if world > 1 else None
    NT = S_SAMPLES * N_IC
    c_hist = 6
    traj_buf = torch.empty((NT, c_hist + 2 * T_STEPS, NEL, NEL), device=dev)

    def job(ic):
        """device-resident job: sample weights, rollout, predictive moments (+ gather across ranks)"""
        flat, names, numels = swag.sample_flat(S_SAMPLES, generator=gen)
        traj = engine.rollout(swag, ic, T_STEPS, flat_weights=flat, named_params=named, out=traj_buf)
        betas = flat[:, 0].exp()
        mean, var = engine.predictive_moments(traj, S_SAMPLES, betas, '2d')
        if world > 1:
            dist.all_gather_into_tensor(gather_m, mean)
            dist.all_gather_into_tensor(gather_v, var)
        return mean, var

    def job_e2e():
        """the call a user makes: host ICs in, host predictive mean/variance out"""
        ic = ic_host.to(dev, non_blocking=True)
        mean, var = job(ic)
        mean_host.copy_(mean, non_blocking=True)
        var_host.copy_(var, non_blocking=True)
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

    for _ in range(args.warmup):
        job(ic_dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: job(ic_dev), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    units = world * S_SAMPLES * N_IC * T_STEPS * args.steps
    value = units / (ms * 1e-3)

    for _ in range(max(1, args.warmup // 3)):
        job_e2e()
    ms_e2e = timed(job_e2e, args.steps)
    e2e_value = units / (ms_e2e * 1e-3)

    # ---- per-kernel rooflines, measured live with CUDA events on the launching stream ------------------
    # (1) one full AR step (9 fused convs: In_conv1, In_conv3, conv1x1 and the folded conv3x3-up on the tensor-core kernel, the four
    #     4-channel dense layers and the 16->2 output conv on the row-tiled FFMA kernel): difference of a 25-step and a 5-step rollout,
    #     so that the per-rollout setup (weight split, buffers) drops out
    flat, _, _ = swag.sample_flat(S_SAMPLES, generator=gen)
    engine.rollout(swag, ic_dev, 1, flat_weights=flat, named_params=named)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    engine.rollout(swag, ic_dev, 5, flat_weights=flat, named_params=named)
    e1.record()
    engine.rollout(swag, ic_dev, 25, flat_weights=flat, named_params=named)
    e2.record(); torch.cuda.synchronize()
    step_ms = (e1.elapsed_time(e2) - e0.elapsed_time(e1)) / 20
    conv_tflops = NT * MODEL_MFLOP * 1e6 / (step_ms * 1e-3) / 1e12
    # (1b) the dominant launch of the step, timed alone through the C ABI on the rollout batch:
    #      conv_tc_kernel on In_conv3 (96 -> 48, 3x3 stride 2, 64x64 -> 32x32), 1920 samples, 30 weight groups
    import arpde_b200._lib as L
    xin = torch.randn((NT, 96, NEL, NEL), device=dev)
    w3 = torch.randn((S_SAMPLES, 48, 96, 3, 3), device=dev) / (96 * 9) ** 0.5
    y3 = torch.empty((NT, 48, NEL // 2, NEL // 2), device=dev)
    n_scr = int(L.lib.arpde_conv_fwd_tc_scratch_floats(96, 48, 3, 3, S_SAMPLES))
    scr = torch.empty((n_scr,), device=dev)
    def k1():
        L.call('arpde_conv_fwd_tc', L.ptr(xin), 96 * NEL * NEL, L.ptr(w3), 48 * 96 * 9, 0, 0, 0, L.ptr(y3), 0, NT, 96, NEL, NEL, 48, 3, 3, 2, 0, 0, 0,
               48, 0, N_IC, L.ptr(scr), n_scr, L.stream())
    for _ in range(3):
        k1()
    torch.cuda.synchronize()
    nrep = 10
    e0.record()
    for _ in range(nrep):
        k1()
    e1.record(); torch.cuda.synchronize()
    k1_ms = e0.elapsed_time(e1) / nrep
    k1_flops = 2.0 * NT * (NEL // 2) ** 2 * 48 * 96 * 9
    k1_tflops = k1_flops / (k1_ms * 1e-3) / 1e12
    del xin, y3
    # (2) the stencil-loss kernel (K2) on the rollout batch: 1920 states x 64x64 x 24 B = 189 MB > L2
    up = traj_buf[:, 6:8]
    u0 = traj_buf[:, 4:6]
    ustar = torch.empty((NT, 2, NEL, NEL), device=dev)
    bs = traj_buf.stride(0)
    def k2():
        L.call('arpde_cn_burgers2d_fwd', L.ptr(up), bs, L.ptr(u0), bs, L.ptr(ustar), 0, NT, NEL, NEL, 1.0 / NEL, 0.005, 0.005, 0, L.stream())
    for _ in range(3):
        k2()
    torch.cuda.synchronize()
    nrep = 20
    e0.record()
    for _ in range(nrep):
        k2()
    e1.record(); torch.cuda.synchronize()
    k2_ms = e0.elapsed_time(e1) / nrep
    k2_bytes = NT * NEL * NEL * STENCIL_BYTES_PT
    k2_gbs = k2_bytes / (k2_ms * 1e-3) / 1e9

    train = None
    if not args.no_train_probe:
        train = train_step_probe(dev, rank=rank, world=world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    n_ops = 9
    launches = args.steps * (T_STEPS * n_ops + 4 + 1 + 1 + 1)        # convs + 4 weight-split launches (tensor-core layers) + bn fold + swag sample + moments
    line = {
        'metric': 'ar_steps_x_samples_per_s', 'value': value, 'unit': 'sample-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(world), 'clocks': clocks, 'gpu_launches': launches,
        'e2e': {'value': e2e_value, 'unit': 'sample-steps/s', 'ms_per_step': ms_e2e / args.steps,
                'h2d_bytes_per_step': ic_host.numel() * 4, 'd2h_bytes_per_step': 2 * mean_host.numel() * 4},
        'roofline': {'kernel': 'conv_tc_kernel on In_conv3 (96->48, 3x3/s2, 1920 samples): tcgen05 implicit GEMM, 3xTF32 split (+ its weight-split launch)',
                     'bound': 'tensor', 'achieved': k1_tflops, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': k1_tflops / peaks['bf16_tflops'],
                     'traffic': 4.486e9, 'traffic_source': 'ncu --set full, profiles/r01d_conv_tc_in3_ncu_full.md: dram read 4.108 GB + write 0.377 GB per launch '
                                                           '(algorithmic input + output 3.40 GB; the excess is halo rows re-read past L2)',
                     'peak_source': peaks['source'] + ' bf16 dense (burst)',
                     'note': 'achieved = algorithmic fp32 FLOPs (2*N*Ho*Wo*cout*cin*k^2) / CUDA-event time; the 1e-5 parity contract needs the 3-term '
                             'TF32 split (3 tensor MACs per algorithmic MAC) and TF32 runs at half the bf16 rate, so the attainable ceiling is '
                             'peak/6 = %.0f TFLOP/s -> frac_of_3xtf32_ceiling %.3f' % (peaks['bf16_tflops'] / 6, k1_tflops / (peaks['bf16_tflops'] / 6)),
                     'flops_per_launch': k1_flops, 'ms': k1_ms,
                     'ar_step': {'ms': step_ms, 'tflops': conv_tflops, 'flops': NT * MODEL_MFLOP * 1e6,
                                 'what': 'all 9 convs of one AR step (4 on conv_tc_kernel, the 4-channel dense layers and the 16->2 output conv on the row-tiled fp32 FFMA kernel)'}},
        'roofline_stencil': {'kernel': 'burgers2d_fwd_vec4<RY=4> (fused periodic stencil + Crank-Nicolson, API form: read uPred, uPred0, write ustar)', 'bound': 'hbm',
                             'achieved': k2_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': k2_gbs / peaks['hbm_gbs'],
                             'traffic': 1.601e8, 'traffic_source': 'ncu --set full, profiles/r01b_stencil_ncu_full.md: dram read 125.9 MB + write 34.2 MB '
                                                                   '(the rest of the 62.9 MB of ustar was still in L2 at kernel end)',
                             'bytes_per_launch': k2_bytes, 'ms': k2_ms, 'peak_source': peaks['source']},
    }
    if train is not None:
        line['train_step'] = train
    if not args.no_cpu_baseline:
        S, N, T = 2, 64, 4
        rate, dt = cpu_reference_rate(S, N, T)
        line['cpu_baseline'] = {'value': rate, 'unit': 'sample-steps/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': 'S=%d x %d ICs x T=%d (%.1f s) of the 30x64x100 job, oracle port of the reference PyTorch CPU loops' % (S, N, T, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
