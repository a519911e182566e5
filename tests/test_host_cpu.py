"""CPU: host logic, ABI surface, drop-in structure.  No compute calls (there is no GPU here)."""
import copy
import os
import re
import sys

import pytest
import torch

from _util import ROOT, SYSTEMS, golden, sub
import _models as M

import arpde_b200
from arpde_b200 import _lib as L
from arpde_b200 import program as P


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'arpde.h')).read()
    hdr = re.sub(r'#ifdef ARPDE_DEV.*?#endif', '', hdr, flags=re.S)        # development-build-only hooks are not in the shipped library
    assert 'getenv' not in ''.join(open(os.path.join(ROOT, 'ar-pde-cnn_b200', 'csrc', f)).read() for f in
                                   os.listdir(os.path.join(ROOT, 'ar-pde-cnn_b200', 'csrc')) if f.endswith('.cu')), \
        'the shipped library must not read the environment (use arpde_dev_env)'
    assert not hasattr(L.lib, 'arpde_conv_tc_set_debug')
    declared = set(re.findall(r'^(?:int|int64_t|const char\*)\s+(arpde_\w+)\s*\(', hdr, flags=re.M))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L.lib, name), name
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert L.lib.arpde_version() == 100


@pytest.mark.parametrize('system', SYSTEMS)
def test_state_dict_keys_and_strict_load(system):
    g, cfg, args, model, bayes, swag = M.build(system)
    ref_keys = set(k[len('sd/'):] for k in g if k.startswith('sd/'))
    assert set(bayes.state_dict().keys()) == ref_keys
    M.load_golden_weights(bayes, g)
    ref_swag = set(k[len('swag/sd/'):] for k in g if k.startswith('swag/sd/'))
    assert set(swag.state_dict().keys()) == ref_swag
    swag.load_state_dict(sub(g, 'swag/sd/'), strict=True)
    # parameter order drives SWAG sampling and the optimizer groups
    names = [n for n, _ in bayes.named_parameters()]
    gold_names = [k[len('train/g/'):] for k in g if k.startswith('train/g/')]
    assert names == gold_names and names[0] == 'model.log_beta'


@pytest.mark.parametrize('system', ['burgers1d', 'ks1d'])
def test_shipped_checkpoint_keys_load_strict(system):
    gk = golden(system + '_ckpt200')
    K = int(gk['max_models'])
    g, cfg, args, model, bayes, swag = M.build(system, max_models=K)
    sd = sub(gk, 'sd/')
    if system == 'burgers1d':     # fixture omits the 1.6 MB low-rank rows
        for k, v in swag.state_dict().items():
            if k.endswith('_cov_mat_sqrt'):
                sd[k] = v
    swag.load_state_dict(sd, strict=True)
    assert int(swag.n_models) in (90, 100)


def test_default_init_matches_reference_rng_stream():
    """same torch seed -> same default (Kaiming-uniform) weights as the reference constructor order"""
    g = golden('burgers2d')
    torch.manual_seed(12345)
    m = M.CLS['burgers2d'](6, 2, [4], growth_rate=4, init_features=96)
    w = m.features.In_conv1.weight
    assert torch.allclose(w, torch.from_numpy(g['sd/model.features.In_conv1.weight']))
    assert torch.allclose(m.features.LastTransUp.conv3.weight, torch.from_numpy(g['sd/model.features.LastTransUp.conv3.weight']))


def test_program_of_default_2d_net():
    g, cfg, args, model, bayes, swag = M.build('burgers2d')
    plan = P.compile_module(model.features)
    assert [(o['cin'], o['cout'], o['kw'], o['stride'], o['up']) for o in plan.ops] == [
        (6, 96, 5, 1, 0), (96, 48, 3, 2, 0), (48, 4, 3, 1, 0), (52, 4, 3, 1, 0), (56, 4, 3, 1, 0), (60, 4, 3, 1, 0),
        (64, 32, 1, 1, 0), (32, 16, 3, 1, 1), (16, 2, 5, 1, 0)]
    assert [b['ctot'] for b in plan.bufs] == [96, 64, 32, 16]
    assert [o['out_coff'] for o in plan.ops[2:6]] == [48, 52, 56, 60]
    assert plan.ops[-1]['out_buf'] == -2 and plan.bn_total == 96 + 48 + 52 + 56 + 60 + 64 + 32 + 16
    bufs, work, stats = P.layout(plan, 2, 64, 64)
    # the first buffer (In_conv1 -> In_conv3) is sized for the fp16 operand format of csrc/conv_z.cu: 12 octets x (hi, lo) x 4 parity
    # planes x 33 x 33 positions x 16 bytes = 418176 floats per sample instead of 96 * 4096
    assert bufs[1][2:] == (64, 32, 32) and work == 2 * (12 * 2 * 4 * 33 * 33 * 4 + 64 * 1024 + 32 * 1024 + 16 * 4096)
    # plan survives deepcopy and points at the copied modules (SwagNN.sample deep-copies the model)
    m2 = copy.deepcopy(model)
    p2 = P.compile_module(m2.features)
    assert p2.param_slots[0][0] is m2.features.In_conv1


def test_program_generic_topology():
    m = M.CLS['burgers2d'](1, 1, (2, 3, 2), growth_rate=16, init_features=48, out_activation='Tanh')
    plan = P.compile_module(m.features)
    ks = [(o['kw'], o['stride'], o['up']) for o in plan.ops]
    assert (4, 2, 0) in ks and (3, 1, 1) in ks and plan.ops[-1]['act'] == 1
    assert plan.out_scale == 0 and plan.c_in == 1 and plan.c_out == 1
    with pytest.raises(ValueError):
        M.CLS['burgers2d'](1, 1, (2, 2), growth_rate=16, init_features=48)


def test_no_cpu_fallback():
    g, cfg, args, model, bayes, swag = M.build('ks1d')
    x = torch.zeros(2, 2, 96)
    with pytest.raises(RuntimeError, match='CUDA'):
        model(x)
    integ = M.integrator('ks1d', g, 'cpu')
    with pytest.raises(RuntimeError, match='CUDA'):
        integ.crankNicolson(torch.zeros(2, 1, 96), torch.zeros(2, 1, 96), 0.1)
    with pytest.raises(RuntimeError, match='CUDA'):
        bayes.calc_neg_log_joint(torch.zeros(2, 1, 96), torch.zeros(2, 1, 96), 3)


def test_product_never_imports_oracle():
    pkg = arpde_b200.PACKAGE_DIR
    n = 0
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.sh')):
                n += 1
                for line in open(os.path.join(root, f)):
                    assert not ('oracle' in line and ('import' in line or 'include' in line)), (f, line)
    assert n > 10


def test_dropin_module_names():
    import importlib
    import subprocess
    code = ("import sys; sys.path.insert(0, %r); "
            "from nn.denseEDcirc2d import DenseED; from nn.bayesNN import BayesNN; from nn.swag import SwagNN; "
            "from nn.burger2DFiniteDifference import Burger2DIntegrate; print(DenseED.__name__)")
    out = subprocess.check_output([sys.executable, '-c', code % os.path.join(arpde_b200.PACKAGE_DIR, 'dropin', '2D-Burgers-SWAG')])
    assert b'DenseED2d' in out
    for proj, mods in (('1D-Burger-SWAG', 'from nn.denseEDcirc import DenseED; from nn.burgerFiniteDifference import BurgerIntegrate'),
                       ('1D-KS-SWAG', 'from nn.denseEDcirc import DenseED; from nn.ksFiniteDifference import KSIntegrate')):
        code = "import sys; sys.path.insert(0, %r); %s; from nn.swag import SwagNN; from nn.bayesNN import BayesNN"
        subprocess.check_call([sys.executable, '-c', code % (os.path.join(arpde_b200.PACKAGE_DIR, 'dropin', proj), mods)])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU): one JSON line with the keys the driver reads; the reference's own modules when they are staged."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['metric'] == 'ar_steps_x_samples_per_s' and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and 'workload' in d['config']
    from oracle import ref_shims
    if ref_shims.ref_root() is not None:
        assert d['cpu_baseline']['kind'] == 'reference'


def test_committed_bench_records_carry_the_contract_keys():
    """The bench records under profiles/ are what DESIGN.md quotes: the latest one-GPU record and the latest 8-GPU record must hold every
    key of the measurement contract (metric line, end-to-end number, roofline of the dominant kernel and of the stencil, CPU baseline,
    clocks, launch counter), with fractions that follow from their own achieved / peak numbers."""
    import glob
    import json
    recs = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02*_bench_1gpu.json')))
    assert recs, 'no one-GPU bench record committed'
    d = json.loads(open(recs[-1]).read().strip().splitlines()[-1])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data',
              'config', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'roofline_stencil', 'cpu_baseline'):
        assert k in d, k
    assert d['metric'] == 'ar_steps_x_samples_per_s' and d['n_gpus'] == 1 and d['config']['workload'].startswith('C4')
    assert d['gpu_launches'] > 0 and d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    assert 0 < d['e2e']['value'] <= d['value'] * 1.001
    for r in (d['roofline'], d['roofline_stencil']):
        assert r['bound'] in ('hbm', 'tensor') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert d['roofline_stencil']['frac'] >= 0.6                    # the north star's stencil target, on the API form at the C4 batch
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
    recs8 = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r02*_bench_8gpu.json')))
    if recs8:
        d8 = json.loads(open(recs8[-1]).read().strip().splitlines()[-1])
        assert d8['n_gpus'] == 8 and d8['scaling'] == 'strong' and d8['config']['ics_per_gpu'] == 8
