"""
Generate tests/golden/t13_main.json: the UNMODIFIED reference drivers (2D-Burgers-SWAG/main.py, 1D-Burger-SWAG/main.py,
1D-KS-SWAG/main.py) run end to end on the reference's own nn/ modules, CPU, for two tiny epochs.

    python tests/golden/make_t13.py          # build container only (needs /root/reference or baseline/_ref)

The harness is oracle/ref_shims.py:run_main (shims: halved circular padding, mocked matplotlib, synthetic validation
files, plot functions replaced by a checksum print).  Recorded per system: the command line, the "Epoch: n, Loss, MSE,
Noise" lines main.py prints, and the checksums of the first saved steps of the test() rollout.  tests/test_gpu_main_dropin.py
re-runs the same command with nn = this repo's drop-in package on the GPU and compares.
"""
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

ARGS = {
    'burgers2d': ['--epochs', '2', '--ntrain', '16', '--batch-size', '8', '--plot-freq', '2', '--ckpt-freq', '2', '--nel', '32'],
    # 1D Burgers back-propagates through windows of up to 9 AR steps: with the default learning rate two epochs on an untrained net are
    # already chaotic at the 1e-2 level even between two runs of the reference itself (cuDNN vs MKL: loss 5 %, predictions 7 %), so this
    # driver is exercised with a smaller step size
    'burgers1d': ['--epochs', '2', '--ntrain', '16', '--batch-size', '8', '--plot-freq', '2', '--ckpt-freq', '2', '--nel', '64',
                  '--lr', '0.0001', '--lr-beta', '0.0001'],
    'ks1d': ['--epochs', '2', '--ntrain', '16', '--batch-size', '8', '--plot-freq', '2', '--ckpt-freq', '2'],
}
EPOCH_RE = re.compile(r'^Epoch: (\d+), Loss: ([0-9.E+-]+), MSE: ([0-9.E+-]+), Noise ([0-9.]+)')
PRED_RE = re.compile(r'^T13 (\w+) pred_absmean_first4 ([0-9.e+-]+) shape (.*)$')


def parse(stdout):
    epochs, preds = [], []
    for line in stdout.splitlines():
        m = EPOCH_RE.match(line)
        if m:
            epochs.append(dict(epoch=int(m.group(1)), loss=float(m.group(2)), mse=float(m.group(3)), noise=float(m.group(4))))
        m = PRED_RE.match(line)
        if m:
            preds.append(dict(what=m.group(1), absmean_first4=float(m.group(2)), shape=m.group(3)))
    return dict(epochs=epochs, preds=preds)


def run(system, nn, env=None):
    with tempfile.TemporaryDirectory() as wd:
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'ref_shims.py'), system, nn, wd] + ARGS[system],
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=1800)
    if out.returncode != 0:
        raise RuntimeError('main.py (%s, nn=%s) failed:\n%s' % (system, nn, out.stdout[-4000:]))
    return parse(out.stdout)


if __name__ == '__main__':
    res = {}
    for system in ARGS:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
        r = run(system, 'ref', env)
        r['argv'] = ARGS[system]
        res[system] = r
        print(system, r['epochs'], r['preds'][:2])
    json.dump(res, open(os.path.join(HERE, 't13_main.json'), 'w'), indent=1)
