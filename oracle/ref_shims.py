"""
oracle/ref_shims.py -- TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

How the UNMODIFIED reference (cics-nd/ar-pde-cnn) is run in this repository:

  * where it lives: `/root/reference` exists only in the build container.  `stage_reference()` (called by
    `__graft_entry__.build()`) copies the three project folders -- Python sources and the shipped 1D checkpoints,
    nothing else -- into `baseline/_ref/` (git-ignored, so no reference source enters the history; NOT gpurun-ignored,
    so it travels to the GPU box like a built `.so`).  `ref_root()` finds whichever copy exists.
  * shims (none edits a reference file):
      1. torch<=1.2 circular padding: every Conv{1,2}d built with padding_mode='circular' gets padding //= 2
         (reference README.md:40 prescribes exactly this for newer torch);
      2. matplotlib is not installed: MagicMock modules are placed in sys.modules;
      3. `main.py` needs validation files from FEniCS / MATLAB solvers that cannot run here: `synth_validation_data`
         writes smooth synthetic stand-ins with the file names and shapes the loaders read
         (2D-Burgers-SWAG/utils/burgerLoader2D.py:37-77, 1D-Burger-SWAG/utils/burgerLoader.py:86-116,
          1D-KS-SWAG/utils/ksLoader.py:87-114).
  * `run_main(project, nn, workdir, argv)` executes a project's `main.py` with runpy, either on the reference's own `nn/`
    package (nn='ref') or on this repo's drop-in package (nn='dropin', ar-pde-cnn_b200/dropin/<project>/nn shadows it).

Users: tests/golden/make_golden.py, tests/ (T13: main.py end to end), bench.py --impl reference / cpu_baseline.
The product package never imports this module.
"""
import os
import shutil
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference'
REF_STAGED = os.path.join(ROOT, 'baseline', '_ref')
PROJ = {'burgers2d': '2D-Burgers-SWAG', 'burgers1d': '1D-Burger-SWAG', 'ks1d': '1D-KS-SWAG'}


def ref_root():
    """Directory holding the reference project folders: the staged copy if present, else the container's /root/reference."""
    for d in (REF_STAGED, REF_SRC):
        if os.path.isdir(os.path.join(d, PROJ['burgers2d'], 'nn')):
            return d
    return None


def stage_reference(force=False):
    """Copy the reference's Python sources + shipped checkpoints to baseline/_ref (build container only)."""
    if not os.path.isdir(REF_SRC):
        return ref_root()
    marker = os.path.join(REF_STAGED, '.staged')
    if os.path.exists(marker) and not force:
        return REF_STAGED
    for proj in PROJ.values():
        for dp, dn, fn in os.walk(os.path.join(REF_SRC, proj)):
            rel = os.path.relpath(dp, REF_SRC)
            for f in fn:
                if f.endswith('.py') or f.endswith('.pth'):
                    os.makedirs(os.path.join(REF_STAGED, rel), exist_ok=True)
                    shutil.copy2(os.path.join(dp, f), os.path.join(REF_STAGED, rel, f))
    for f in ('README.md', 'LICENSE'):
        if os.path.exists(os.path.join(REF_SRC, f)):
            shutil.copy2(os.path.join(REF_SRC, f), os.path.join(REF_STAGED, f))
    open(marker, 'w').write('copied from %s by oracle/ref_shims.py:stage_reference (never committed)\n' % REF_SRC)
    return REF_STAGED


def mock_matplotlib():
    for name in ['matplotlib', 'matplotlib.pyplot', 'matplotlib.cm', 'matplotlib.ticker', 'matplotlib.lines', 'matplotlib.colorbar',
                 'matplotlib.gridspec', 'matplotlib.patches', 'matplotlib.colors', 'mpl_toolkits', 'mpl_toolkits.axes_grid1']:
        sys.modules.setdefault(name, mock.MagicMock())


_PAD_PATCHED = False


def halve_circular_padding():
    """Shim 1: the reference passes padding=2p, which torch<=1.2 applied as p per side."""
    global _PAD_PATCHED
    if _PAD_PATCHED:
        return
    import torch.nn as nn
    for cls in (nn.Conv1d, nn.Conv2d):
        orig = cls.__init__

        def patched(self, *a, __orig=orig, **kw):
            __orig(self, *a, **kw)
            if self.padding_mode == 'circular':
                self.padding = tuple(p // 2 for p in self.padding)
                self._reversed_padding_repeated_twice = tuple(x for p in reversed(self.padding) for x in (p, p))
        cls.__init__ = patched
    _PAD_PATCHED = True


import contextlib


@contextlib.contextmanager
def reference_modules(system):
    """`with reference_modules('burgers2d') as proj:` -- inside the block `import nn.xxx` / `import main` resolve to the reference
    project and Conv layers are built with the halved circular padding; on exit torch.nn and sys.path / sys.modules are restored, so
    the product's own modules can be built in the same process (pytest).  Reference modules built inside keep working afterwards."""
    global _PAD_PATCHED
    import torch.nn as nn
    root = ref_root()
    if root is None:
        raise FileNotFoundError('reference not available')
    saved_init = (nn.Conv1d.__init__, nn.Conv2d.__init__)
    saved_path = list(sys.path)
    was = _PAD_PATCHED
    clash = lambda k: k == 'nn' or k.startswith('nn.') or k == 'utils' or k.startswith('utils.') or k in ('args', 'main')
    saved_mods = {k: sys.modules.pop(k) for k in list(sys.modules) if clash(k)}
    mock_matplotlib()
    _PAD_PATCHED = False
    halve_circular_padding()
    proj = os.path.join(root, PROJ[system])
    sys.path.insert(0, proj)
    try:
        yield proj
    finally:
        nn.Conv1d.__init__, nn.Conv2d.__init__ = saved_init
        _PAD_PATCHED = was
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if clash(k)]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def install_shims():
    """Shims for importing the reference's own nn/ modules."""
    mock_matplotlib()
    halve_circular_padding()


def import_reference(system):
    """sys.path set up for one reference project + shims installed; returns the project directory."""
    root = ref_root()
    if root is None:
        raise FileNotFoundError('the reference is neither at %s nor staged at %s' % (REF_SRC, REF_STAGED))
    install_shims()
    proj = os.path.join(root, PROJ[system])
    for k in [k for k in sys.modules if k == 'nn' or k.startswith('nn.') or k == 'utils' or k.startswith('utils.') or k in ('args', 'main')]:
        del sys.modules[k]
    sys.path.insert(0, proj)
    return proj


def synth_validation_data(system, data_dir, nel=None):
    """Smooth synthetic stand-ins for the solver outputs main.py loads as test cases (values are irrelevant to training)."""
    import numpy as np
    rng = np.random.RandomState(4242)
    os.makedirs(data_dir, exist_ok=True)
    if system == 'burgers2d':
        n = nel or 64
        x = np.linspace(0, 1, n + 1)[:-1]
        xx, yy = np.meshgrid(x, x)
        for case in (400, 401):
            d = os.path.join(data_dir, 'run%d' % case)
            os.makedirs(d, exist_ok=True)
            a, b = rng.randn(2), rng.rand(2)
            for j in range(0, 201, 2):
                damp = np.exp(-0.002 * j)
                u = np.stack([a[0] * np.sin(2 * np.pi * (xx + b[0])) * np.cos(2 * np.pi * yy), a[1] * np.cos(2 * np.pi * (yy + b[1])) * np.sin(2 * np.pi * xx)]) * damp
                np.save(os.path.join(d, 'u%d.npy' % j), u.astype(np.float64))
    elif system == 'burgers1d':
        n = nel or 512
        x = np.linspace(0, 1, n + 1)
        t = np.arange(2001)[:, None]
        for case in range(495, 500):
            a, b = rng.randn(), rng.rand()
            u = a * np.sin(2 * np.pi * (x[None, :] + b)) * np.exp(-0.0005 * t)
            np.save(os.path.join(data_dir, 'u%d.npy' % case), u.astype(np.float64))
    else:
        n = nel or 96
        x = np.linspace(0, 22 * np.pi, n + 1)
        t = np.arange(2010)[:, None]
        for case in range(1, 6):
            a, b = 2.5 + rng.rand(), rng.rand() * 6
            u = a * np.sin(2 * np.pi * 3 * x[None, :] / (22 * np.pi) + b + 0.01 * t)
            np.savetxt(os.path.join(data_dir, 'ks_data_%d.dat' % case), u, delimiter=',', fmt='%.6e')


def run_main(system, nn, workdir, argv, seed=12345, stub_plots=True):
    """Execute the reference's UNMODIFIED <project>/main.py (runpy, __main__) inside `workdir`.

    nn='ref'   : the reference's own nn/ package (padding shim installed);
    nn='dropin': this repo's ar-pde-cnn_b200/dropin/<project>/nn package shadows it (no padding shim: the drop-in
                 convolutions already hold the effective padding).
    The project's Python files are copied to workdir (main.py writes ./experiments next to itself), synthetic validation
    files are written under workdir/solver/validation_data, numpy is seeded (args.py seeds only random and torch).
    stub_plots: utils.post.plotPred / plotSamples (matplotlib) are replaced by a function that prints a checksum of the
    prediction it was handed, so the test()/testSample() rollouts of main.py are exercised and comparable."""
    import runpy
    import numpy as np
    import torch
    root = ref_root()
    if root is None:
        raise FileNotFoundError('reference not available')
    proj_src = os.path.join(root, PROJ[system])
    proj = os.path.join(workdir, PROJ[system])
    if os.path.isdir(proj):
        shutil.rmtree(proj)
    shutil.copytree(proj_src, proj, ignore=shutil.ignore_patterns('*.pth', '__pycache__', 'experiments'))
    argv = list(argv)
    nel = int(argv[argv.index('--nel') + 1]) if '--nel' in argv else None
    synth_validation_data(system, os.path.join(proj, 'solver', 'validation_data'), nel)
    mock_matplotlib()
    if nn == 'ref':
        halve_circular_padding()
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    elif nn == 'dropin':
        assert not _PAD_PATCHED, 'the padding shim must not be active when the drop-in modules are used'
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'ar-pde-cnn_b200', 'dropin', PROJ[system]))
    else:
        raise ValueError(nn)
    sys.path.insert(1 if nn == 'dropin' else 0, proj)
    if nn == 'dropin':
        # the drop-in nn/ must win over <project>/nn: it is first on sys.path, the project directory second
        assert sys.path[0].endswith(os.path.join('dropin', PROJ[system]))
    os.chdir(proj)
    np.random.seed(seed)
    if stub_plots:
        import importlib
        post = importlib.import_module('utils.post')

        def _stub(name):
            def f(*a, **k):
                # first array with a time axis = the prediction; the first few saved steps only (an untrained net's long
                # rollout blows up and is not comparable between implementations)
                arrs = [x for x in a if isinstance(x, np.ndarray) and x.ndim >= 2]
                if arrs:
                    arr = arrs[0]
                    head = arr[:, :4] if (arr.ndim == 3 and arr.shape[0] <= 8) else arr[:4]      # [cases, time, x] vs [time, ...]
                    print('T13 %s pred_absmean_first4 %.6e shape %s' % (name, float(np.abs(head).mean()), tuple(arr.shape)))
            return f
        post.plotPred, post.plotSamples = _stub('plotPred'), _stub('plotSamples')
    old_argv = sys.argv
    sys.argv = ['main.py'] + list(argv)
    try:
        runpy.run_path(os.path.join(proj, 'main.py'), run_name='__main__')
    finally:
        sys.argv = old_argv


if __name__ == '__main__':
    # python oracle/ref_shims.py <system> <ref|dropin> <workdir> [main.py args...]
    run_main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:])
