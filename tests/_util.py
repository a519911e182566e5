"""Shared helpers for the test-suite: golden fixture access and the oracle import."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
SYSTEMS = ('burgers2d', 'burgers1d', 'ks1d')

_cache = {}


def golden(name):
    if name not in _cache:
        with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
            _cache[name] = {k: z[k] for k in z.files}
    return _cache[name]


def sub(d, prefix, strip=''):
    """Tensors of fixture dict d under prefix, key prefix `strip` removed after it."""
    out = {}
    for k, v in d.items():
        if k.startswith(prefix):
            kk = k[len(prefix):]
            if strip and kk.startswith(strip):
                kk = kk[len(strip):]
            out[kk] = torch.from_numpy(np.array(v))
    return out


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.detach().double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def cfg_of(g):
    return dict(blocks=[int(v) for v in g['cfg/blocks']], growth_rate=int(g['cfg/growth_rate']),
                init_features=int(g['cfg/init_features']), in_channels=int(g['cfg/in_channels']),
                out_channels=int(g['cfg/out_channels']), nic=int(g['cfg/nic']))
