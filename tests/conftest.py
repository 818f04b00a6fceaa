import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def _gpu_ready():
    try:
        import torch
        if not torch.cuda.is_available():
            return False, 'no CUDA device'
    except Exception as exc:                      # pragma: no cover
        return False, f'torch unavailable: {exc}'
    if not os.path.exists(os.path.join(ROOT, 'feabas_b200', 'csrc', 'libfeabas_cuda.so')):
        return False, 'libfeabas_cuda.so is not built (python -m feabas_b200.csrc.build)'
    return True, ''


def pytest_collection_modifyitems(config, items):
    """A plain ``pytest tests`` on a CPU-only box skips the GPU tier instead of failing in it; with ``-m gpu``
    asked for explicitly (the GPU box) nothing is skipped -- a missing device or library must fail loudly there."""
    if 'gpu' in (config.getoption('-m') or ''):
        return
    ok, why = _gpu_ready()
    if ok:
        return
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Return {case: {key: array}} from a flat npz written by oracle/make_golden.py."""
    out = {}
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        for key in z.files:
            case, _, rest = key.partition('/')
            out.setdefault(case, {})[rest] = z[key]
    return out


def case_kwargs(rec):
    kw = {}
    for k, v in rec.items():
        if k.startswith('kw/'):
            v = v if v.ndim else v.item()
            kw[k[3:]] = v
    return kw


@pytest.fixture(scope='session')
def golden_small():
    return load_golden('xcorr_small.npz')


@pytest.fixture(scope='session')
def golden_seeded():
    return load_golden('xcorr_seeded.npz')


@pytest.fixture(scope='session')
def golden_host():
    out = {}
    with np.load(os.path.join(GOLDEN, 'matcher_host.npz'), allow_pickle=False) as z:
        for key in z.files:
            out[key] = z[key]
    return out
