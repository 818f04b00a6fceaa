"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Only usable where
``/root/reference`` exists (the build container); the GPU box has no copy, so
nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this.  Used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
``not gpu`` tests (skipped when the reference is absent) to validate the
restatement directly.

``import feabas.matcher`` needs shapely/h5py/triangle/rtree/pyamg/tensorstore/
skimage/matplotlib/dask, none of which are installed here.  The
functions on the hot path (``xcorr_fft``, ``global_translation_matcher``,
``common.masked_dog_filter``, ``common.divide_bbox`` ...) touch none of them,
so those modules are replaced by inert mocks for the import only.
"""
import importlib
import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('FEABAS_REFERENCE_ROOT', '/root/reference')

_MOCKED = [
    'shapely', 'shapely.geometry', 'shapely.ops', 'shapely.affinity', 'shapely.strtree',
    'triangle', 'rtree', 'rtree.index', 'h5py', 'matplotlib', 'matplotlib.tri',
    'matplotlib.pyplot', 'tensorstore', 'pyamg', 'skimage', 'skimage.morphology',
    'dask', 'dask.distributed', 'dask_jobqueue', 'google', 'google.cloud',
    'google.cloud.storage',
]


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'feabas', 'matcher.py'))


def load():
    """Return the reference modules ``(matcher, common, constant)``."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    for name in _MOCKED:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    matcher = importlib.import_module('feabas.matcher')
    common = importlib.import_module('feabas.common')
    constant = importlib.import_module('feabas.constant')
    return matcher, common, constant
