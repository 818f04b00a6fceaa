"""CPU tier: the C-ABI library builds, loads and exports every symbol include/*.h declares;
host-only entry points work; compute entry points fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import xcorr_oracle as xo


@pytest.fixture(scope='module')
def lib():
    from feabas_b200.csrc import build
    build.build()
    from feabas_b200.cuda import _lib
    return _lib


def _declared_functions():
    names = set()
    inc = os.path.join(ROOT, 'include')
    for fn in os.listdir(inc):
        text = open(os.path.join(inc, fn)).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        names.update(re.findall(r'\b(fb_[a-z0-9_]+)\s*\(', text))
    return names


def test_exports_match_header(lib):
    declared = _declared_functions()
    assert declared == set(lib.SYMBOLS), (declared ^ set(lib.SYMBOLS))
    handle = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name


def test_version_and_next_fast_len(lib):
    assert b'sm_100a' in lib.lib().fb_version()
    for t in list(range(1, 700)) + [1499, 2999, 4097, 8191]:
        assert lib.lib().fb_next_fast_len(t) == xo.next_fast_len_5smooth(t)


def test_plan_info(lib):
    from feabas_b200.cuda import fft_shape
    assert fft_shape((74, 67), (74, 67), True) == (150, 135)
    assert fft_shape((74, 67), (74, 67), False) == (75, 72)
    big = lib.plan_info(512, 512, 512, 512, lib.FB_F32, 1024, 1024, 0x2 | (2 << 2))
    assert big['path'] == 'staged-fast' and big['launches_per_chunk'] == 4
    # F0 + F1 + G(P,Q) ~ 12.6 MB per 512^2 pair
    assert 12.5e6 < big['ws_bytes_per_pair'] < 13.0e6
    small = lib.plan_info(74, 67, 74, 67, lib.FB_F32, 150, 135, 0x2 | (2 << 2))
    assert small['path'] == 'fused-warp' and small['smem_fused'] <= 227 * 1024
    odd = lib.plan_info(72, 72, 72, 72, lib.FB_F32, 144, 144, 0x2 | (2 << 2))     # 144 x 144: fits one SM, not in the warp-fused table
    assert odd['path'] == 'fused'


def test_bad_arguments(lib):
    L = lib.lib()
    info = (ctypes.c_longlong * 8)()
    assert L.fb_xcorr_plan_info(8, 8, 8, 8, 0, 14, 16, 0, info) == -2       # 14 = 2*7 not 5-smooth
    assert b'2^a 3^b 5^c' in L.fb_last_error()
    assert L.fb_xcorr_plan_info(8, 8, 8, 8, 7, 16, 16, 0, info) == -1
    assert L.fb_xcorr_plan_info(8, 8, 8, 8, 0, 4, 16, 0, info) == -1        # grid smaller than the images
    assert L.fb_set_option(b'nope', 1) == -1


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from feabas_b200.cuda import xcorr_fft
    a = np.zeros((1, 8, 8), np.float32)
    with pytest.raises(lib.FeabasCudaError, match='no CUDA device'):
        xcorr_fft(a, a)
    assert lib.launch_count() == 0


def test_image_entry_points_reject_bad_arguments_and_have_no_cpu_fallback(lib):
    """Argument errors are reported before anything touches a device (FB_EINVAL / FB_ESIZE with a message); with valid
    arguments and no GPU every image entry point answers FB_ECUDA "no CUDA device": none computes on the host."""
    import torch
    L = lib.lib()
    buf = np.zeros(64 * 64, np.float64)
    ptr = buf.ctypes.data
    assert L.fb_resize_area_frac(ptr, 1, 64, 64, lib.FB_U8, 0.5, 2.5, ptr, 26, 128, 0, None) == -1        # 1/fx < 1: enlarging
    assert b'shrinks' in L.fb_last_error()
    assert L.fb_resize_area_frac(ptr, 1, 64, 64, lib.FB_F64, 2.5, 2.5, ptr, 26, 26, 0, None) == -1        # dtype
    assert L.fb_resize_area_frac(ptr, 1, 64, 64, lib.FB_U8, 2.5, 2.5, ptr, 40, 26, 0, None) == -1         # output too large
    assert L.fb_masked_dog_f64(ptr, None, 1, 64, 64, 1, 2.5, float('nan'), 0, ptr, ptr, 16, 0, None) == -1   # workspace too small
    assert b'workspace' in L.fb_last_error()
    need = L.fb_masked_dog_f64_workspace(1, 64, 64)
    assert need >= 2 * 64 * 64 * 8
    assert L.fb_masked_dog_f64(ptr, None, 1, 64, 64, 1, 100.0, float('nan'), 0, ptr, ptr, need, 0, None) == -2   # sigma: radius too large
    assert L.fb_masked_dog_f64(ptr, ptr, 3, 64, 64, 2, 2.5, float('nan'), 0, ptr, ptr, L.fb_masked_dog_f64_workspace(3, 64, 64), 0, None) == -1
    assert b'mask_n' in L.fb_last_error()
    assert L.fb_profile_read(0, ctypes.c_void_p(-1), None, None, 0) in (0, -3)      # FB_ALL_STREAMS: no context yet is not an error
    if not torch.cuda.is_available():
        assert L.fb_resize_area_frac(ptr, 1, 64, 64, lib.FB_U8, 2.5, 2.5, ptr, 26, 26, 0, None) == -3
        assert b'no CUDA device' in L.fb_last_error()
        assert L.fb_masked_dog_f64(ptr, None, 1, 64, 64, 1, 2.5, float('nan'), 0, ptr, ptr, need, 0, None) == -3
        assert b'no CUDA device' in L.fb_last_error()
        assert lib.launch_count() == 0
