"""CPU tier: the CUDA kernel bodies (feabas_b200/csrc/fb_xcorr.cuh), executed by the host
emulator (tests/host_emu/emu.cpp, test infrastructure), against the golden vectors of the
unmodified reference and against the oracle.  This checks the kernels' index arithmetic,
radix passes, packing and reductions without a GPU; the -m gpu tier checks the real thing."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, case_kwargs
from oracle import xcorr_oracle as xo
from feabas_b200 import synth
import parity

EMU_DIR = os.path.join(ROOT, 'tests', 'host_emu')
_DT = {np.dtype('float32'): 0, np.dtype('uint8'): 1, np.dtype('float64'): 2}


@pytest.fixture(scope='module')
def emu():
    so = os.path.join(EMU_DIR, 'libfb_emu.so')
    src = os.path.join(EMU_DIR, 'emu.cpp')
    deps = [src] + [os.path.join(ROOT, 'feabas_b200', 'csrc', f) for f in ('fb_xcorr.cuh', 'fb_fft.cuh', 'fb_gfft.cuh', 'fb_host_plan.h')]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(['g++', '-std=c++20', '-O2', '-fPIC', '-shared', '-pthread', '-o', so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.emu_xcorr.restype = ctypes.c_int

    def run(a, b, conf_mode=2, subpixel=False, pad=True, path=0, nthr=8):
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        n, h0, w0 = a.shape
        _, h1, w1 = b.shape
        ny, nx = xo.fft_shape((h0, w0), (h1, w1), pad)
        out = np.zeros((5, n))
        rc = lib.emu_xcorr(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), n, h0, w0, h1, w1,
                           _DT[a.dtype], ny, nx, int(conf_mode), int(subpixel), path,
                           out.ctypes.data_as(ctypes.c_void_p), nthr)
        if rc == -4:
            return None                     # fused path does not fit this problem
        assert rc > 0, rc
        cdt = np.float64 if conf_mode == 1 else np.float32
        return out[0].copy(), out[1].copy(), out[2].astype(cdt), rc
    return run


UNSUPPORTED = ('multichannel', 'normalize_masks', 'normalize_default')


@pytest.mark.parametrize('path', [1, 2])
def test_golden_small_through_kernel_bodies(emu, golden_small, path):
    n = 0
    for name, rec in golden_small.items():
        if name in UNSUPPORTED:
            continue
        kw = case_kwargs(rec)
        res = emu(rec['img0'], rec['img1'], path=path, **kw)
        if res is None:
            assert path == 1
            continue
        dx, dy, cf, rc = res
        assert rc == path
        tol = dict(conf_rtol=5e-2) if kw.get('conf_mode', 2) == 1 else {}   # STD: float32 pow amplifies rounding
        parity.compare(dx, dy, cf, rec['dx'], rec['dy'], rec['conf'], rec['img0'], rec['img1'], **tol, **kw)
        n += 1
    assert n >= (20 if path == 2 else 14)


def test_seeded_mid_size_staged(emu):
    s0, s1, shifts = synth.block_pairs(2, 256, 21, max_shift=32)
    dx, dy, cf, rc = emu(s0, s1, subpixel=True, pad=True)
    assert rc == 2
    parity.check_against_oracle((dx, dy, cf), s0, s1, subpixel=True, pad=True)
    np.testing.assert_array_equal(np.round(dx), shifts[:, 0])
    np.testing.assert_array_equal(np.round(dy), shifts[:, 1])


def test_thread_count_independent(emu):
    s0, s1, _ = synth.block_pairs(2, (60, 75), 12, max_shift=7)
    ref = emu(s0, s1, subpixel=True, nthr=1)
    for nthr in (3, 8, 32):
        got = emu(s0, s1, subpixel=True, nthr=nthr)
        for x, y in zip(ref[:3], got[:3]):
            np.testing.assert_array_equal(x, y)


def test_finalize_one_row_at_a_time(emu, golden_small):
    """K4's narrow variant (long lines: FFT 8192 float32 / 4096 float64, where a 4-line tile no longer fits in
    shared memory) recomputes the rows around the peak one by one: same numbers as the 3-rows-at-once variant."""
    for name in ('roll_f32_sub', 'diffshape_pad', 'stitch_fine_pad', 'uint8_in', 'thumb50_std', 'edge_wrap_nopad', 'odd_sizes'):
        rec = golden_small[name]
        kw = case_kwargs(rec)
        wide = emu(rec['img0'], rec['img1'], path=2, **kw)
        narrow = emu(rec['img0'], rec['img1'], path=3, **kw)
        assert narrow[3] == 3
        for x, y in zip(wide[:3], narrow[:3]):
            np.testing.assert_array_equal(x, y)


def test_register_fft_and_pass_planner(tmp_path):
    """fb_gfft.cuh (compile-time mixed-radix register FFT: the fast path's per-lane transform for lengths with
    factors 3 / 5 and the composite-radix butterfly of the fused kernel's passes) against a direct DFT, forward and
    inverse, float and double; radix planner: products, radix bounds, digit-position tables are permutations."""
    exe = tmp_path / 'gfft_check'
    subprocess.run(['g++', '-std=c++17', '-O1', '-o', str(exe), os.path.join(EMU_DIR, 'gfft_check.cpp')], check=True)
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
