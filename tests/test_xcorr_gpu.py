"""GPU tier (-m gpu): the CUDA path, called through the C ABI (ctypes shim in
feabas_b200.cuda), against the golden vectors of the unmodified reference, against the oracle
on seeded inputs, and -- at the benchmark's full size -- against the synthetic ground truth."""
import numpy as np
import pytest

from conftest import case_kwargs
from feabas_b200 import synth
import parity

pytestmark = pytest.mark.gpu

UNSUPPORTED = ()


@pytest.fixture(scope='module')
def fc():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import feabas_b200.cuda as fc
    return fc


def _tol(kw):
    # FFT_CONF_STD raises a float32-rounded base to the power ny*nx (matcher.py:133): one ulp of
    # the base moves the result by ~ny*nx*6e-8 relative, in the reference itself as well.
    return dict(conf_rtol=5e-2) if kw.get('conf_mode', 2) == 1 else {}


@pytest.mark.parametrize('force', [None, 'staged', 'fused', 'fused_smem'])
def test_golden_small(fc, golden_small, force):
    n = 0
    for name, rec in golden_small.items():
        if name in UNSUPPORTED:
            continue
        kw = case_kwargs(rec)
        try:
            dx, dy, cf = fc.xcorr_fft(rec['img0'], rec['img1'], force=force, **kw)
        except fc._lib.FeabasCudaError as e:
            assert force in ('fused', 'fused_smem') and 'unavailable' in str(e), (name, e)
            continue
        parity.compare(dx, dy, cf, rec['dx'], rec['dy'], rec['conf'], rec['img0'], rec['img1'], **_tol(kw), **kw)
        n += 1
    assert n >= (14 if force in ('fused', 'fused_smem') else 20)
    assert fc._lib.launch_count() > 0


def test_golden_seeded(fc, golden_seeded):
    for name, rec in golden_seeded.items():
        s0, s1, shifts = synth.block_pairs(int(rec['n']), rec['size'].tolist(), int(rec['seed']), max_shift=int(rec['max_shift']))
        kw = case_kwargs(rec)
        dx, dy, cf = fc.xcorr_fft(s0, s1, **kw)
        parity.compare(dx, dy, cf, rec['dx'], rec['dy'], rec['conf'], s0, s1, **kw)
        np.testing.assert_array_equal(np.round(dx), shifts[:, 0])
        np.testing.assert_array_equal(np.round(dy), shifts[:, 1])


@pytest.mark.parametrize('shape0,shape1,kw', [
    ((5, 74, 67), (5, 74, 67), dict(subpixel=True)),
    ((5, 74, 67), (5, 74, 67), dict(subpixel=True, pad=False)),
    ((3, 33, 20), (3, 50, 64), dict(subpixel=True, pad=False)),
    ((3, 50, 64), (3, 33, 20), dict(subpixel=True)),
    ((4, 150, 150), (4, 150, 150), dict(subpixel=True)),                  # thumbnail 300^2 FFT
    ((2, 280, 280), (2, 280, 280), dict(subpixel=True, conf_mode=0)),     # alignment 576^2 FFT
    ((2, 250, 1500), (2, 250, 1500), dict(subpixel=False)),               # coarse strip 500 x 3000
    ((1, 2000, 200), (1, 2000, 200), dict(subpixel=False)),               # coarse strip 4000 x 400
    ((3, 256, 256), (3, 256, 256), dict(subpixel=True, conf_mode=1)),
    ((2, 100, 90), (2, 100, 90), dict(subpixel=True, conf_mode=1, pad=False)),
])
def test_random_against_oracle(fc, shape0, shape1, kw):
    rng = np.random.default_rng(abs(hash((shape0, shape1))) % (2 ** 31))
    if shape0 == shape1:
        a, b, _ = synth.block_pairs(shape0[0], shape0[1:], seed=shape0[1], max_shift=min(shape0[1:]) // 8)
    else:
        a = rng.standard_normal(shape0).astype(np.float32)
        b = rng.standard_normal(shape1).astype(np.float32)
    got = fc.xcorr_fft(a, b, **kw)
    parity.check_against_oracle(got, a, b, **_tol(kw), **kw)


@pytest.mark.parametrize('dtype', [np.uint8, np.float64])
def test_float64_pipeline(fc, dtype):
    a, b, _ = synth.block_pairs(3, 96, seed=5, max_shift=10, band_pass=False)
    a, b = a.clip(0, 255).astype(dtype), b.clip(0, 255).astype(dtype)
    for kw in (dict(subpixel=True), dict(subpixel=True, pad=False), dict(subpixel=True, conf_mode=1)):
        got = fc.xcorr_fft(a, b, **kw)
        parity.check_against_oracle(got, a, b, conf_rtol=1e-6 if kw.get('conf_mode', 2) != 1 else 5e-2, **kw)


def test_sigma_masks_normalize_multichannel(fc):
    """The rarely used arguments of xcorr_fft (matcher.py:44-56,66-81): DoG inside the call, mask
    normalisation with explicit / default masks, channel-mean of the cross-power."""
    from oracle import matcher_oracle as mo
    rng = np.random.default_rng(21)
    a, b, _ = synth.block_pairs(3, (90, 70), seed=9, max_shift=8, band_pass=False)
    m0 = np.ones((90, 70), bool); m0[:, :15] = False
    m1 = np.ones((90, 70), bool); m1[60:, :] = False
    # sigma > 0 with and without masks: oracle = masked DoG, then xcorr
    for masks in ((None, None), (m0, m1)):
        fa = mo.masked_dog_oracle(a, 2.5, mask=masks[0])
        fb = mo.masked_dog_oracle(b, 2.5, mask=masks[1])
        got = fc.xcorr_fft(a, b, sigma=2.5, mask0=masks[0], mask1=masks[1], subpixel=True)
        parity.check_against_oracle(got, fa.astype(np.float32), fb.astype(np.float32), subpixel=True)
    # normalize, float32 masks (the reference's default mask dtype) in every confidence mode
    f0, f1 = m0.astype(np.float32), m1.astype(np.float32)
    fa, fb, _ = synth.block_pairs(3, (90, 70), seed=10, max_shift=8)
    for kw in (dict(subpixel=True), dict(subpixel=True, pad=False), dict(subpixel=True, conf_mode=0), dict(subpixel=False, conf_mode=1)):
        for masks in ((None, None), (f0, f1)):
            got = fc.xcorr_fft(fa, fb, normalize=True, mask0=masks[0], mask1=masks[1], **kw)
            parity.check_against_oracle(got, fa, fb, normalize=True, mask0=masks[0], mask1=masks[1], **_tol(kw), **kw)
    # multi-channel, incl. the float64 pipeline and more pairs than one chunk holds
    c0 = rng.standard_normal((5, 40, 56, 3)).astype(np.float32)
    c1 = np.roll(c0, (3, -5), axis=(1, 2)) + 0.1 * rng.standard_normal(c0.shape).astype(np.float32)
    for kw in (dict(subpixel=True), dict(subpixel=True, pad=False, conf_mode=1)):
        parity.check_against_oracle(fc.xcorr_fft(c0, c1, **kw), c0, c1, **_tol(kw), **kw)
    u0 = (c0 * 40 + 128).clip(0, 255).astype(np.uint8)
    u1 = (c1 * 40 + 128).clip(0, 255).astype(np.uint8)
    parity.check_against_oracle(fc.xcorr_fft(u0, u1, subpixel=True), u0, u1, conf_rtol=1e-6, subpixel=True)
    fc._lib.set_option('ws_bytes', 1 << 20)
    try:
        parity.check_against_oracle(fc.xcorr_fft(c0, c1, subpixel=True), c0, c1, subpixel=True)
    finally:
        fc._lib.set_option('ws_bytes', 2 << 30)


def test_device_tensors_and_host_arrays_agree(fc):
    import torch
    a, b, _ = synth.block_pairs(6, 128, seed=3, max_shift=16)
    host = fc.xcorr_fft(a, b, subpixel=True)
    dev = fc.xcorr_fft(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), subpixel=True)
    pinned = fc.xcorr_fft(torch.from_numpy(a).pin_memory(), torch.from_numpy(b).pin_memory(), subpixel=True)
    for x, y, z in zip(host, dev, pinned):
        np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(x, z)


def test_empty_batch(fc):
    a = np.zeros((0, 16, 16), np.float32)
    dx, dy, cf = fc.xcorr_fft(a, a)
    assert dx.shape == dy.shape == cf.shape == (0,)


def test_host_path_chunking(fc):
    """Chunked, double-buffered host path must not depend on the chunk size."""
    a, b, _ = synth.block_pairs(37, 64, seed=8, max_shift=8)
    ref = fc.xcorr_fft(a, b, subpixel=True)
    fc._lib.set_option('host_chunk_bytes', 5 * 2 * 64 * 64 * 4)     # 5 pairs per chunk
    fc._lib.set_option('ws_bytes', 1 << 20)
    try:
        for force in (None, 'staged'):
            got = fc.xcorr_fft(a, b, subpixel=True, force=force)
            for x, y in zip(ref, got):
                np.testing.assert_allclose(x, y, rtol=0, atol=1e-4)
    finally:
        fc._lib.set_option('host_chunk_bytes', 64 << 20)
        fc._lib.set_option('ws_bytes', 2 << 30)


def test_full_size_ground_truth(fc):
    """Benchmark-sized work (512^2 blocks, FFT 1024^2): size-independent properties.
    (1) the synthetic ground-truth shift is recovered for every pair; (2) swapping the two
    stacks negates the displacement; (3) results do not depend on batch composition."""
    import torch
    n = 24
    a, b, shifts = synth.block_pairs(n, 512, seed=4, max_shift=32)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = fc.xcorr_fft_device(ta, tb, subpixel=True).cpu().numpy()
    np.testing.assert_array_equal(np.round(out[0]), shifts[:, 0])
    np.testing.assert_array_equal(np.round(out[1]), shifts[:, 1])
    assert np.all(out[2] > 0.5)
    swapped = fc.xcorr_fft_device(tb, ta, subpixel=True).cpu().numpy()
    np.testing.assert_allclose(swapped[0], -out[0], atol=0.02)
    np.testing.assert_allclose(swapped[1], -out[1], atol=0.02)
    np.testing.assert_allclose(swapped[2], out[2], rtol=1e-4, atol=1e-6)
    part = fc.xcorr_fft_device(ta[5:9].contiguous(), tb[5:9].contiguous(), subpixel=True).cpu().numpy()
    np.testing.assert_array_equal(part, out[:, 5:9])
    # and the first two pairs against the oracle
    parity.check_against_oracle((out[0, :2], out[1, :2], out[2, :2].astype(np.float32)), a[:2], b[:2], subpixel=True)


FAST_CASES = [
    # (n, shape0, shape1, kwargs) -- every FFT grid here is a power of two in {256, 512, 1024, 2048, 4096}
    (5, (128, 128), (128, 128), dict(subpixel=True)),                          # 256 x 256
    (3, (127, 127), (127, 127), dict(subpixel=True)),                          # 256 x 256, ragged rows
    (3, (110, 120), (147, 137), dict(subpixel=True)),                          # 256 x 256, different shapes
    (3, (256, 256), (256, 256), dict(subpixel=True)),                          # 512 x 512
    (3, (256, 256), (256, 256), dict(subpixel=True, pad=False)),               # 256 x 256, no pruning
    (2, (512, 512), (512, 512), dict(subpixel=True)),                          # 1024 x 1024
    (2, (512, 512), (512, 512), dict(subpixel=True, pad=False)),               # 512 x 512 circular
    (2, (1024, 256), (1024, 256), dict(subpixel=True, pad=False)),             # 1024 x 256
    (2, (128, 512), (128, 512), dict(subpixel=False)),                         # 256 x 1024
    (3, (256, 128), (256, 128), dict(subpixel=True, conf_mode=0)),             # 512 x 256, NONE
    (3, (256, 256), (256, 256), dict(subpixel=True, conf_mode=1)),             # STD
    (3, (255, 255), (255, 255), dict(subpixel=True, conf_mode=0, pad=False)),  # 256^2 odd rows, NONE pairs rows
    (2, (501, 512), (501, 512), dict(subpixel=True, conf_mode=1, pad=False)),  # 512 x 512 (odd rows)
    (2, (1024, 1024), (1024, 1024), dict(subpixel=True)),                      # 2048 x 2048 (64 points per lane)
    (2, (1024, 128), (1024, 128), dict(subpixel=True)),                        # 2048 x 256
    (2, (256, 2048), (256, 2048), dict(subpixel=True, pad=False)),             # 256 x 2048, no pruning
    (2, (1020, 1017), (1020, 1017), dict(subpixel=True, conf_mode=1)),         # 2048 x 2048 ragged, STD
    (2, (2048, 512), (2048, 512), dict(subpixel=False, conf_mode=0, pad=False)),  # 2048 x 512, unpruned columns, NONE
    (1, (2048, 2048), (2048, 2048), dict(subpixel=True)),                      # 4096 x 4096 (a line spans two warps)
    (2, (2040, 128), (2040, 128), dict(subpixel=True)),                        # 4096 x 256 (ragged), column pieces
    (2, (128, 2048), (128, 2048), dict(subpixel=True, conf_mode=1)),           # 256 x 4096, STD
    (2, (4096, 256), (4096, 256), dict(subpixel=True, conf_mode=0, pad=False)),  # 4096 x 256 unpruned, NONE
    # 5-smooth grids with a radix-3 factor: 576 = 24 x 24 and 288 = 24 x 12 points (lanes 24..31 shadow lanes 0..7)
    (3, (280, 280), (280, 280), dict(subpixel=True)),                          # 576 x 576: default fine-alignment blocks
    (3, (140, 140), (140, 140), dict(subpixel=True)),                          # 288 x 288
    (2, (285, 281), (285, 281), dict(subpixel=True, conf_mode=1)),             # 576 x 576 ragged, STD
    (2, (288, 576), (288, 576), dict(subpixel=True, pad=False)),               # 288 x 576 unpruned
    (2, (280, 512), (280, 512), dict(subpixel=True, conf_mode=0)),             # 576 x 1024, NONE
    (2, (512, 140), (512, 140), dict(subpixel=False)),                         # 1024 x 288
    (3, (130, 130), (150, 150), dict(subpixel=True)),                          # 288 x 288, different shapes
    # 300 = 30 x 10 points (radix 5; three lines per warp, K3 tiles of 4 rows with shadow threads)
    (4, (150, 150), (150, 150), dict(subpixel=True)),                          # 300 x 300: thumbnail blocks
    (3, (300, 300), (300, 300), dict(subpixel=True, pad=False)),               # 300 x 300 unpruned
    (2, (141, 150), (141, 150), dict(subpixel=False)),                         # 288 x 300 (ragged rows)
    # further 5-smooth lengths: 200 = 20 x 10, 400 = 40 x 10, 800 = 40 x 20, 384 = 48 x 8, 768 = 48 x 16, 1152 = 48 x 24
    (3, (100, 100), (100, 100), dict(subpixel=True)),                          # 200 x 200
    (2, (200, 200), (200, 200), dict(subpixel=True, conf_mode=1)),             # 400 x 400, STD
    (2, (400, 400), (400, 400), dict(subpixel=True)),                          # 800 x 800
    (2, (192, 192), (192, 192), dict(subpixel=True, conf_mode=0)),             # 384 x 384, NONE
    (2, (384, 384), (384, 384), dict(subpixel=True)),                          # 768 x 768
    (2, (576, 576), (576, 576), dict(subpixel=True)),                          # 1152 x 1152
    (2, (100, 400), (100, 400), dict(subpixel=True)),                          # 200 x 800
    (2, (400, 200), (400, 200), dict(subpixel=True, pad=False)),               # 400 x 200 unpruned
    # 60 / 50 points per lane: 600 = 60 x 10, 500 = 50 x 10, 1200 = 60 x 20, 720 = 60 x 12
    (2, (300, 300), (300, 300), dict(subpixel=True)),                          # 600 x 600
    (2, (250, 250), (250, 250), dict(subpixel=True)),                          # 500 x 500 (K3 tiles of 4 rows)
    (2, (600, 600), (600, 600), dict(subpixel=True, conf_mode=1)),             # 1200 x 1200, STD
    (2, (360, 360), (360, 360), dict(subpixel=True, conf_mode=0)),             # 720 x 720, NONE
    (2, (600, 720), (600, 720), dict(subpixel=True, pad=False)),               # 600 x 720 unpruned
]


@pytest.mark.parametrize('n,shape0,shape1,kw', FAST_CASES)
def test_fast_path_against_oracle(fc, n, shape0, shape1, kw):
    from feabas_b200.cuda import _lib
    ny, nx = fc.fft_shape(shape0, shape1, kw.get('pad', True))
    from feabas_b200.cuda.xcorr import _flags
    info = _lib.plan_info(*shape0, *shape1, _lib.FB_F32, ny, nx, _flags(kw.get('conf_mode', 2), kw.get('subpixel', False), kw.get('pad', True)))
    assert info['path'] == 'staged-fast', (ny, nx, info)
    if shape0 == shape1:
        a, b, _ = synth.block_pairs(n, shape0, seed=shape0[0] + shape0[1], max_shift=min(shape0) // 8)
    else:
        rng = np.random.default_rng(shape0[0])
        b = rng.standard_normal((n,) + shape1).astype(np.float32)
        a = b[:, 20:20 + shape0[0], 10:10 + shape0[1]] + 0.1 * rng.standard_normal((n,) + shape0).astype(np.float32)
        a = np.ascontiguousarray(a)
    got = fc.xcorr_fft(a, b, **kw)
    parity.check_against_oracle(got, a, b, **_tol(kw), **kw)
    gen = fc.xcorr_fft(a, b, force='generic', **kw)
    np.testing.assert_allclose(got[0], gen[0], atol=2e-3)
    np.testing.assert_allclose(got[1], gen[1], atol=2e-3)


@pytest.mark.parametrize('size,n', [(128, 48), (256, 20), (140, 60), (280, 16), (150, 80)])
def test_fast_path_many_work_items_per_cta(fc, size, n):
    """More (pair, column group) / (pair, tile) work items than resident CTAs: the persistent loops of the
    fast-path kernels (TMA store / load recycling, odd last column group) run several iterations."""
    a, b, shifts = synth.block_pairs(n, size, seed=size + n, max_shift=size // 8)
    got = fc.xcorr_fft(a, b, subpixel=True)
    gen = fc.xcorr_fft(a, b, subpixel=True, force='generic')
    np.testing.assert_array_equal(np.round(got[0]), shifts[:, 0])
    np.testing.assert_array_equal(np.round(got[1]), shifts[:, 1])
    np.testing.assert_allclose(got[0], gen[0], atol=2e-3)
    np.testing.assert_allclose(got[1], gen[1], atol=2e-3)
    np.testing.assert_allclose(got[2], gen[2], rtol=1e-4, atol=1e-6)
    parity.check_against_oracle(tuple(v[:3] for v in got), a[:3], b[:3], subpixel=True)


@pytest.mark.parametrize('size,n', [(280, 12), (150, 24), (1024, 6), (100, 40)])
def test_fast_sizes_size_independent_properties(fc, size, n):
    """Grids served by the wider fast path (576 = 24 x 24, 300 = 30 x 10, 2048 = 64 x 32, 200 = 20 x 10 points per
    line), through the two-stream schedule (n >= 64 pairs would be needed to split -- here the serial one) and the
    pipelined one (option): ground truth recovered, swapping the stacks negates the displacement, results do not
    depend on batch composition or on the schedule."""
    import torch
    a, b, shifts = synth.block_pairs(n, size, seed=size + 1, max_shift=min(32, size // 8))
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = fc.xcorr_fft_device(ta, tb, subpixel=True).cpu().numpy()
    np.testing.assert_array_equal(np.round(out[0]), shifts[:, 0])
    np.testing.assert_array_equal(np.round(out[1]), shifts[:, 1])
    swapped = fc.xcorr_fft_device(tb, ta, subpixel=True).cpu().numpy()
    np.testing.assert_allclose(swapped[0], -out[0], atol=0.02)
    np.testing.assert_allclose(swapped[1], -out[1], atol=0.02)
    np.testing.assert_allclose(swapped[2], out[2], rtol=1e-4, atol=1e-6)
    part = fc.xcorr_fft_device(ta[2:5].contiguous(), tb[2:5].contiguous(), subpixel=True).cpu().numpy()
    np.testing.assert_array_equal(part, out[:, 2:5])


def test_two_stream_schedule_matches_serial(fc):
    """fb_set_option('pipeline', 2) (default: a chunk runs as two halves on two streams) and 1 (serial) give
    bit-identical results; so does a workspace so small that the batch is cut into several chunks."""
    import torch
    n = 150
    a, b, shifts = synth.block_pairs(n, 128, seed=77, max_shift=16)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    L = fc._lib
    try:
        L.set_option('pipeline', 1)
        serial = fc.xcorr_fft_device(ta, tb, subpixel=True).cpu().numpy()
        L.set_option('pipeline', 2)
        two = fc.xcorr_fft_device(ta, tb, subpixel=True).cpu().numpy()
        L.set_option('ws_bytes', 70 * 450000)                     # ~70 pairs per chunk: chunks of 70 / 70 / 10 pairs
        chunked = fc.xcorr_fft_device(ta, tb, subpixel=True).cpu().numpy()
    finally:
        L.set_option('pipeline', 2)
        L.set_option('ws_bytes', 2 << 30)
    np.testing.assert_array_equal(serial, two)
    np.testing.assert_array_equal(serial, chunked)
    np.testing.assert_array_equal(np.round(two[0]), shifts[:, 0])


# warp-fused kernel (fb_xcorr_wf.cuh): every grid of its size table, the three confidence modes, equal and different
# image shapes, padded and not, float32 and uint8 (computed in float32 on request)
WF_CASES = [
    (40, (74, 67), (74, 67), dict(subpixel=True)),                               # FFT 150 x 135
    (40, (74, 67), (74, 67), dict(subpixel=True, pad=False)),                    # 75 x 72
    (24, (60, 75), (60, 75), dict(subpixel=True)),                               # 120 x 150
    (24, (60, 75), (60, 75), dict(subpixel=False, pad=False)),                   # 60 x 75
    (24, (60, 67), (60, 67), dict(subpixel=True)),                               # 120 x 135
    (50, (50, 50), (50, 50), dict(subpixel=True)),                               # 100 x 100
    (50, (50, 50), (50, 50), dict(subpixel=True, pad=False)),                    # 50 x 50
    (30, (64, 64), (64, 64), dict(subpixel=True)),                               # 128 x 128
    (30, (64, 64), (64, 64), dict(subpixel=True, pad=False)),                    # 64 x 64
    (13, (74, 67), (74, 67), dict(subpixel=True, conf_mode=0)),
    (13, (74, 67), (74, 67), dict(subpixel=True, conf_mode=1)),
    (13, (73, 66), (75, 68), dict(subpixel=True)),                               # different shapes, odd heights, same grid
    (7, (50, 50), (48, 50), dict(subpixel=True, conf_mode=1)),                   # 100 x 100 from unequal blocks, STD
    (5, (75, 72), (75, 72), dict(subpixel=True, pad=False, conf_mode=0)),        # no free rows to stage the images in
]


@pytest.mark.parametrize('n,shape0,shape1,kw', WF_CASES)
def test_warp_fused_against_oracle(fc, n, shape0, shape1, kw):
    from feabas_b200.cuda import _lib, fft_shape
    ny, nx = fft_shape(shape0, shape1, kw.get('pad', True))
    info = _lib.plan_info(shape0[0], shape0[1], shape1[0], shape1[1], _lib.FB_F32, ny, nx, 0)
    assert info['path'] == 'fused-warp', (ny, nx, info)
    if shape0 == shape1:
        a, b, shifts = synth.block_pairs(n, shape0, seed=sum(shape0) + n, max_shift=min(shape0) // 8)
    else:
        big = (max(shape0[0], shape1[0]) + 8, max(shape0[1], shape1[1]) + 8)
        c0, c1, _ = synth.block_pairs(n, big, seed=77, max_shift=3)
        a = np.ascontiguousarray(c0[:, 2:2 + shape0[0], 3:3 + shape0[1]])
        b = np.ascontiguousarray(c1[:, 4:4 + shape1[0], 1:1 + shape1[1]])
        shifts = None
    got = fc.xcorr_fft(a, b, **kw)
    parity.check_against_oracle(got, a, b, **_tol(kw), **kw)
    if shifts is not None and kw.get('pad', True):
        assert np.mean((np.round(got[0]) == shifts[:, 0]) & (np.round(got[1]) == shifts[:, 1])) > 0.95
    # the first fused kernel (shared-memory passes) agrees within the same gates
    old = fc.xcorr_fft(a, b, force='fused_smem', **kw)
    parity.compare(got[0], got[1], got[2], old[0], old[1], old[2], a, b, **_tol(kw), **kw)


def test_warp_fused_uint8_as_float32_and_large_batch(fc):
    import torch
    a, b, shifts = synth.block_pairs(700, (74, 67), seed=5, max_shift=9, dtype=np.float32, band_pass=False)   # > 2 waves of CTAs
    ua, ub = a.clip(0, 255).astype(np.uint8), b.clip(0, 255).astype(np.uint8)
    out = fc.xcorr_fft_device(torch.from_numpy(ua).cuda(), torch.from_numpy(ub).cuda(), subpixel=True, u8_as_f32=True).cpu().numpy()
    ref = fc.xcorr_fft(ua.astype(np.float32), ub.astype(np.float32), subpixel=True)
    np.testing.assert_array_equal(out[0], ref[0])
    np.testing.assert_array_equal(out[1], ref[1])
    np.testing.assert_array_equal(out[2].astype(np.float32), ref[2])
    parity.check_against_oracle(ref, ua.astype(np.float32), ub.astype(np.float32), subpixel=True)   # (raw, un-band-passed pixels)
