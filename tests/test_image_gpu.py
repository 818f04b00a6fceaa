"""GPU tier: image operators either side of the matcher (fb_masked_dog, fb_resize_area, fb_resize_nearest,
fb_crop_blocks, fb_stack_minmax) against golden vectors of the unmodified reference, the oracle, and the
third-party calls the reference itself makes (cv2.resize / cv2.remap, scipy.ndimage)."""
import numpy as np
import pytest

from feabas_b200 import synth
from oracle import matcher_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fc():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import feabas_b200.cuda as fc
    return fc


@pytest.mark.parametrize('name', ['plain', 'masked', 'u8', 'stack_masked', 'allmask_true'])
def test_dog_golden(fc, golden_host, name):
    img, sigma = golden_host[f'dog/{name}/img'], golden_host[f'dog/{name}/sigma'].item()
    mask = golden_host.get(f'dog/{name}/mask')
    want = golden_host[f'dog/{name}/out']
    exact = fc.masked_dog_filter(img, sigma, mask=mask, exact=True)
    assert exact.dtype == want.dtype and exact.shape == want.shape
    span = float(np.ptp(img.astype(np.float32)))
    # float64 accumulation in scipy's order: identical up to the last float32 ulp of the Gaussian taps
    np.testing.assert_allclose(exact, want, rtol=0, atol=2e-7 * span)
    assert np.mean(exact == want) > 0.99
    fast = fc.masked_dog_filter(img, sigma, mask=mask)
    np.testing.assert_allclose(fast, want, rtol=0, atol=3e-6 * span)


def test_dog_random_against_oracle(fc):
    import torch
    rng = np.random.default_rng(11)
    for shape, sigma in [((3, 130, 97), 2.5), ((1, 64, 300), 1.25), ((2, 200, 200), 3.5), ((70, 45), 6.0)]:
        img = (rng.standard_normal(shape) * 30 + 120).astype(np.float32)
        mask = rng.random(shape) > 0.2
        for m in (None, mask):
            for signed in (True, False):
                want = mo.masked_dog_oracle(img, sigma, mask=m, signed=signed)
                got = fc.masked_dog_filter(img, sigma, mask=m, signed=signed, exact=True)
                np.testing.assert_allclose(got, want, rtol=0, atol=3e-7 * np.ptp(img))
        # device tensors in -> device tensor out, explicit ptp (a shard of a bigger stack)
        t = fc.masked_dog_filter(torch.from_numpy(img).cuda(), sigma, mask=torch.from_numpy(mask).cuda(), ptp=500.0, exact=True)
        assert t.is_cuda
        np.testing.assert_allclose(t.cpu().numpy(), mo.masked_dog_oracle(img, sigma, mask=mask, ptp=np.float32(500.0)),
                                   rtol=0, atol=3e-7 * 500)
    u8 = rng.integers(0, 256, (2, 90, 110), dtype=np.uint8)
    np.testing.assert_allclose(fc.masked_dog_filter(u8, 2.5, exact=True), mo.masked_dog_oracle(u8, 2.5), rtol=0, atol=1e-4)


@pytest.mark.parametrize('k', [2, 3, 4])
def test_resize_area_vs_cv2(fc, k):
    rng = np.random.default_rng(k)
    for shape in [(300, 500), (301, 501), (303, 505), (64, 64), (7, 9)]:
        u8 = rng.integers(0, 256, shape, dtype=np.uint8)
        want = mo.resize_area_oracle(u8, 1 / k)
        got = fc.resize_area(u8, 1 / k)
        assert got.shape == want.shape and got.dtype == np.uint8
        np.testing.assert_array_equal(got, want)
        f32 = (rng.standard_normal(shape) * 40 + 128).astype(np.float32)
        np.testing.assert_allclose(fc.resize_area(f32, 1 / k), mo.resize_area_oracle(f32, 1 / k), rtol=1e-6, atol=1e-4)
        m = rng.random(shape) > 0.5
        np.testing.assert_array_equal(fc.resize_mask(m, 1 / k), mo.resize_mask_oracle(m, 1 / k))
    with pytest.raises(NotImplementedError):
        fc.resize_area(np.zeros((10, 10), np.uint8), 1.5)      # enlarging: not INTER_AREA proper, no FEABAS caller


@pytest.mark.parametrize('factor', [0.4, 0.3, 0.7, 0.9, 0.45, 0.15, 0.8, 1 / 2.5, 0.999])
def test_resize_area_any_factor_vs_cv2(fc, factor):
    """coarse_downsample / fine_downsample are free parameters (matcher.py:233-234): OpenCV's general INTER_AREA path,
    bit-exact for uint8 AND float32 (same tables, same float32 operation order)."""
    import torch
    rng = np.random.default_rng(int(factor * 1000))
    for shape in [(301, 517), (64, 64), (500, 333), (9, 7), (64, 700), (3, 200, 130)]:
        u8 = rng.integers(0, 256, shape, dtype=np.uint8)
        want = np.stack([mo.resize_area_oracle(x, factor) for x in u8.reshape((-1,) + shape[-2:])]).reshape(shape[:-2] + (-1,))
        want = want.reshape(shape[:-2] + mo.resize_area_oracle(u8.reshape((-1,) + shape[-2:])[0], factor).shape)
        got = fc.resize_area(u8, factor)
        assert got.shape == want.shape and got.dtype == np.uint8
        np.testing.assert_array_equal(got, want)
        f32 = (rng.standard_normal(shape[-2:]) * 40 + 128).astype(np.float32)
        np.testing.assert_array_equal(fc.resize_area(f32, factor), mo.resize_area_oracle(f32, factor))
        m = rng.random(shape[-2:]) > 0.5
        np.testing.assert_array_equal(fc.resize_mask(m, factor), mo.resize_mask_oracle(m, factor))
    t = fc.resize_area(torch.from_numpy(u8[0]).cuda(), factor)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), mo.resize_area_oracle(u8[0], factor))


def test_dog_float64(fc):
    """common.py:363 converts only non-floating dtypes: float64 images are filtered and returned in float64."""
    import torch
    rng = np.random.default_rng(5)
    for shape, sigma in [((2, 90, 131), 2.5), ((70, 45), 6.0), ((1, 40, 300), 1.0), ((33, 29), 12.0)]:
        img = rng.standard_normal(shape) * 30 + 120
        mask = rng.random(shape) > 0.2
        one = rng.random(shape[-2:]) > 0.3
        for m in (None, mask, one):
            for signed in (True, False):
                want = mo.masked_dog_oracle(img, sigma, mask=m, signed=signed)
                got = fc.masked_dog_filter(img, sigma, mask=m, signed=signed)
                assert got.dtype == np.float64 and got.shape == want.shape
                # every operation in double in scipy's order; exp() of the taps may differ in the last ulp
                np.testing.assert_allclose(got, want, rtol=0, atol=1e-13 * np.ptp(img))
        t = fc.masked_dog_filter(torch.from_numpy(img).cuda(), sigma, mask=torch.from_numpy(mask).cuda(), ptp=500.0)
        assert t.is_cuda and t.dtype == torch.float64
        np.testing.assert_allclose(t.cpu().numpy(), mo.masked_dog_oracle(img, sigma, mask=mask, ptp=500.0), rtol=0, atol=1e-13 * 500)
    # float64 straight into xcorr_fft with sigma > 0 (matcher.py:54-58): complex128 pipeline
    a = rng.standard_normal((2, 64, 64)) * 20 + 100
    b = np.roll(a, (3, -2), axis=(1, 2))
    from oracle import xcorr_oracle as xo
    got = fc.xcorr_fft(a, b, sigma=2.0, subpixel=True)
    want = xo.xcorr_oracle(mo.masked_dog_oracle(a, 2.0), mo.masked_dog_oracle(b, 2.0), subpixel=True)
    np.testing.assert_allclose(got[0], want[0], atol=1e-6)
    np.testing.assert_allclose(got[1], want[1], atol=1e-6)
    np.testing.assert_allclose(got[2], want[2], rtol=1e-6)


def _blocks(boxes, ainv, tinv):
    b = np.asarray(boxes, dtype=np.float64)
    rows = np.empty((len(b), 10))
    rows[:, 0], rows[:, 1] = b[:, 0], b[:, 1]
    rows[:, 2] = (b[:, 2] - b[:, 0]) / np.round(b[:, 2] - b[:, 0])
    rows[:, 3] = (b[:, 3] - b[:, 1]) / np.round(b[:, 3] - b[:, 1])
    rows[:, 4], rows[:, 5], rows[:, 6] = ainv[0, 0], ainv[1, 0], tinv[0]
    rows[:, 7], rows[:, 8], rows[:, 9] = ainv[0, 1], ainv[1, 1], tinv[1]
    return rows


@pytest.mark.parametrize('dtype', [np.uint8, np.float32])
def test_crop_blocks_vs_cv2_remap(fc, dtype):
    import torch
    from feabas_b200.cuda import image as im
    rng = np.random.default_rng(3)
    canvas = synth.em_canvas(400, 600, seed=9)
    img = canvas if dtype == np.uint8 else synth.dog_f32(canvas)
    t = torch.from_numpy(img).cuda()
    boxes = np.array([[10, 20, 84, 87], [-30, -12, 44, 55], [560, 350, 634, 417], [300, 100, 374, 167]])
    cases = [(np.eye(2), np.array([0.0, 0.0])),                      # identity: exact copies with fill outside
             (np.eye(2), np.array([-17.0, 5.0])),                    # integer translation
             (np.eye(2), np.array([3.3, -7.77])),                    # fractional translation
             (np.array([[1.01, 0.02], [-0.015, 0.99]]), np.array([4.2, -3.1]))]
    for ainv, tinv in cases:
        want, _ = mo.render_blocks_oracle(img, boxes, ainv, tinv, fillval=0)
        got = im.crop_blocks(t, _blocks(boxes, ainv, tinv), (67, 74)).cpu().numpy()
        if dtype == np.uint8:
            np.testing.assert_array_equal(got, want)
        else:
            np.testing.assert_allclose(got, want, rtol=0, atol=1e-5 * np.ptp(img))
            if np.allclose(ainv, np.eye(2)) and np.allclose(tinv, np.round(tinv)):
                np.testing.assert_array_equal(got, want)
    # coverage mask + fill value
    ainv, tinv = cases[3]
    want, wmask = mo.render_blocks_oracle(img, boxes, ainv, tinv, fillval=7, cover=(0, 0, 600, 400))
    got, gmask = im.crop_blocks_masked(t, _blocks(boxes, ainv, tinv), (67, 74), fillval=7, cover=(0, 0, 600, 400))
    np.testing.assert_array_equal(gmask.cpu().numpy().astype(bool), wmask)
    if dtype == np.uint8:
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    else:
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=1e-5 * np.ptp(img))


def test_stack_minmax(fc):
    import torch
    from feabas_b200.cuda import image as im
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 33, 47)).astype(np.float32)
    a[2] = 4.0
    mm = im.stack_minmax(torch.from_numpy(a).cuda()).cpu().numpy()
    np.testing.assert_array_equal(mm[:, 0], a.reshape(5, -1).min(1))
    np.testing.assert_array_equal(mm[:, 1], a.reshape(5, -1).max(1))
    u = rng.integers(0, 256, (3, 20, 20), dtype=np.uint8)
    mm = im.stack_minmax(torch.from_numpy(u).cuda()).cpu().numpy()
    np.testing.assert_array_equal(mm[:, 1] - mm[:, 0], np.ptp(u.reshape(3, -1), axis=1))
