"""The oracle (oracle/*.py) against golden vectors produced by the unmodified
reference (tests/golden, written by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import case_kwargs
from oracle import xcorr_oracle as xo
from oracle import matcher_oracle as mo
from feabas_b200 import synth


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    np.testing.assert_array_equal(a, b)          # NaNs compare equal positionally


def test_xcorr_small_bit_exact(golden_small):
    assert len(golden_small) >= 23
    for name, rec in golden_small.items():
        kw = case_kwargs(rec)
        dx, dy, cf = xo.xcorr_oracle(rec['img0'], rec['img1'], **kw)
        _same(dx, rec['dx']), _same(dy, rec['dy']), _same(cf, rec['conf'])


def test_xcorr_seeded_bit_exact(golden_seeded):
    for name, rec in golden_seeded.items():
        size = rec['size'].tolist()
        s0, s1, shifts = synth.block_pairs(int(rec['n']), size, int(rec['seed']), max_shift=int(rec['max_shift']))
        np.testing.assert_allclose([s0.astype(np.float64).sum(), s1.astype(np.float64).sum()], rec['input_sum'], rtol=1e-9)
        dx, dy, cf = xo.xcorr_oracle(s0, s1, **case_kwargs(rec))
        _same(dx, rec['dx']), _same(dy, rec['dy']), _same(cf, rec['conf'])
        # the synthetic ground truth is what the reference reports
        np.testing.assert_array_equal(np.round(dx), shifts[:, 0])
        np.testing.assert_array_equal(np.round(dy), shifts[:, 1])


def test_next_fast_len_matches_fftpack():
    from scipy.fftpack import next_fast_len
    for t in list(range(1, 3000)) + [4095, 4096, 4097, 5999, 8191, 12345]:
        assert xo.next_fast_len_5smooth(t) == next_fast_len(t)
    # the sizes quoted in SURVEY.md §7
    assert [xo.next_fast_len_5smooth(t) for t in (67, 133, 499, 1499, 2999, 559)] == [72, 135, 500, 1500, 3000, 576]


def test_divide_bbox(golden_host):
    i = 0
    while f'divide/{i}/bbox' in golden_host:
        kw = {k.split('/')[-1]: (v if v.ndim else v.item()) for k, v in golden_host.items() if k.startswith(f'divide/{i}/kw/')}
        if 'min_num_blocks' in kw and np.ndim(kw['min_num_blocks']):
            kw['min_num_blocks'] = tuple(kw['min_num_blocks'])
        out = np.stack(mo.divide_bbox_oracle(tuple(golden_host[f'divide/{i}/bbox']), **kw), 0)
        _same(out, golden_host[f'divide/{i}/out'])
        i += 1
    assert i == 5


def test_zorder_and_bbox_helpers(golden_host):
    _same(mo.z_order_oracle(golden_host['zorder/in']), golden_host['zorder/out'])
    _same(mo.bbox_centers_oracle(golden_host['bbox/in']), golden_host['bbox/centers'])
    _same(mo.bbox_sizes_oracle(golden_host['bbox/in']), golden_host['bbox/sizes'])


def test_cartesian_distributor(golden_host):
    i = 0
    while f'cart/{i}/bbox0' in golden_host:
        kw = {k.split('/')[-1]: v.item() for k, v in golden_host.items() if k.startswith(f'cart/{i}/kw/')}
        o0, o1 = mo.cartesian_blocks_oracle(golden_host[f'cart/{i}/bbox0'], golden_host[f'cart/{i}/bbox1'],
                                            golden_host[f'cart/{i}/spacing'].item(), **kw)
        _same(o0, golden_host[f'cart/{i}/out0']), _same(o1, golden_host[f'cart/{i}/out1'])
        i += 1
    assert i == 3


@pytest.mark.parametrize('name', ['plain', 'masked', 'u8', 'stack_masked', 'allmask_true'])
def test_masked_dog(golden_host, name):
    mask = golden_host.get(f'dog/{name}/mask')
    out = mo.masked_dog_oracle(golden_host[f'dog/{name}/img'], golden_host[f'dog/{name}/sigma'].item(), mask=mask)
    _same(out, golden_host[f'dog/{name}/out'])


@pytest.mark.parametrize('name,kw', [('conf', dict(conf_thresh=0.3)), ('retry', dict(conf_thresh=0.9)),
                                     ('retry_df', dict(conf_thresh=0.9, divide_factor=(1, 4))),
                                     ('flat', dict(conf_thresh=0.9))])
def test_global_translation(golden_host, name, kw):
    src = 'retry' if name == 'retry_df' else name
    out = mo.global_translation_oracle(golden_host[f'gt/{src}/img0'], golden_host[f'gt/{src}/img1'], **kw)
    np.testing.assert_array_equal(np.asarray(out, dtype=np.float64), golden_host[f'gt/{name}/out'])


def test_global_translation_cases_cover_both_branches(golden_host):
    # 'conf' returns from the whole-image xcorr, 'retry' must have used a sub-block
    whole = mo.xcorr_oracle(golden_host['gt/retry/img0'][None], golden_host['gt/retry/img1'][None])
    assert golden_host['gt/conf/out'][2] > 0.3
    assert golden_host['gt/retry/out'][2] != pytest.approx(float(whole[2][0]))


def test_reference_direct_when_present():
    """In the build container, also compare against the live reference on fresh random input."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not present (GPU box)')
    matcher, common, const = ref_loader.load()
    rng = np.random.default_rng(99)
    for shp0, shp1, kw in [((3, 45, 61), (3, 45, 61), dict(subpixel=True)),
                           ((2, 33, 20), (2, 50, 64), dict(subpixel=True, pad=False)),
                           ((2, 64, 64), (2, 64, 64), dict(subpixel=True, conf_mode=1))]:
        a = rng.standard_normal(shp0).astype(np.float32)
        b = rng.standard_normal(shp1).astype(np.float32)
        r = matcher.xcorr_fft(a, b, **kw)
        o = xo.xcorr_oracle(a, b, **kw)
        for x, y in zip(r, o):
            _same(x, y)


def test_reference_direct_dog_dtypes():
    """The oracle's band-pass is the live reference's for every dtype branch of common.py:363 (uint8 -> float32, float32,
    float64 kept), with and without masks, signed and unsigned.  (The float64 branch is what ``fb_masked_dog_f64`` is
    compared with on the GPU.)"""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not present (GPU box)')
    matcher, common, const = ref_loader.load()
    rng = np.random.default_rng(17)
    for dtype in (np.uint8, np.float32, np.float64):
        img = (rng.random((2, 61, 83)) * 255).astype(dtype)
        mask = rng.random((2, 61, 83)) > 0.15
        for m in (None, mask):
            for signed in (True, False):
                want = common.masked_dog_filter(img, 2.5, mask=m, signed=signed)
                got = mo.masked_dog_oracle(img, 2.5, mask=m, signed=signed)
                assert got.dtype == want.dtype == (np.float64 if dtype == np.float64 else np.float32)
                np.testing.assert_array_equal(got, want)
