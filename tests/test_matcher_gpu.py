"""GPU tier: matcher layers above xcorr_fft (global_translation_matcher, bboxes_mesh_renderer_matcher,
iterative loop, stitching_matcher) against golden vectors of the unmodified reference and the oracle."""
import numpy as np
import pytest

import loop_cases as lc
from feabas_b200 import synth
from oracle import matcher_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fc():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import feabas_b200.cuda as fc
    return fc


@pytest.mark.parametrize('name,kw', [('conf', dict(conf_thresh=0.3)), ('retry', dict(conf_thresh=0.9)),
                                     ('retry_df', dict(conf_thresh=0.9, divide_factor=(1, 4))),
                                     ('flat', dict(conf_thresh=0.9))])
def test_global_translation_golden(fc, golden_host, name, kw):
    src = 'retry' if name == 'retry_df' else name
    tx, ty, cf = fc.global_translation_matcher(golden_host[f'gt/{src}/img0'], golden_host[f'gt/{src}/img1'], **kw)
    want = golden_host[f'gt/{name}/out']
    assert isinstance(tx, float) and isinstance(ty, float) and isinstance(cf, float)
    assert (tx, ty) == (want[0], want[1])
    assert cf == pytest.approx(want[2], rel=1e-4, abs=1e-6)


def test_global_translation_sigma_and_shapes(fc):
    tiles, _, _ = synth.tile_grid(1, 2, tile_hw=(300, 400), overlap=0.25, jitter=6, seed=3)
    a, b = tiles[0][:, -100:], tiles[1][:, :110]                      # different widths
    want = mo.global_translation_oracle(mo.masked_dog_oracle(a, 2.0), mo.masked_dog_oracle(b, 2.0))
    got = fc.global_translation_matcher(a, b, sigma=2.0)
    assert got[:2] == want[:2]
    assert got[2] == pytest.approx(want[2], rel=1e-4, abs=1e-6)


def _mesh(fc, shape, shift=(0, 0), uid=0):
    m = fc.AffineMesh((0, 0, shape[1], shape[0]), uid=uid)
    m.apply_translation(shift, 0)
    return m


@pytest.mark.parametrize('batch_size,pad,subpixel', [(None, True, True), (7, True, False), (50, False, True)])
def test_block_grid_pass_translation(fc, batch_size, pad, subpixel):
    """bboxes_mesh_renderer_matcher with translation-only meshes == the oracle's integer-shift crops."""
    canvas = synth.dog_f32(synth.em_canvas(420, 900, seed=4))
    img0, img1 = np.ascontiguousarray(canvas[5:405, 10:810]), np.ascontiguousarray(canvas[0:400, 40:840])
    shift0 = (-30, 5)                                                 # img0 placed in img1's frame
    m0, m1 = _mesh(fc, img0.shape, shift0, 0), _mesh(fc, img1.shape, (0, 0), 1)
    b0, b1 = fc.distributor_cartesian_bbox(m0, m1, 75, min_num_blocks=2, zorder=True)
    big0, big1 = fc.distributor_cartesian_bbox(m0, m1, 300, zorder=True)
    b0, b1 = np.concatenate((big0, b0)), np.concatenate((big1, b1))   # two block sizes -> at least two batches
    got = fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(img0), fc.ArrayLoader(img1), b0, b1,
                                          batch_size=batch_size, pad=pad, subpixel=subpixel)
    want = mo.block_grid_match_oracle(img0, img1, b0, b1, shift0=shift0, shift1=(0, 0), batch_size=batch_size,
                                      pad=pad, subpixel=subpixel)
    assert got[0].shape == want[0].shape == (b0.shape[0], 2)
    agree = np.all(np.abs(got[0] - want[0]) < 0.02, axis=1) & np.all(np.abs(got[1] - want[1]) < 0.02, axis=1)
    # blocks that hang over the image border correlate weakly and may tie; everything confident must agree
    strong = want[2] > 0.3
    assert np.all(agree[strong]) and strong.sum() > 0.6 * strong.size
    np.testing.assert_allclose(got[2][agree], want[2][agree], rtol=1e-4, atol=1e-5)
    assert got[2].dtype == np.float32


def test_block_grid_pass_empty(fc):
    m0, m1 = _mesh(fc, (50, 50)), _mesh(fc, (50, 50), uid=1)
    img = np.zeros((50, 50), np.float32)
    out = fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(img), fc.ArrayLoader(img), np.empty((0, 4)), np.empty((0, 4)))
    assert out[0].shape == (0, 2) and out[1].shape == (0, 2) and out[2].shape == (0,)


def _strips(seed, shape=(700, 260), jitter=9, overlap_px=200):
    """Two uint8 strips of a horizontal overlap (+ margin), true offset known."""
    h, w = shape
    canvas = synth.em_canvas(h + 40, 2 * w + 40, seed=seed)
    rng = np.random.default_rng(seed)
    jx, jy = rng.integers(-jitter, jitter + 1, 2)
    a = canvas[20:20 + h, 20:20 + w]
    b = canvas[20 + jy:20 + jy + h, 20 + jx:20 + jx + w]
    noise = rng.normal(0, 6, (2, h, w))
    a = np.clip(a + noise[0], 0, 255).astype(np.uint8)
    b = np.clip(b + noise[1], 0, 255).astype(np.uint8)
    return a, b, (-jx, -jy)


@pytest.mark.parametrize('cfg', [dict(sigma=2.5, coarse_downsample=0.5, fine_downsample=1.0, pad=True, conf_thresh=0.33, residue_len=2),
                                 dict(sigma=2.5, coarse_downsample=1, fine_downsample=1, conf_thresh=0.3),
                                 dict(sigma=2.0, coarse_downsample=0.5, fine_downsample=0.5, spacings=[60, 200], pad=True)])
def test_stitching_matcher_vs_oracle(fc, cfg):
    a, b, true = _strips(seed=21)
    trace = []
    want = mo.stitching_oracle(a, b, trace=trace, **cfg)
    got = fc.stitching_matcher(a, b, **cfg)
    assert len(got) == 5 and got[4] is None
    assert want[0] is not None and got[0] is not None
    assert got[0].shape == want[0].shape and got[1].shape == want[1].shape
    # the recovered displacement field is the true offset
    d = got[1] - got[0]
    assert np.all(np.abs(np.median(d, axis=0) - np.array(true)) < 0.5)
    # same blocks, same matches: sub-pixel agreement of every point pair, weights within 1e-3
    np.testing.assert_allclose(got[0], want[0], atol=0.02)
    np.testing.assert_allclose(got[1], want[1], atol=0.02)
    np.testing.assert_allclose(got[2], want[2], rtol=1e-4, atol=1e-6)
    assert isinstance(got[3], float) or np.isscalar(got[3])


def test_stitching_matcher_failure_tuple(fc):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (300, 200), dtype=np.uint8)
    b = rng.integers(0, 256, (300, 200), dtype=np.uint8)
    out = fc.stitching_matcher(a, b, conf_thresh=0.33, coarse_downsample=0.5)
    assert out == (None, None, 0.33, None, None)


def test_section_matcher_surrogate(fc):
    """section_matcher on thumbnail-like float32 (already band-passed) images, sigma=0, as thumbnail.py:509 calls it."""
    canvas = synth.dog_f32(synth.em_canvas(700, 700, seed=8), 3.5)
    img0, img1 = np.ascontiguousarray(canvas[20:620, 30:630]), np.ascontiguousarray(canvas[26:626, 22:622])
    m0 = fc.AffineMesh((0, 0, 600, 600), uid=0)
    m1 = fc.AffineMesh((0, 0, 600, 600), uid=1)
    xy0, xy1, wt, strain = fc.section_matcher(m0, m1, fc.ArrayLoader(img0), fc.ArrayLoader(img1), sigma=0.0,
                                              spacings=[150, 50], conf_thresh=0.35, pad=True, distributor='cartesian_bbox',
                                              residue_mode='huber', residue_len=3)
    assert xy0.shape == xy1.shape and xy0.shape[0] > 50 and wt.shape[0] == xy0.shape[0]
    d = np.median(xy1 - xy0, axis=0)
    np.testing.assert_allclose(d, [8, -6], atol=0.3)


class _DuckRenderer:
    """The attributes of ``feabas.renderer.MeshRenderer`` that ``renderer_block_rows`` reads (renderer.py:36-48), filled
    in for an affine mesh the way the reference's ``from_mesh`` does (:90-109), and a ``crop_multiple`` that renders on
    the host through the oracle (what the reference's own method does for such a mesh)."""

    def __init__(self, mesh, img, tol=0.1, origin=(0, 0)):
        from oracle import convex
        ainv, tinv = mesh.render_map()
        self._affine_approximator = {'global_affine': np.concatenate((ainv, tinv.reshape(1, 2)), axis=0), 'global_residue': 0.0}
        self._affine_approx_tol = tol
        self._offset = np.zeros((1, 2))
        self._geodesic_mask = False
        self.resolution = mesh.resolution
        x0, y0, x1, y1 = mesh.bounds
        self._cover = (x0 + 0.5, y0 + 0.5, x1 - 0.5, y1 - 0.5)
        self._covered_region = convex.box(*self._cover)
        self._mesh, self._img, self._origin = mesh, img, origin
        self.host_batches = 0

    def crop_multiple(self, bboxes, **kwargs):
        self.host_batches += 1
        sigma = kwargs.get('log_sigma', 0)
        ainv, tinv = self._mesh.render_map()
        cover = None if not sigma > 0 else tuple(np.array(self._cover) - np.tile(self._origin, 2))
        stack, mask = mo.render_blocks_oracle(self._img, bboxes, ainv, tinv, fillval=0, origin_xy=self._origin, cover=cover)
        if not mask.any():
            return None
        return mo.masked_dog_oracle(stack, sigma, mask=mask) if sigma > 0 else stack


class _RegionLoader:
    """A loader without an in-RAM image attribute (FEABAS's tile-backed loaders): only ``crop`` gives pixels."""

    def __init__(self, img, origin=(0, 0), resolution=4.0):
        self.img, self.origin, self.resolution, self.default_fillval, self.dtype = img, origin, resolution, 0, img.dtype
        self.crops = 0

    def crop(self, bbox, return_empty=False, **kwargs):
        self.crops += 1
        x0, y0 = self.origin
        return mo.crop_with_fill(self.img, (bbox[0] - x0, bbox[1] - y0, bbox[2] - x0, bbox[3] - y0), self.default_fillval)


def test_bboxes_renderer_matcher_device_and_host_batches(fc):
    """Reference-style renderers: batches that are affine within tolerance and fully covered are cut on the device (same
    numbers as the AffineMesh path, bit for bit), a batch with a block hanging over the mesh border goes through the
    renderer's own ``crop_multiple`` on the host and still meets the gates."""
    import types
    from oracle import convex
    from feabas_b200.cuda import matcher as pm
    mod = types.SimpleNamespace(shpgeo=types.SimpleNamespace(box=convex.box),
                                shapely=types.SimpleNamespace(affinity=types.SimpleNamespace(affine_transform=convex.affine_transform)))
    img0, img1 = lc.section_pair(71, size=520, angle=0.006, scale=1.003, shift=(5.0, -3.0))
    h, w = img0.shape
    m0 = fc.AffineMesh.from_bbox((0, 0, w, h), cartesian=True, uid=0.0)
    m1 = fc.AffineMesh.from_bbox((0, 0, w, h), cartesian=True, uid=1.0)
    ang = 0.006
    m1.set_map(np.array([[np.cos(ang), np.sin(ang)], [-np.sin(ang), np.cos(ang)]]) * 1.003, np.array([-4.0, 2.5]))
    inner = np.array([(x, y, x + 128, y + 128) for y in (40, 190, 340) for x in (40, 190, 340)], dtype=np.float64)
    for sigma in (0.0, 2.5):
        kw = dict(sigma=sigma, batch_size=9, pad=True, subpixel=True)
        want = fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(img0), fc.ArrayLoader(img1), inner, inner, **kw)
        r0, r1 = _DuckRenderer(m0, img0), _DuckRenderer(m1, img1)
        ld0, ld1 = _RegionLoader(img0), _RegionLoader(img1)
        got = pm.bboxes_renderer_matcher(r0, r1, ld0, ld1, inner, inner, renderer_module=mod, **kw)
        assert r0.host_batches == 0 and r1.host_batches == 0 and ld0.crops == 1 and ld1.crops == 1
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    # a block over the border with the band-pass on: host batch
    border = np.concatenate((inner[:4], np.array([(430, 430, 558, 558)], dtype=np.float64)), axis=0)
    kw = dict(sigma=2.5, batch_size=9, pad=True, subpixel=True)
    want = fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(img0), fc.ArrayLoader(img1), border, border, **kw)
    r0, r1 = _DuckRenderer(m0, img0), _DuckRenderer(m1, img1)
    got = pm.bboxes_renderer_matcher(r0, r1, _RegionLoader(img0), _RegionLoader(img1), border, border, renderer_module=mod, **kw)
    assert r0.host_batches == 1 and r1.host_batches == 1
    np.testing.assert_allclose(got[0], want[0], atol=0.02)
    np.testing.assert_allclose(got[1], want[1], atol=0.02)
    np.testing.assert_allclose(got[2], want[2], rtol=1e-4, atol=1e-6)
