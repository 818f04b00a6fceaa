"""CPU tier: the oracle's restatement of the coarse-to-fine loop (``oracle/matcher_oracle.py``) against the traces of
the UNMODIFIED reference (``tests/golden/loop_stitch.npz``, written by ``oracle/make_golden.py`` through
``oracle/ref_harness.py``), the reference-side rebinding ``install()``, and -- where ``/root/reference`` exists --
a live re-run of the reference against the committed traces."""
import json
import warnings

import numpy as np
import pytest

import loop_cases as lc
from conftest import load_golden
from oracle import matcher_oracle as mo
from oracle import ref_loader

_ORACLE_KEYS = ('sigma', 'coarse_downsample', 'fine_downsample', 'spacings', 'residue_len', 'conf_thresh', 'min_num_blocks', 'pad',
                'residue_mode')


@pytest.fixture(scope='module')
def golden_stitch():
    return load_golden('loop_stitch.npz')


@pytest.mark.parametrize('name', ['stitch_yaml_h', 'stitch_yaml_v', 'stitch_yaml_long', 'stitch_levels_half', 'stitch_autopad'])
def test_stitching_oracle_equals_reference_trace(golden_stitch, name):
    spec, rec = lc.stitch_cases()[name], golden_stitch[name]
    a, b = spec['make']()
    np.testing.assert_array_equal(rec['input_sum'], [a.astype(np.float64).sum(), b.astype(np.float64).sum()])
    trace = []
    kw = {k: v for k, v in spec['kwargs'].items() if k in _ORACLE_KEYS}
    xy0, xy1, wt = mo.stitching_oracle(a, b, trace=trace, **kw)
    levels = [t for t in trace if 'spacing' in t]
    ref_levels = [i for i in range(int(rec['trace/n'])) if str(rec[f'trace/{i}/kind']) == 'level']
    assert len(levels) == len(ref_levels)
    for lv, i in zip(levels, ref_levels):
        assert lv['pad'] == bool(rec[f'trace/{i}/pad']) and lv['subpixel'] == bool(rec[f'trace/{i}/subpixel'])
        np.testing.assert_array_equal(lv['xy0'], rec[f'trace/{i}/xy0'])
        np.testing.assert_array_equal(lv['conf'], rec[f'trace/{i}/conf'])
    np.testing.assert_array_equal(xy0, rec['xy0'])
    np.testing.assert_array_equal(xy1, rec['xy1'])
    np.testing.assert_array_equal(wt, rec['weight'])


def test_stitching_oracle_failure(golden_stitch):
    spec, rec = lc.stitch_cases()['stitch_fail'], golden_stitch['stitch_fail']
    a, b = spec['make']()
    out = mo.stitching_oracle(a, b, **{k: v for k, v in spec['kwargs'].items() if k in _ORACLE_KEYS})
    assert bool(rec['failed']) and out[0] is None and out[2] == float(rec['weight_or_conf'])


def test_loop_goldens_cover_the_branches():
    """The recorded cases exercise: single / multi level, auto pad rule (pad dropped on the adjacent level), dwell,
    enlarge, level skipping with weight decay, initial matches, per-block DoG with coverage masks, failures."""
    st, se = load_golden('loop_stitch.npz'), load_golden('loop_section.npz')

    def levels(rec):
        return [dict(pad=bool(rec[f'trace/{i}/pad']), sub=bool(rec[f'trace/{i}/subpixel']), n=rec[f'trace/{i}/conf'].shape[0],
                     sigma=float(rec[f'trace/{i}/sigma']), bs=int(rec[f'trace/{i}/batch_size']))
                for i in range(int(rec['trace/n'])) if str(rec[f'trace/{i}/kind']) == 'level']
    assert [lv['n'] for lv in levels(st['stitch_yaml_long'])] == [4, 108]
    auto = levels(st['stitch_autopad'])
    assert auto[0]['pad'] and not auto[1]['pad'] and auto[1]['sub'] and not auto[0]['sub']
    assert len(levels(se['section_thumb'])) == 4                                   # allow_dwell=1: every spacing twice
    assert [lv['n'] for lv in levels(se['loop_enlarge_skip_decay'])] == [100, 49, 100, 400]   # enlarged once, then down
    assert all(lv['sigma'] == 3.5 and lv['bs'] == 100 for lv in levels(se['section_align']))
    assert bool(st['stitch_fail']['failed']) and bool(se['section_fail']['failed'])
    assert 'phtm' in st['stitch_masks_photometric'] and 'phtm' in st['stitch_photometric_nodog']


@pytest.mark.skipif(not ref_loader.available(), reason='reference sources not present (GPU box)')
def test_install_rebinds_reference_globals():
    """``feabas_b200.cuda.install()`` swaps the three module globals the reference's own
    control flow looks up (feabas/matcher.py:153,213,846; feabas/common.py:353)."""
    import feabas_b200.cuda as fc
    matcher, common, _ = ref_loader.load()
    saved = matcher.xcorr_fft, matcher.global_translation_matcher, common.masked_dog_filter
    try:
        out = fc.install(matcher, common)
        assert out is matcher
        assert matcher.xcorr_fft is fc.xcorr_fft
        assert matcher.global_translation_matcher is fc.global_translation_matcher
        assert common.masked_dog_filter is fc.masked_dog_filter
        # the rebinding is what the reference's callers see: stitching_matcher resolves both names at call time
        assert matcher.stitching_matcher.__globals__['xcorr_fft'] is fc.xcorr_fft
        assert matcher.bboxes_mesh_renderer_matcher.__globals__['xcorr_fft'] is fc.xcorr_fft
    finally:
        matcher.xcorr_fft, matcher.global_translation_matcher, common.masked_dog_filter = saved


@pytest.mark.skipif(not ref_loader.available(), reason='reference sources not present (GPU box)')
def test_reference_live_rerun_matches_committed_trace(golden_stitch):
    """The committed trace is what the unmodified reference produces here and now."""
    from oracle.ref_harness import Harness
    warnings.simplefilter('ignore')
    h = Harness()
    spec, rec = lc.stitch_cases()['stitch_levels_half'], golden_stitch['stitch_levels_half']
    a, b = spec['make']()
    with h:
        xy0, xy1, wt, strain, _ = h.matcher.stitching_matcher(a, b, **json.loads(str(rec['kwargs_json'])))
    np.testing.assert_array_equal(xy0, rec['xy0'])
    np.testing.assert_array_equal(wt, rec['weight'])
    assert strain == float(rec['strain'])


@pytest.mark.skipif(not ref_loader.available(), reason='reference sources not present (GPU box)')
def test_renderer_block_rows_follow_the_reference_renderer():
    """``renderer_block_rows`` reads a REFERENCE ``MeshRenderer`` (the unmodified class; its constructor is the affine
    stand-in's) the way ``crop_field`` does: the rows it builds are the field ``crop_field_affine`` returns, pixel for
    pixel, blocks that hang over the covered region by a square pixel or more send the batch to the host renderer, and a
    loader at another resolution rescales the map as ``crop_multiple`` does."""
    import types
    from oracle.ref_harness import Harness
    from feabas_b200.cuda import matcher as pm
    h = Harness()
    rng = np.random.default_rng(2)
    with h:
        mesh = h.AffineMesh.from_bbox((0, 0, 640, 480), cartesian=True, uid=1.0, resolution=4.0)
        ang = 0.013
        mesh.set_map(np.array([[np.cos(ang), np.sin(ang)], [-np.sin(ang), np.cos(ang)]]) * 1.004, np.array([5.3, -7.1]))
        img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
        for res_loader in (4.0, 8.0):
            loader = h.stream_loader(img, resolution=res_loader)
            render = h.renderer_cls.from_mesh(mesh, image_loader=loader, affine_approx_tol=0.1)
            inner = np.array([(100, 80, 228, 208), (300, 200, 428, 328), (60, 300, 188, 428)], dtype=np.float64)
            for sigma in (0.0, 2.5):
                got = pm.renderer_block_rows(render, inner, sigma, loader.resolution, h.renderer)
                assert got is not None
                rows, shape = got
                assert shape == (128, 128)
                for k, bbox in enumerate(inner):
                    xf, yf, mask = render.crop_field(bbox, log_sigma=sigma)
                    if loader.resolution != render.resolution:
                        sc = render.resolution / loader.resolution
                        xf, yf = (xf + 0.5) * sc - 0.5, (yf + 0.5) * sc - 0.5
                    assert mask.all()
                    cols, rws = np.meshgrid(np.arange(128), np.arange(128))
                    xx, yy = rows[k, 0] + cols * rows[k, 2], rows[k, 1] + rws * rows[k, 3]
                    mine_x = xx * rows[k, 4] + yy * rows[k, 5] + rows[k, 6]
                    mine_y = xx * rows[k, 7] + yy * rows[k, 8] + rows[k, 9]
                    if loader.resolution == render.resolution:          # the same float64 operations in the same order
                        np.testing.assert_array_equal(mine_x, xf)
                        np.testing.assert_array_equal(mine_y, yf)
                    else:                                               # the rescale is folded into the map
                        np.testing.assert_allclose(mine_x, xf, rtol=0, atol=1e-9)
                        np.testing.assert_allclose(mine_y, yf, rtol=0, atol=1e-9)
            # a block that sticks out of the mesh: whole when rendered without the band-pass, host renderer with it
            border = np.array([(100, 80, 228, 208), (560, 380, 688, 508)], dtype=np.float64)
            assert pm.renderer_block_rows(render, border, 0.0, loader.resolution, h.renderer) is not None
            assert pm.renderer_block_rows(render, border, 2.5, loader.resolution, h.renderer) is None
            # ... and mixed block sizes, a zero tolerance or a geodesic mask are not affine-gather batches
            assert pm.renderer_block_rows(render, np.array([(0, 0, 64, 64), (0, 0, 128, 64)], dtype=np.float64), 0.0, loader.resolution, h.renderer) is None
            plain = h.renderer_cls.from_mesh(mesh, image_loader=loader, affine_approx_tol=0.1)
            plain._affine_approx_tol = 0
            assert pm.renderer_block_rows(plain, inner, 0.0, loader.resolution, h.renderer) is None
            geo = types.SimpleNamespace(_affine_approximator=render._affine_approximator, _affine_approx_tol=0.1, _geodesic_mask=True)
            assert pm.renderer_block_rows(geo, inner, 0.0, loader.resolution, h.renderer) is None
        # a renderer with an offset (MeshRenderer keeps vertex coordinates relative to it, renderer.py:57,86): crop_field
        # subtracts it from the block before anything else, and so do the rows
        from oracle import convex
        approx = dict(render._affine_approximator)
        shifted = h.renderer_cls([None], offset=np.array([[37.0, -21.0]]), resolution=4.0, affine_approximator=approx,
                                 affine_approx_tol=0.1, covered_region=convex.box(0.0, 0.0, 639.0, 479.0))
        boxes = inner + np.tile([37.0, -21.0], 2)
        rows, shape = pm.renderer_block_rows(shifted, boxes, 2.5, 4.0, h.renderer)
        for k, bbox in enumerate(boxes):
            xf, yf, mask = shifted.crop_field(bbox, log_sigma=2.5)
            cols, rws = np.meshgrid(np.arange(128), np.arange(128))
            xx, yy = rows[k, 0] + cols * rows[k, 2], rows[k, 1] + rws * rows[k, 3]
            np.testing.assert_array_equal(xx * rows[k, 4] + yy * rows[k, 5] + rows[k, 6], xf)
            np.testing.assert_array_equal(xx * rows[k, 7] + yy * rows[k, 8] + rows[k, 9], yf)
            assert mask.all()


@pytest.mark.skipif(not ref_loader.available(), reason='reference sources not present (GPU box)')
def test_renderer_block_rows_per_block_fit_branch():
    """An ELASTIC mesh: the global affine fit misses the tolerance, so ``crop_field`` fits an affine map per block over the
    triangles that touch it (the reference's own ``bbox_affine_tform`` + ``spatial.fit_affine``, renderer.py:395-416; the
    STRtree is replaced by a brute-force intersection query).  ``renderer_block_rows`` must take the same branch and
    reproduce the reference's field bit for bit where the local fit is within tolerance, and give the batch up where
    it is not."""
    from oracle import convex
    from oracle.ref_harness import Harness
    from feabas_b200.cuda import matcher as pm
    h = Harness()
    # regular triangulation of a 640 x 480 section, vertices every 80 px; moving = initial + a smooth bend
    gx, gy = np.meshgrid(np.arange(0, 641, 80.0) - 0.5, np.arange(0, 481, 80.0) - 0.5)
    v_init = np.stack((gx.ravel(), gy.ravel()), axis=-1)
    nxv = gx.shape[1]
    tris = []
    for r in range(gx.shape[0] - 1):
        for c in range(nxv - 1):
            i = r * nxv + c
            tris += [(i, i + 1, i + nxv), (i + 1, i + nxv + 1, i + nxv)]
    tris = np.array(tris)
    bend = 4e-6 * np.stack(((v_init[:, 1] - 240) ** 2, (v_init[:, 0] - 320) ** 2), axis=-1)     # ~0.3 px over the section, locally affine
    v_mov = v_init * 1.002 + np.array([4.0, -3.0]) + bend

    class Tree:                                            # shapely.STRtree.query(geom, predicate='intersects')
        def __init__(self, polys):
            self.polys = polys

        def query(self, geom, predicate=None):
            assert predicate == 'intersects'
            return np.array([k for k, p in enumerate(self.polys) if p.intersection(geom).area > 0], dtype=np.int64)

    with h:
        spatial = __import__('feabas.spatial', fromlist=['fit_affine'])
        full = spatial.fit_affine(v_init, v_mov)                       # global fit, renderer.py:97-102 (moving -> image)
        resid = np.max(np.sum((v_init - (v_mov @ full[:2, :2] + full[-1, :2])) ** 2, axis=-1)) ** 0.5
        assert resid > 0.12                                            # ... misses a 0.1 px tolerance
        tree = Tree([convex.ConvexPoly(v_mov[t]) for t in tris])
        approx = {'global_affine': full, 'global_residue': resid, 'vertices': (tree, tris, v_mov, v_init)}
        render = h.renderer_cls([None], offset=np.zeros((1, 2)), resolution=4.0, affine_approximator=approx,
                                affine_approx_tol=0.1, covered_region=convex.box(0.0, 0.0, 639.0, 479.0))
        boxes = np.array([(100, 80, 164, 144), (300, 200, 364, 264), (420, 330, 484, 394)], dtype=np.float64)
        got = pm.renderer_block_rows(render, boxes, 2.5, 4.0, h.renderer)
        assert got is not None
        rows, shape = got
        assert shape == (64, 64) and np.ptp(rows[:, 4]) > 0            # one affine map PER BLOCK, not the global one
        for k, bbox in enumerate(boxes):
            xf, yf, mask = render.crop_field(bbox, log_sigma=2.5)      # the reference takes its bbox_affine_tform branch
            cols, rws = np.meshgrid(np.arange(64), np.arange(64))
            xx, yy = rows[k, 0] + cols * rows[k, 2], rows[k, 1] + rws * rows[k, 3]
            np.testing.assert_array_equal(xx * rows[k, 4] + yy * rows[k, 5] + rows[k, 6], xf)
            np.testing.assert_array_equal(xx * rows[k, 7] + yy * rows[k, 8] + rows[k, 9], yf)
            assert mask.all()
        # a tolerance the local fits cannot meet: the batch is the host renderer's
        tight = h.renderer_cls([None], offset=np.zeros((1, 2)), resolution=4.0, affine_approximator=approx,
                               affine_approx_tol=1e-6, covered_region=convex.box(0.0, 0.0, 639.0, 479.0))
        assert pm.renderer_block_rows(tight, boxes, 0.0, 4.0, h.renderer) is None
