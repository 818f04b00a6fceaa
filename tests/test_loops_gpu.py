"""GPU tier: the coarse-to-fine matcher layers against traces of the UNMODIFIED
reference.

``tests/golden/loop_*.npz`` were written by ``oracle/make_golden.py``: the reference's own ``stitching_matcher``,
``section_matcher``, ``iterative_xcorr_matcher_w_mesh``, ``bboxes_mesh_renderer_matcher`` and
``MeshRenderer.crop_multiple`` executed in the build container with only the geometry layer (Mesh / SLM /
MeshRenderer.from_mesh) replaced by the affine stand-ins (``oracle/ref_harness.py``).  Every pyramid level the
reference went through is recorded: block lists, pad / sub-pixel flags, batch size, the mesh maps the blocks were
rendered through, the matches and confidences; plus the final ``(xy0, xy1, weight, strain)``.

Gates (BASELINE.json north_star): same levels, same flags, bit-equal block lists; matched points within 0.02 px;
confidences within 1e-4 relative (1e-6 absolute floor).
"""
import json

import numpy as np
import pytest

import loop_cases as lc
from conftest import load_golden
from oracle import matcher_oracle as mo

pytestmark = pytest.mark.gpu

PX_TOL = 0.02
CONF_RTOL = 1e-4
CONF_ATOL = 1e-6


@pytest.fixture(scope='module')
def fc():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import feabas_b200.cuda as fc
    return fc


@pytest.fixture(scope='module')
def golden_stitch():
    return load_golden('loop_stitch.npz')


@pytest.fixture(scope='module')
def golden_section():
    return load_golden('loop_section.npz')


def _levels(rec):
    out = []
    for i in range(int(rec['trace/n'])):
        if str(rec[f'trace/{i}/kind']) == 'level':
            out.append({k.split('/', 2)[2]: v for k, v in rec.items() if k.startswith(f'trace/{i}/')})
    return out


class LevelRecorder:
    """Records every call of the product's ``bboxes_mesh_renderer_matcher`` made by its coarse-to-fine loop."""

    def __init__(self, monkeypatch, fc):
        import feabas_b200.cuda.matcher as pm
        self.levels = []
        inner = pm.bboxes_mesh_renderer_matcher

        def logged(mesh0, mesh1, loader0, loader1, bboxes0, bboxes1, **kwargs):
            out = inner(mesh0, mesh1, loader0, loader1, bboxes0, bboxes1, **kwargs)
            self.levels.append(dict(bboxes0=np.array(bboxes0, dtype=np.float64), bboxes1=np.array(bboxes1, dtype=np.float64),
                                    pad=bool(kwargs.get('pad', True)), subpixel=bool(kwargs.get('subpixel', False)),
                                    batch_size=-1 if kwargs.get('batch_size', None) is None else int(kwargs['batch_size']),
                                    sigma=float(kwargs.get('sigma', 0.0)),
                                    map0=np.concatenate([np.ravel(v) for v in mesh0.get_map()]),
                                    map1=np.concatenate([np.ravel(v) for v in mesh1.get_map()]),
                                    xy0=np.array(out[0]), xy1=np.array(out[1]), conf=np.array(out[2])))
            return out
        monkeypatch.setattr(pm, 'bboxes_mesh_renderer_matcher', logged)


def _check_levels(got, want, conf_thresh):
    assert len(got) == len(want), f'{len(got)} levels, the reference went through {len(want)}'
    for i, (g, w) in enumerate(zip(got, want)):
        assert g['pad'] == bool(w['pad']) and g['subpixel'] == bool(w['subpixel']), f'level {i}: pad / sub-pixel flags differ'
        assert g['batch_size'] == int(w['batch_size']) and g['sigma'] == float(w['sigma'])
        np.testing.assert_array_equal(g['bboxes0'], w['bboxes0'], err_msg=f'level {i}: block list differs')
        np.testing.assert_array_equal(g['bboxes1'], w['bboxes1'])
        np.testing.assert_allclose(g['map0'], w['map0'], rtol=0, atol=1e-4)       # maps carry the sub-pixel noise of the level before
        np.testing.assert_allclose(g['map1'], w['map1'], rtol=0, atol=1e-4)
        assert g['conf'].dtype == w['conf'].dtype and g['xy0'].dtype == np.float64
        _check_matches(g['xy0'], g['xy1'], g['conf'], w['xy0'], w['xy1'], w['conf'], f'level {i}')


def _check_matches(xy0, xy1, wt, rxy0, rxy1, rwt, what, wt_rtol=CONF_RTOL):
    assert xy0.shape == rxy0.shape and xy1.shape == rxy1.shape and wt.shape == rwt.shape, what
    err = max(np.abs(xy0 - rxy0).max(initial=0), np.abs(xy1 - rxy1).max(initial=0))
    assert err <= PX_TOL, f'{what}: matched points differ by {err:.4f} px'
    werr = np.abs(wt.astype(np.float64) - rwt.astype(np.float64))
    lim = wt_rtol * np.abs(rwt.astype(np.float64)) + CONF_ATOL
    assert np.all(werr <= lim), f'{what}: confidence / weight differs by {werr.max():.3e} (rel {np.max(werr / np.maximum(np.abs(rwt), 1e-30)):.3e})'


def _kwargs(rec):
    return json.loads(str(rec['kwargs_json']))


@pytest.mark.parametrize('name', list(lc.stitch_cases()))
def test_stitching_matcher_reference_trace(fc, golden_stitch, monkeypatch, name):
    spec, rec = lc.stitch_cases()[name], golden_stitch[name]
    img0, img1 = spec['make']()
    np.testing.assert_array_equal(rec['input_sum'], [np.asarray(img0, dtype=np.float64).sum(), np.asarray(img1, dtype=np.float64).sum()])
    kwargs = _kwargs(rec)
    assert kwargs == json.loads(json.dumps(spec['kwargs']))
    if 'masks' in spec:
        kwargs['mask0'], kwargs['mask1'] = spec['masks'](img0, img1)
    recorder = LevelRecorder(monkeypatch, fc)
    xy0, xy1, weight, strain, phtm = fc.stitching_matcher(img0, img1, **kwargs)
    if bool(rec['failed']):
        assert xy0 is None and xy1 is None and strain is None and phtm is None
        assert weight == float(rec['weight_or_conf'])
        return
    _check_levels(recorder.levels, _levels(rec), kwargs.get('conf_thresh', 0.3))
    _check_matches(xy0, xy1, weight, rec['xy0'], rec['xy1'], rec['weight'], 'final')
    assert strain == pytest.approx(float(rec['strain']), rel=2e-2, abs=1e-7)
    if 'phtm' in rec:
        np.testing.assert_allclose(np.asarray(phtm, dtype=np.float64), rec['phtm'], rtol=1e-5)
    else:
        assert phtm is None


def _section_inputs(fc, spec):
    img0, img1 = spec['make']()
    sums = [np.asarray(img0, dtype=np.float64).sum(), np.asarray(img1, dtype=np.float64).sum()]
    if spec['prep'] == 'dog':
        img0, img1 = mo.masked_dog_oracle(img0, spec['dog_sigma']), mo.masked_dog_oracle(img1, spec['dog_sigma'])
    (h0, w0), (h1, w1) = img0.shape, img1.shape
    mesh0 = fc.AffineMesh.from_bbox((0, 0, w0, h0), cartesian=True, uid=0.0, resolution=4.0)
    mesh1 = fc.AffineMesh.from_bbox((0, 0, w1, h1), cartesian=True, uid=1.0, resolution=4.0)
    return sums, mesh0, mesh1, fc.ArrayLoader(np.ascontiguousarray(img0), resolution=4.0), fc.ArrayLoader(np.ascontiguousarray(img1), resolution=4.0)


@pytest.mark.parametrize('name', list(lc.section_cases()))
def test_section_matcher_reference_trace(fc, golden_section, monkeypatch, name):
    spec, rec = lc.section_cases()[name], golden_section[name]
    sums, mesh0, mesh1, ld0, ld1 = _section_inputs(fc, spec)
    np.testing.assert_array_equal(rec['input_sum'], sums)
    kwargs = _kwargs(rec)
    if spec.get('initial'):
        p0, p1, w = lc.initial_matches_for(600, 0.02, (30.0, -24.0))
        kwargs['initial_matches'] = fc.Match(p0, p1, w)
    recorder = LevelRecorder(monkeypatch, fc)
    if spec['entry'] == 'section':
        xy0, xy1, weight, strain = fc.section_matcher(mesh0, mesh1, ld0, ld1, **kwargs)
    else:
        spacings = kwargs.pop('spacings')
        xy0, xy1, weight, strain = fc.iterative_xcorr_matcher_w_mesh(mesh0, mesh1, ld0, ld1, spacings, **kwargs)
    _check_levels(recorder.levels, _levels(rec), kwargs.get('conf_thresh', 0.3))
    if bool(rec['failed']):
        assert xy0 is None and xy1 is None and weight == 0
    else:
        _check_matches(xy0, xy1, weight, rec['xy0'], rec['xy1'], rec['weight'], 'final')
    assert strain == pytest.approx(float(rec['strain']), rel=2e-2, abs=1e-7)
