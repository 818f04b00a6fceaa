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
