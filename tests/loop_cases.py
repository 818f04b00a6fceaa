"""Seeded inputs and keyword sets of the coarse-to-fine matcher cases (rows a3 - a5 of SURVEY.md section 8).

Shared by ``oracle/make_golden.py`` (which runs the UNMODIFIED reference on them, ``oracle/ref_harness.py``, and
writes ``tests/golden/loop_*.npz``) and by the tests (which run ``feabas_b200.cuda`` on the same inputs and compare
with the recorded traces).  Inputs are regenerated from seeds; ``input_sum`` in the golden file guards against a
drifting generator.  Keyword sets follow the YAML files the reference ships (``configs/default_*_configs.yaml``).
"""
import numpy as np

from feabas_b200 import synth

# configs/default_stitching_configs.yaml:11-20
YAML_STITCH = dict(sigma=2.5, coarse_downsample=0.5, fine_downsample=1.0, pad=True, conf_thresh=0.33, residue_len=2,
                   residue_mode='huber', min_num_blocks=2)
# configs/default_thumbnail_configs.yaml:43-53 as thumbnail.py:494-509 hands them on (sigma popped, allow_dwell=1); the region
# distributors belong to the geometry layer, the cartesian one is the reference loop's own
YAML_THUMB = dict(conf_thresh=0.35, pad=True, spacings=[150, 50], shrink_factor=1, distributor='cartesian_bbox', residue_mode='huber',
                  residue_len=3, allow_dwell=1)
# configs/default_alignment_configs.yaml:14-28 (matcher_config), spacings scaled to the test images
YAML_ALIGN = dict(batch_size=100, conf_thresh=0.35, sigma=3.5, pad=True, distributor='cartesian_bbox', spacings=[200, 100],
                  shrink_factor=0.7, residue_mode='huber', residue_len=3, render_weight_threshold=0.1,
                  stiffness_multiplier_threshold=0.1)


def strips(seed, shape=(700, 260), jitter=9, noise=6.0):
    """Two uint8 strips of one canvas, the second cut at an integer offset (+ independent noise)."""
    h, w = shape
    canvas = synth.em_canvas(h + 40, 2 * w + 40, seed=seed)
    rng = np.random.default_rng(seed)
    jx, jy = rng.integers(-jitter, jitter + 1, 2)
    a = canvas[20:20 + h, 20:20 + w]
    b = canvas[20 + jy:20 + jy + h, 20 + jx:20 + jx + w]
    nz = rng.normal(0, noise, (2, h, w))
    a = np.clip(a + nz[0], 0, 255).astype(np.uint8)
    b = np.clip(b + nz[1], 0, 255).astype(np.uint8)
    return a, b, (-int(jx), -int(jy))


def _warp_nearest_free(canvas, a, t, shape):
    """Sample ``canvas`` at ``p @ a + t`` (bilinear, float64 arithmetic, numpy only: deterministic everywhere)."""
    h, w = shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = xx * a[0, 0] + yy * a[1, 0] + t[0]
    sy = xx * a[0, 1] + yy * a[1, 1] + t[1]
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    fx, fy = sx - x0, sy - y0
    c = canvas.astype(np.float64)
    v = (c[y0, x0] * (1 - fx) * (1 - fy) + c[y0, x0 + 1] * fx * (1 - fy) + c[y0 + 1, x0] * (1 - fx) * fy + c[y0 + 1, x0 + 1] * fx * fy)
    return v


def section_pair(seed, size=600, angle=0.012, scale=1.008, shift=(7.0, -5.0), noise=5.0):
    """Two uint8 "sections": the second is the first seen through a small similarity transform + noise."""
    margin = 60
    canvas = synth.em_canvas(size + 2 * margin, size + 2 * margin, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    img0 = canvas[margin:margin + size, margin:margin + size]
    ca, sa = np.cos(angle) * scale, np.sin(angle) * scale
    a = np.array([[ca, sa], [-sa, ca]])
    ctr = np.array([size / 2, size / 2])
    t = ctr - ctr @ a + np.array(shift) + margin
    img1 = _warp_nearest_free(canvas, a, t, (size, size))
    nz = rng.normal(0, noise, (2, size, size))
    img0 = np.clip(img0 + nz[0], 0, 255).astype(np.uint8)
    img1 = np.clip(np.round(img1 + nz[1]), 0, 255).astype(np.uint8)
    return img0, img1


def _edge_mask(shape, left=0, top=0, right=0, bottom=0):
    m = np.ones(shape, dtype=bool)
    if left:
        m[:, :left] = False
    if top:
        m[:top, :] = False
    if right:
        m[:, -right:] = False
    if bottom:
        m[-bottom:, :] = False
    return m


# name -> dict(kind, inputs=callable returning the positional arrays, kwargs)
def stitch_cases():
    cases = {}
    cases['stitch_yaml_h'] = dict(make=lambda: strips(21)[:2], kwargs=dict(YAML_STITCH))                       # single level (75)
    cases['stitch_yaml_v'] = dict(make=lambda: tuple(np.ascontiguousarray(x.T) for x in strips(22, (900, 240))[:2]),
                                  kwargs=dict(YAML_STITCH))
    cases['stitch_yaml_long'] = dict(make=lambda: strips(23, (2400, 200), jitter=14)[:2], kwargs=dict(YAML_STITCH))   # two levels: 600 / 66.7
    cases['stitch_levels_half'] = dict(make=lambda: strips(24)[:2],
                                       kwargs=dict(sigma=2.0, coarse_downsample=0.5, fine_downsample=0.5, spacings=[60, 200], pad=True))
    cases['stitch_autopad'] = dict(make=lambda: strips(25, (800, 300), jitter=5)[:2],
                                   kwargs=dict(sigma=2.5, coarse_downsample=1, fine_downsample=1, spacings=[50, 200], conf_thresh=0.3))
    cases['stitch_relative_spacings'] = dict(make=lambda: strips(26)[:2],
                                             kwargs=dict(sigma=2.5, coarse_downsample=0.5, fine_downsample=1, spacings=[0.1, 0.4], pad=True,
                                                         residue_mode='threshold', residue_len=1.5))

    # factors that are not 1/k: OpenCV's general INTER_AREA path (coverage tables), odd output sizes, sigma * factor
    cases['stitch_frac_downsample'] = dict(make=lambda: strips(29, (810, 275), jitter=7)[:2],
                                           kwargs=dict(sigma=2.5, coarse_downsample=0.4, fine_downsample=0.8, pad=True, conf_thresh=0.33,
                                                       residue_len=2))

    def masked():
        a, b, _ = strips(27)
        return a, b
    cases['stitch_masks_photometric'] = dict(make=masked, kwargs=dict(YAML_STITCH, compute_photometric=True),
                                             masks=lambda a, b: (_edge_mask(a.shape, left=30, top=12), _edge_mask(b.shape, right=40)))
    cases['stitch_photometric_nodog'] = dict(make=lambda: tuple(synth.dog_f32(x) for x in strips(28)[:2]),
                                             kwargs=dict(sigma=0, compute_photometric=True, pad=True))

    def noise():
        rng = np.random.default_rng(0)
        return rng.integers(0, 256, (300, 200), dtype=np.uint8), rng.integers(0, 256, (300, 200), dtype=np.uint8)
    cases['stitch_fail'] = dict(make=noise, kwargs=dict(conf_thresh=0.33, coarse_downsample=0.5))
    return cases


def section_cases():
    """``section_matcher`` / ``iterative_xcorr_matcher_w_mesh`` on AffineMesh.from_bbox meshes over StreamLoaders.

    ``prep``: 'dog' -> the loaders hold float32 band-passed images (thumbnail.py:503-507, sigma=0 inside), 'raw' -> uint8
    images and per-block DoG (aligner.py:129, sigma from the YAML)."""
    cases = {}
    cases['section_thumb'] = dict(make=lambda: section_pair(31), prep='dog', dog_sigma=3.5, entry='section', kwargs=dict(YAML_THUMB, sigma=0.0))
    cases['section_align'] = dict(make=lambda: section_pair(32, size=700, angle=0.006, scale=1.004, shift=(4.0, 6.0)), prep='raw',
                                  entry='section', kwargs=dict(YAML_ALIGN))
    cases['section_align_border'] = dict(make=lambda: section_pair(33, size=520, angle=0.0, scale=1.0, shift=(11.0, -9.0)), prep='raw',
                                         entry='section', kwargs=dict(YAML_ALIGN, spacings=[130], shrink_factor=1))
    cases['loop_enlarge_skip_decay'] = dict(make=lambda: section_pair(34, size=640, angle=0.004, scale=1.0, shift=(-21.0, 13.0)), prep='dog',
                                            dog_sigma=2.5, entry='loop',
                                            kwargs=dict(spacings=[64, 32], conf_thresh=0.3, allow_enlarge=True, max_spacing_skip=1,
                                                        link_weight_decay=0.5, compute_strain=True, residue_len=2, residue_mode='huber',
                                                        batch_size=40))
    cases['loop_initial_matches'] = dict(make=lambda: section_pair(35, size=600, angle=0.02, scale=1.0, shift=(30.0, -24.0)), prep='dog',
                                         dog_sigma=3.0, entry='section', initial=True,
                                         kwargs=dict(YAML_THUMB, sigma=0.0, spacings=[100, 50], compute_strain=True))
    cases['section_fail'] = dict(make=lambda: (np.random.default_rng(5).integers(0, 256, (300, 300), dtype=np.uint8),
                                               np.random.default_rng(6).integers(0, 256, (300, 300), dtype=np.uint8)),
                                 prep='raw', entry='section', kwargs=dict(YAML_ALIGN, spacings=[100]))
    return cases


def initial_matches_for(size, angle, shift, n=40, seed=9):
    """Coarse point matches (what the thumbnail stage hands to the fine matcher): xy1 = similarity(xy0) + 1 px noise."""
    rng = np.random.default_rng(seed)
    xy1 = rng.uniform(40, size - 40, (n, 2))
    ca, sa = np.cos(angle), np.sin(angle)
    a = np.array([[ca, sa], [-sa, ca]])
    ctr = np.array([size / 2, size / 2])
    xy0 = xy1 @ a + (ctr - ctr @ a + np.array(shift)) + rng.normal(0, 1.0, (n, 2))
    return xy0, xy1, np.ones(n)
