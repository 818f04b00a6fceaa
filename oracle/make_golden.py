"""Generate ``tests/golden/*.npz`` from the UNMODIFIED reference.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Run in the build container
(where ``/root/reference`` exists):

    python -m oracle.make_golden

Every record stores the inputs (small arrays) or the seed that regenerates
them, the keyword arguments, and the reference's outputs.  The reference has
no tests or golden vectors of its own; these are outputs of its
functions run here with scipy 1.18.1 / numpy 2.3.5 / OpenCV 4.13.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader                     # noqa: E402
from feabas_b200 import synth                     # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def xcorr_cases():
    """name -> (img0, img1, kwargs).  Inputs are kept small enough to commit."""
    rng = np.random.default_rng(20261017)
    cases = {}
    a = rng.standard_normal((2, 128, 128)).astype(np.float32)
    cases['roll_f32_sub'] = (a, np.roll(a, (3, -5), axis=(1, 2)), dict(subpixel=True))
    z = np.zeros((2, 32, 32), np.float32)
    cases['zeros_mirror'] = (z, z, dict(conf_mode=2))
    cases['zeros_none'] = (z, z, dict(conf_mode=0))
    big = rng.standard_normal((1, 64, 50)).astype(np.float32)
    small = big[:, 14:54, 12:42].copy() + 0.05 * rng.standard_normal((1, 40, 30)).astype(np.float32)
    cases['diffshape_pad'] = (small, big, dict(subpixel=True, pad=True))
    cases['diffshape_nopad'] = (small, big, dict(subpixel=True, pad=False))
    cases['diffshape_swapped'] = (big, small, dict(subpixel=True, pad=True))
    s0, s1, _ = synth.block_pairs(4, (74, 67), seed=11, max_shift=9)
    cases['stitch_fine_pad'] = (s0, s1, dict(subpixel=True, pad=True))          # FFT 150 x 135
    cases['stitch_fine_nopad'] = (s0, s1, dict(subpixel=True, pad=False))       # FFT 75 x 72
    s0, s1, _ = synth.block_pairs(3, (60, 75), seed=12, max_shift=7)
    cases['stitch_v_pad'] = (s0, s1, dict(subpixel=True, pad=True))             # FFT 120 x 150
    s0, s1, _ = synth.block_pairs(3, 50, seed=13, max_shift=6)
    cases['thumb50_pad_nosub'] = (s0, s1, dict(subpixel=False, pad=True))       # FFT 100 x 100
    cases['thumb50_std'] = (s0, s1, dict(subpixel=True, pad=True, conf_mode=1))
    cases['thumb50_none'] = (s0, s1, dict(subpixel=True, pad=True, conf_mode=0))
    s0, s1, _ = synth.block_pairs(2, 128, seed=14, max_shift=20)
    cases['pow2_128_pad'] = (s0, s1, dict(subpixel=True, pad=True))             # FFT 256 x 256
    cases['pow2_128_nopad'] = (s0, s1, dict(subpixel=True, pad=False))          # FFT 128 x 128
    u0, u1, _ = synth.block_pairs(2, 64, seed=15, max_shift=10, dtype=np.float32, band_pass=False)
    cases['uint8_in'] = (u0.clip(0, 255).astype(np.uint8), u1.clip(0, 255).astype(np.uint8),
                         dict(subpixel=True, pad=True))
    cases['float64_in'] = (u0.astype(np.float64), u1.astype(np.float64), dict(subpixel=True, pad=True))
    c0 = rng.standard_normal((2, 48, 40, 3)).astype(np.float32)
    c1 = np.roll(c0, (-4, 6), axis=(1, 2)) + 0.1 * rng.standard_normal(c0.shape).astype(np.float32)
    cases['multichannel'] = (c0, c1, dict(subpixel=True, pad=True))
    s0, s1, _ = synth.block_pairs(2, 64, seed=16, max_shift=8)
    m0 = np.ones((64, 64), np.float32)
    m0[:, :20] = 0
    m1 = np.ones((64, 64), np.float32)
    m1[40:, :] = 0
    cases['normalize_masks'] = (s0, s1, dict(subpixel=True, pad=True, normalize=True, mask0=m0, mask1=m1))
    cases['normalize_default'] = (s0, s1, dict(subpixel=True, pad=True, normalize=True))
    # uncorrelated pair: low confidence, exercises clip and det<=0 paths
    n0 = rng.standard_normal((3, 40, 40)).astype(np.float32)
    n1 = rng.standard_normal((3, 40, 40)).astype(np.float32)
    cases['uncorrelated'] = (n0, n1, dict(subpixel=True, pad=True))
    # peak on the wrap boundary: shift by half the padded period
    w = rng.standard_normal((1, 32, 32)).astype(np.float32)
    cases['edge_wrap_nopad'] = (w, np.roll(w, (16, 16), axis=(1, 2)), dict(subpixel=True, pad=False))
    cases['single_pixel'] = (np.ones((1, 1, 1), np.float32), np.ones((1, 1, 1), np.float32), dict(subpixel=True))
    cases['odd_sizes'] = (rng.standard_normal((2, 37, 53)).astype(np.float32),
                          rng.standard_normal((2, 41, 29)).astype(np.float32), dict(subpixel=True, pad=True))
    return cases


def seeded_xcorr_cases():
    """Larger cases: only the seed, shapes and outputs are stored."""
    return {
        'seed_256_pad': dict(n=3, size=256, seed=21, max_shift=32, kwargs=dict(subpixel=True, pad=True)),
        'seed_512_pad': dict(n=2, size=512, seed=22, max_shift=32, kwargs=dict(subpixel=True, pad=True)),
        'seed_512_nopad': dict(n=2, size=512, seed=22, max_shift=32, kwargs=dict(subpixel=True, pad=False)),
        'seed_280_pad': dict(n=2, size=280, seed=23, max_shift=20, kwargs=dict(subpixel=True, pad=True)),  # FFT 576
        'seed_600x400': dict(n=2, size=(600, 400), seed=24, max_shift=30, kwargs=dict(subpixel=False, pad=True)),  # 1200 x 800
        # grids of the wider register-resident path
        'seed_150_pad': dict(n=3, size=150, seed=25, max_shift=16, kwargs=dict(subpixel=True, pad=True)),      # FFT 300
        'seed_300_pad': dict(n=2, size=300, seed=26, max_shift=24, kwargs=dict(subpixel=True, pad=True)),      # FFT 600
        'seed_140_std': dict(n=2, size=140, seed=27, max_shift=16, kwargs=dict(subpixel=True, pad=True, conf_mode=1)),   # FFT 288, STD
        'seed_1024_pad': dict(n=1, size=1024, seed=28, max_shift=32, kwargs=dict(subpixel=True, pad=True)),    # FFT 2048
        'seed_2048_pad': dict(n=1, size=2048, seed=29, max_shift=32, kwargs=dict(subpixel=True, pad=True)),    # FFT 4096
        'seed_192_none': dict(n=2, size=192, seed=30, max_shift=20, kwargs=dict(subpixel=True, pad=True, conf_mode=0)),  # FFT 384, NONE
    }


def run_loop_case(h, spec, kind):
    """One case of tests/loop_cases.py through the UNMODIFIED reference (oracle/ref_harness.py).  Returns the flat
    dict of arrays stored in the golden file."""
    import json
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import loop_cases as lc
    from oracle import matcher_oracle as mo
    from oracle.ref_harness import flatten_trace
    img0, img1 = spec['make']()
    kwargs = dict(spec['kwargs'])
    rec = {'kwargs_json': np.asarray(json.dumps(kwargs, sort_keys=True)),
           'input_sum': np.asarray([np.asarray(img0, dtype=np.float64).sum(), np.asarray(img1, dtype=np.float64).sum()])}
    with h:
        if kind == 'stitch':
            if 'masks' in spec:
                kwargs['mask0'], kwargs['mask1'] = spec['masks'](img0, img1)
            out = h.matcher.stitching_matcher(img0, img1, **kwargs)
            xy0, xy1, weight, strain, phtm = out
            if phtm is not None:
                rec['phtm'] = np.asarray(phtm, dtype=np.float64)
        else:
            if spec['prep'] == 'dog':
                img0, img1 = mo.masked_dog_oracle(img0, spec['dog_sigma']), mo.masked_dog_oracle(img1, spec['dog_sigma'])
            h0, w0 = img0.shape
            h1, w1 = img1.shape
            mesh0 = h.AffineMesh.from_bbox((0, 0, w0, h0), cartesian=True, uid=0.0, resolution=4.0)
            mesh1 = h.AffineMesh.from_bbox((0, 0, w1, h1), cartesian=True, uid=1.0, resolution=4.0)
            ld0, ld1 = h.stream_loader(img0, resolution=4.0), h.stream_loader(img1, resolution=4.0)
            if spec.get('initial'):
                p0, p1, w = lc.initial_matches_for(*spec['initial_args']) if 'initial_args' in spec else lc.initial_matches_for(600, 0.02, (30.0, -24.0))
                kwargs['initial_matches'] = h.common.Match(p0, p1, w)
            if spec['entry'] == 'section':
                xy0, xy1, weight, strain = h.matcher.section_matcher(mesh0, mesh1, ld0, ld1, **kwargs)
            else:
                spacings = kwargs.pop('spacings')
                xy0, xy1, weight, strain = h.matcher.iterative_xcorr_matcher_w_mesh(mesh0, mesh1, ld0, ld1, spacings, **kwargs)
        trace = h.take_trace()
    rec['failed'] = np.asarray(xy0 is None)
    if xy0 is not None:
        rec['xy0'], rec['xy1'], rec['weight'] = np.asarray(xy0), np.asarray(xy1), np.asarray(weight)
    rec['weight_or_conf'] = np.asarray(weight if xy0 is None else 0.0, dtype=np.float64)
    rec['strain'] = np.asarray(np.nan if strain is None else strain, dtype=np.float64)
    flatten_trace('trace', trace, rec)
    return rec


def loop_goldens():
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import loop_cases as lc
    from oracle.ref_harness import Harness
    h = Harness()
    for fname, cases, kind in (('loop_stitch.npz', lc.stitch_cases(), 'stitch'), ('loop_section.npz', lc.section_cases(), 'section')):
        blob = {}
        for name, spec in cases.items():
            rec = run_loop_case(h, spec, kind)
            levels = [int(rec[f'trace/{i}/conf'].shape[0]) for i in range(int(rec['trace/n'])) if str(rec[f'trace/{i}/kind']) == 'level']
            print(f'{name}: failed={bool(rec["failed"])} levels={levels} matches={0 if bool(rec["failed"]) else rec["xy0"].shape[0]} strain={float(rec["strain"]):.3e}')
            for k, v in rec.items():
                blob[f'{name}/{k}'] = v
        np.savez_compressed(os.path.join(OUT, fname), **blob)


def main():
    if not ref_loader.available():
        raise SystemExit('reference not present; golden vectors can only be generated in the build container')
    matcher, common, const = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter('ignore')

    # ---- xcorr_fft, stored inputs ---------------------------------------
    blob = {}
    for name, (i0, i1, kw) in xcorr_cases().items():
        dx, dy, cf = matcher.xcorr_fft(i0, i1, **kw)
        blob[f'{name}/img0'] = i0
        blob[f'{name}/img1'] = i1
        for k, v in kw.items():
            blob[f'{name}/kw/{k}'] = np.asarray(v)
        blob[f'{name}/dx'], blob[f'{name}/dy'], blob[f'{name}/conf'] = dx, dy, cf
    np.savez_compressed(os.path.join(OUT, 'xcorr_small.npz'), **blob)

    # ---- xcorr_fft, seeded inputs ---------------------------------------
    blob = {}
    for name, spec in seeded_xcorr_cases().items():
        s0, s1, shifts = synth.block_pairs(spec['n'], spec['size'], spec['seed'], max_shift=spec['max_shift'])
        dx, dy, cf = matcher.xcorr_fft(s0, s1, **spec['kwargs'])
        blob[f'{name}/n'] = np.asarray(spec['n'])
        blob[f'{name}/size'] = np.asarray(spec['size'])
        blob[f'{name}/seed'] = np.asarray(spec['seed'])
        blob[f'{name}/max_shift'] = np.asarray(spec['max_shift'])
        blob[f'{name}/input_sum'] = np.asarray([s0.astype(np.float64).sum(), s1.astype(np.float64).sum()])
        blob[f'{name}/shifts'] = shifts
        for k, v in spec['kwargs'].items():
            blob[f'{name}/kw/{k}'] = np.asarray(v)
        blob[f'{name}/dx'], blob[f'{name}/dy'], blob[f'{name}/conf'] = dx, dy, cf
    np.savez_compressed(os.path.join(OUT, 'xcorr_seeded.npz'), **blob)

    # ---- host helpers + global translation + DoG -------------------------
    blob = {}
    boxes = [((0, 0, 3000, 400), dict(block_size=750.0000000000001, min_num_blocks=1)),
             ((0, 0, 3000, 400), dict(block_size=75.0, min_num_blocks=2)),
             ((10, -7, 411, 3013), dict(block_size=74.99999999999999, min_num_blocks=2, shrink_factor=0.7)),
             ((0, 0, 250, 1500), dict(min_num_blocks=(3, 2))),
             ((5, 5, 6, 6), dict(block_size=100))]
    for i, (bb, kw) in enumerate(boxes):
        out = common.divide_bbox(bb, **kw)
        blob[f'divide/{i}/bbox'] = np.asarray(bb)
        for k, v in kw.items():
            blob[f'divide/{i}/kw/{k}'] = np.asarray(v)
        blob[f'divide/{i}/out'] = np.stack(out, axis=0)
    rng = np.random.default_rng(5)
    pts = rng.integers(0, 23, size=(200, 2)).astype(np.float64)
    blob['zorder/in'] = pts
    blob['zorder/out'] = common.z_order(pts)
    bbs = rng.integers(-50, 500, size=(17, 4)).astype(np.float64)
    blob['bbox/in'] = bbs
    blob['bbox/centers'] = common.bbox_centers(bbs)
    blob['bbox/sizes'] = common.bbox_sizes(bbs)

    class _Box:                       # duck-typed stand-in for Mesh.bbox(gear=...)
        def __init__(self, b):
            self._b = b

        def bbox(self, gear=None):
            return self._b
    for i, (b0, b1, sp, kw) in enumerate([
            ((12.0, -3.0, 3012.0, 497.0), (0, 0, 3000, 500), 750.0000000000001, dict(min_num_blocks=1, zorder=True)),
            ((12.0, -3.0, 3012.0, 497.0), (0, 0, 3000, 500), 75.0, dict(min_num_blocks=2, zorder=True)),
            ((0, 0, 2048, 2048), (3, 5, 2051, 2053), 150, dict(min_num_blocks=1, shrink_factor=0.7, zorder=True))]):
        o0, o1 = matcher.distributor_cartesian_bbox(_Box(b0), _Box(b1), sp, **kw)
        blob[f'cart/{i}/bbox0'], blob[f'cart/{i}/bbox1'] = np.asarray(b0), np.asarray(b1)
        blob[f'cart/{i}/spacing'] = np.asarray(sp)
        for k, v in kw.items():
            blob[f'cart/{i}/kw/{k}'] = np.asarray(v)
        blob[f'cart/{i}/out0'], blob[f'cart/{i}/out1'] = o0, o1

    # masked DoG: plain, masked, batch with global ptp, uint8 input
    img = synth.em_canvas(96, 120, seed=31).astype(np.float32)
    msk = np.ones(img.shape, bool)
    msk[:, 90:] = False
    msk[:10, :] = False
    stack = np.stack([img, img[::-1] * 0.5], 0)
    mstack = np.stack([msk, np.ones_like(msk)], 0)
    dog = {'plain': (img, 2.5, None), 'masked': (img, 2.5, msk), 'u8': (img.astype(np.uint8), 3.5, msk),
           'stack_masked': (stack, 1.25, mstack), 'allmask_true': (img, 2.0, np.ones_like(msk))}
    for name, (im, sg, mk) in dog.items():
        blob[f'dog/{name}/img'] = im
        blob[f'dog/{name}/sigma'] = np.asarray(sg)
        if mk is not None:
            blob[f'dog/{name}/mask'] = mk
        blob[f'dog/{name}/out'] = common.masked_dog_filter(im, sg, mask=mk)

    # global translation: one confident pair, one that falls to the block retry
    cv = synth.em_canvas(260, 700, seed=41)
    f = synth.dog_f32(cv)
    g0 = f[5:205, 20:520].copy()
    g1 = f[12:212, 60:560].copy()
    blob['gt/conf/img0'], blob['gt/conf/img1'] = g0, g1
    blob['gt/conf/out'] = np.asarray(matcher.global_translation_matcher(g0, g1, conf_thresh=0.3))
    h0 = g0.copy()
    h1 = g1.copy()
    nrng = np.random.default_rng(42)
    h0[:, 170:] = nrng.standard_normal(h0[:, 170:].shape).astype(np.float32) * 40   # corrupt most of the strip
    h1[:, 200:] = nrng.standard_normal(h1[:, 200:].shape).astype(np.float32) * 40
    blob['gt/retry/img0'], blob['gt/retry/img1'] = h0, h1
    blob['gt/retry/out'] = np.asarray(matcher.global_translation_matcher(h0, h1, conf_thresh=0.9))
    blob['gt/retry_df/out'] = np.asarray(matcher.global_translation_matcher(h0, h1, conf_thresh=0.9, divide_factor=(1, 4)))
    k0 = h0.copy()
    k0[:, :250] = 0.0                   # flat sub-blocks are skipped (np.ptp == 0)
    blob['gt/flat/img0'], blob['gt/flat/img1'] = k0, h1
    blob['gt/flat/out'] = np.asarray(matcher.global_translation_matcher(k0, h1, conf_thresh=0.9))
    np.savez_compressed(os.path.join(OUT, 'matcher_host.npz'), **blob)

    # ---- coarse-to-fine loops: stitching_matcher / section_matcher / iterative_xcorr_matcher_w_mesh -------------
    loop_goldens()

    import scipy
    import cv2
    with open(os.path.join(OUT, 'PROVENANCE.txt'), 'w') as fh:
        fh.write('generated by oracle/make_golden.py from the unmodified reference at /root/reference\n')
        fh.write(f'reference version: feabas 3.0.6 (setup.py:4)\n')
        fh.write(f'scipy {scipy.__version__}, numpy {np.__version__}, opencv {cv2.__version__}, python {sys.version.split()[0]}\n')
        fh.write('loop_*.npz: the unmodified stitching_matcher / section_matcher / iterative_xcorr_matcher_w_mesh / '
                 'bboxes_mesh_renderer_matcher / MeshRenderer.crop_multiple with the geometry layer (Mesh, SLM, MeshRenderer.from_mesh, '
                 'five shapely calls) replaced by the affine stand-ins, see oracle/ref_harness.py\n')
    print('golden vectors written to', OUT)


if __name__ == '__main__':
    main()
