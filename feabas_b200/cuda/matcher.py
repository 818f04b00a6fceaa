"""Matcher layers above ``xcorr_fft``, same names / arguments / return conventions as
``feabas/matcher.py``, with the pixel work on the GPU.

* ``global_translation_matcher``        feabas/matcher.py:138-221
* ``stitching_matcher``                 feabas/matcher.py:224-367
* ``section_matcher``                   feabas/matcher.py:370-427
* ``iterative_xcorr_matcher_w_mesh``    feabas/matcher.py:430-778
* ``bboxes_mesh_renderer_matcher``      feabas/matcher.py:781-861

The control flow (pyramid level selection, confidence filtering, pad rule) is host Python, as in the
reference.  Geometry and relaxation (the reference's ``Mesh`` / ``SLM`` / ``MeshRenderer``, SURVEY
section 2 rows 9, 11, 12) are outside this package: the loop talks to them through the handful of
methods the reference itself uses, so the reference's own objects plug in when FEABAS is
installed; ``feabas_b200.cuda.surrogate`` provides an affine stand-in (``AffineMesh`` /
``AffineSLM``) that keeps every matcher function runnable -- and testable -- without them.
"""
import numpy as np

from . import blocks as _blk
from . import image as _img
from . import _lib
from .constant import (ANNEAL_CONNECTED_RIGID as _ANNEAL_CONNECTED_RIGID, ANNEAL_COPY_EXACT as _ANNEAL_COPY_EXACT,
                       DEFAULT_AVG_DEFORM, DEFAULT_RESOLUTION, DEFAULT_THICKNESS, FFT_CONF_MIRROR, MESH_GEAR_FIXED,
                       MESH_GEAR_INITIAL, MESH_GEAR_MOVING, RENDER_FULL, Match)
from .surrogate import AffineMesh, AffineSLM, ArrayLoader
from .xcorr import fft_shape, xcorr_fft, xcorr_fft_device

try:
    import torch
except Exception:                       # pragma: no cover
    torch = None


# --------------------------------------------------------------------------------------------
# global translation
# --------------------------------------------------------------------------------------------
def _fit_window(lo, hi, want, limit):
    """Grow [lo, hi) symmetrically to ``want`` pixels, slide it back inside [0, limit], clip."""
    grow = int(np.ceil((want - (hi - lo)) / 2))
    lo, hi = lo - grow, hi + grow
    slide = -min(lo, 0) - max(hi - limit, 0)
    return int(min(max(lo + slide, 0), limit)), int(min(max(hi + slide, 0), limit))


def _translation_blocks(windows):
    """integer windows (x_lo, y_lo) -> fb_crop_blocks rows for an identity map"""
    rows = np.zeros((len(windows), 10), dtype=np.float64)
    for i, (x_lo, y_lo) in enumerate(windows):
        rows[i] = (x_lo, y_lo, 1, 1, 1, 0, 0, 0, 1, 0)
    return rows


def global_translation_matcher(img0, img1, **kwargs):
    """Whole-image translation (tx, ty, conf): one padded cross-correlation of the two images and, when its
    confidence is not above ``conf_thresh``, a second try on a grid of sub-blocks.

    ``img0`` / ``img1``: 2-D numpy arrays or CUDA tensors (already band-passed unless ``sigma`` > 0).
    kwargs as the reference: ``sigma`` (0), ``mask0``/``mask1``, ``conf_mode`` (MIRROR), ``conf_thresh`` (0.3),
    ``divide_factor`` (6).  Returns Python floats.
    """
    sigma = kwargs.get('sigma', 0.0)
    conf_mode = kwargs.get('conf_mode', FFT_CONF_MIRROR)
    conf_thresh = kwargs.get('conf_thresh', 0.3)
    divide_factor = kwargs.get('divide_factor', 6)
    dev = kwargs.get('device', None)
    a = _img.to_device(img0, dev)
    b = _img.to_device(img1, a.device.index)
    if a.dtype != b.dtype:
        a, b = a.to(torch.float64), b.to(torch.float64)
    if sigma > 0:
        a = _masked_dog_any(a, sigma, kwargs.get('mask0', None))
        b = _masked_dog_any(b, sigma, kwargs.get('mask1', None))
    tx, ty, conf = _whole_image_translations(a[None], b[None], conf_mode)[0]
    if conf > conf_thresh:
        return tx, ty, conf
    return _translation_retry(a, b, tx, ty, conf, conf_mode, divide_factor)


def _whole_image_translations(stack0, stack1, conf_mode):
    """First shot of ``global_translation_matcher`` (feabas/matcher.py:148-156) for a STACK of image pairs of one
    shape: one batched padded cross-correlation, one read-back.  -> list of (tx, ty, conf)."""
    h0, w0 = stack0.shape[-2:]
    h1, w1 = stack1.shape[-2:]
    res = xcorr_fft_device(stack0, stack1, conf_mode=conf_mode, pad=True).cpu().numpy()
    return [(float(res[0, k]) + (w1 - w0) / 2, float(res[1, k]) + (h1 - h0) / 2, _conf_scalar(res[2, k], conf_mode))
            for k in range(res.shape[1])]


def _translation_retry(a, b, tx, ty, conf, conf_mode, divide_factor):
    """Second shot (feabas/matcher.py:159-221): the images cut into ``divide_factor`` blocks, the most confident
    block wins if it beats the whole-image confidence."""
    h0, w0 = a.shape[-2:]
    h1, w1 = b.shape[-2:]
    rows_cols = _blk.balanced_division(np.minimum((h0, w0), (h1, w1)), divide_factor)
    xa0, ya0, xb0, yb0 = _blk.divide_bbox((0, 0, w0, h0), min_num_blocks=rows_cols)
    xa1, ya1, xb1, yb1 = _blk.divide_bbox((0, 0, w1, h1), min_num_blocks=rows_cols)
    win0, win1 = [], []
    for k in range(xa0.size):
        want_w = max(xb0[k] - xa0[k], xb1[k] - xa1[k])
        want_h = max(yb0[k] - ya0[k], yb1[k] - ya1[k])
        win0.append(_fit_window(xa0[k], xb0[k], want_w, w0) + _fit_window(ya0[k], yb0[k], want_h, h0))
        win1.append(_fit_window(xa1[k], xb1[k], want_w, w1) + _fit_window(ya1[k], yb1[k], want_h, h1))
    win0, win1 = np.array(win0), np.array(win1)          # columns: x_lo, x_hi, y_lo, y_hi
    size0 = np.stack((win0[:, 3] - win0[:, 2], win0[:, 1] - win0[:, 0]), axis=-1)
    size1 = np.stack((win1[:, 3] - win1[:, 2], win1[:, 1] - win1[:, 0]), axis=-1)
    if np.any(size0 != size0[0]) or np.any(size1 != size1[0]):
        # the reference np.stack()s the blocks and raises on ragged ones
        raise ValueError('all input arrays must have the same shape')
    if a.dtype in (torch.float32, torch.uint8):
        s0 = _img.crop_blocks(a, _translation_blocks(win0[:, [0, 2]]), size0[0], origin=(0, 0))
        s1 = _img.crop_blocks(b, _translation_blocks(win1[:, [0, 2]]), size1[0], origin=(0, 0))
        flat = (np.ptp(_img.stack_minmax(s0).cpu().numpy(), axis=-1) == 0) | (np.ptp(_img.stack_minmax(s1).cpu().numpy(), axis=-1) == 0)
    else:
        s0 = torch.stack([a[w[2]:w[3], w[0]:w[1]] for w in win0])
        s1 = torch.stack([b[w[2]:w[3], w[0]:w[1]] for w in win1])
        flat = np.array([bool(x.max() == x.min()) or bool(y.max() == y.min()) for x, y in zip(s0, s1)])
    keep = np.nonzero(~flat)[0]                          # constant blocks carry no signal (matcher.py:196,205)
    if keep.size == 0:
        return tx, ty, conf
    if keep.size < flat.size:
        sel = _img.upload_small(keep, s0.device)
        s0, s1 = s0.index_select(0, sel).contiguous(), s1.index_select(0, sel).contiguous()
    res = xcorr_fft_device(s0, s1, conf_mode=conf_mode, pad=True).cpu().numpy()
    off_x = ((win1[:, 1] - win1[:, 0]) - (win0[:, 1] - win0[:, 0])) / 2 + win1[:, 0] - win0[:, 0]
    off_y = ((win1[:, 3] - win1[:, 2]) - (win0[:, 3] - win0[:, 2])) / 2 + win1[:, 2] - win0[:, 2]
    block_tx, block_ty = res[0] + off_x[keep], res[1] + off_y[keep]
    block_conf = res[2].astype(np.float64 if conf_mode == 1 else np.float32)
    best = int(np.argmax(block_conf))
    if block_conf[best] >= conf:
        tx, ty, conf = float(block_tx[best]), float(block_ty[best]), _conf_scalar(block_conf[best], conf_mode)
    return tx, ty, conf


def _conf_scalar(v, conf_mode):
    # the reference returns conf.item() of a float32 array (float64 for FFT_CONF_STD)
    return float(v) if conf_mode == 1 else float(np.float32(v))


def _masked_dog_any(t, sigma, mask):
    if mask is not None:
        if bool(mask.all()):
            mask = None
        else:
            mask = _img.to_device(mask, t.device.index)
    return _img.masked_dog_device(t, sigma, mask)


# --------------------------------------------------------------------------------------------
# one pass over a block grid
# --------------------------------------------------------------------------------------------
def _as_loader(loader, device=None):
    """ArrayLoader as is; a reference ``StreamLoader`` (feabas/dal.py:1008) or any object exposing the in-RAM
    image as ``_img`` with ``bounds`` is wrapped (image uploaded once)."""
    if isinstance(loader, ArrayLoader):
        return loader
    if hasattr(loader, '_img') and hasattr(loader, 'bounds'):
        cached = getattr(loader, '_fb_device_loader', None)
        if cached is None:
            b = loader.bounds
            cached = ArrayLoader(loader._img, fillval=getattr(loader, '_default_fillval', 0),
                                 resolution=getattr(loader, 'resolution', 4.0), x0=b[0], y0=b[1], device=device)
            try:
                loader._fb_device_loader = cached
            except AttributeError:
                pass
        return cached
    raise TypeError(f'cannot take pixels from {type(loader).__name__}: pass a feabas_b200.cuda.ArrayLoader '
                    '(or a feabas StreamLoader)')


def _block_rows(mesh, loader, bboxes):
    """fb_crop_blocks rows for the blocks ``bboxes`` (output / MOVING frame) of an affine mesh over ``loader``:
    source pixel = (moving @ Ainv + tinv) - loader origin   (feabas/renderer.py:419-450)."""
    ainv, tinv = mesh.render_map()
    if loader.resolution != mesh.resolution:
        # crop_multiple rescales the field to the loader's resolution, pixel centres kept (renderer.py:621-624,
        # spatial.scale_coordinates): one more affine step, folded into the map
        scale = mesh.resolution / loader.resolution
        ainv, tinv = ainv * scale, (tinv + 0.5) * scale - 0.5
    b = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    wd = np.round(b[:, 2] - b[:, 0])
    ht = np.round(b[:, 3] - b[:, 1])
    rows = np.empty((b.shape[0], 10), dtype=np.float64)
    rows[:, 0], rows[:, 1] = b[:, 0], b[:, 1]
    rows[:, 2], rows[:, 3] = (b[:, 2] - b[:, 0]) / wd, (b[:, 3] - b[:, 1]) / ht
    rows[:, 4], rows[:, 5], rows[:, 6] = ainv[0, 0], ainv[1, 0], tinv[0] - loader.x0
    rows[:, 7], rows[:, 8], rows[:, 9] = ainv[0, 1], ainv[1, 1], tinv[1] - loader.y0
    return rows, (int(ht[0]), int(wd[0]))


def _render_stack(mesh, loader, bboxes, sigma, ptp_hint=None, mask_range=None):
    """Blocks of one batch as an ``N x H x W`` CUDA tensor (band-passed when ``sigma`` > 0); ``None`` when no pixel
    of the batch is covered by the mesh (``crop_multiple`` returns None then, feabas/common.py:275-279)."""
    rows, shape = _block_rows(mesh, loader, bboxes)
    cover = None
    if sigma > 0:                                       # precise_mask=log_sigma>0 (feabas/renderer.py:493,506)
        x_lo, y_lo, x_hi, y_hi = mesh.covered_rect()
        if loader.resolution != mesh.resolution:
            scale = mesh.resolution / loader.resolution
            x_lo, y_lo, x_hi, y_hi = ((v + 0.5) * scale - 0.5 for v in (x_lo, y_lo, x_hi, y_hi))
        cover = (x_lo - loader.x0, y_lo - loader.y0, x_hi - loader.x0, y_hi - loader.y0)
    return _render_rows(loader, rows, shape, sigma, cover, ptp_hint, mask_range)


def _render_rows(loader, rows, shape, sigma, cover=None, ptp_hint=None, mask_range=None):
    """``fb_crop_blocks`` rows -> stack (band-passed when ``sigma`` > 0).  ``cover``: rectangle of the source the mesh
    covers (loader pixel frame), or None when every block is known to be rendered whole."""
    if sigma <= 0:
        return _img.crop_blocks_masked(loader.tensor, rows, shape, fillval=loader.default_fillval, cover=None)[0]
    if mask_range is not None:
        # renderer.py:633-636: only intensities inside the range are valid -- a mask for every block
        stack, mask = _img.crop_blocks_masked(loader.tensor, rows, shape, fillval=loader.default_fillval, cover=cover)
        rng = np.atleast_1d(mask_range)
        valid = (stack >= float(rng[0])) & (stack <= float(rng[-1]))
        mask = valid if mask is None else (mask.to(torch.bool) & valid)
        lo, hi = (int(v) for v in torch.stack((mask.min(), mask.max())).cpu())
        if hi == 0:
            return None
        return _img.masked_dog_device(stack, sigma, None if lo == 1 else mask, ptp=ptp_hint)
    # masks only for the blocks that hang over the border of the mesh (a block without masked pixels gets no mask term:
    # the term is zero there); nothing is read back from the device
    stack, mask, mask_images, any_covered = _img.crop_blocks_masked(loader.tensor, rows, shape, fillval=loader.default_fillval,
                                                                    cover=cover, compact=True)
    if not any_covered:
        return None
    return _img.masked_dog_device(stack, sigma, mask, ptp=ptp_hint, mask_images=mask_images)


def bboxes_mesh_renderer_matcher(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1, **kwargs):
    """Render the blocks ``bboxes0`` / ``bboxes1`` of the two sections and cross-correlate them pairwise.

    Returns ``(xy0, xy1, conf)``: matched points (K x 2 float64, in the frame the bboxes live in) and the
    per-block confidence.  kwargs as the reference: ``batch_size``, ``sigma`` (0), ``conf_mode`` (MIRROR),
    ``pad`` (True), ``subpixel`` (False), ``mask_range`` (None); ``render_mode`` / ``geodesic_mask`` /
    ``affine_approx_tol`` / ``render_weight_threshold`` are accepted and only meaningful for reference meshes.
    """
    if not (hasattr(mesh0, 'render_map') and hasattr(mesh1, 'render_map')):
        return _reference_block_pass(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1, **kwargs)
    return _bboxes_collect(_bboxes_enqueue(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1, **kwargs))


def _bboxes_enqueue(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1, **kwargs):
    """Device half of ``bboxes_mesh_renderer_matcher``: every batch rendered and cross-correlated, nothing read back.
    Returns what ``_bboxes_collect`` needs."""
    batch_size = kwargs.get('batch_size', None)
    sigma = kwargs.get('sigma', 0.0)
    conf_mode = kwargs.get('conf_mode', FFT_CONF_MIRROR)
    pad = kwargs.get('pad', True)
    subpixel = kwargs.get('subpixel', False)
    mask_range = kwargs.get('mask_range', None)
    pending = []
    if bboxes0 is None or len(bboxes0) == 0:
        return pending, None, None, conf_mode
    loader0 = _as_loader(image_loader0, kwargs.get('device', None))
    loader1 = _as_loader(image_loader1, loader0.tensor.device.index)
    bboxes0, bboxes1 = np.asarray(bboxes0), np.asarray(bboxes1)
    edges = _blk.split_batches(bboxes0, bboxes1, batch_size)
    for lo, hi in zip(edges[:-1], edges[1:]):
        if hi <= lo:
            continue
        stack0 = _render_stack(mesh0, loader0, bboxes0[lo:hi], sigma, mask_range=mask_range)
        if stack0 is None:
            continue
        stack1 = _render_stack(mesh1, loader1, bboxes1[lo:hi], sigma, mask_range=mask_range)
        if stack1 is None:
            continue
        res = xcorr_fft_device(stack0, stack1, conf_mode=conf_mode, pad=pad, subpixel=subpixel)
        # the read-back is enqueued NOW, right behind this batch (page-locked landing buffer + event): a copy issued at
        # collection time would queue up behind whatever was enqueued in between (the next jobs of the _many entry point)
        landed = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
        landed.copy_(res, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(res.device))
        pending.append((lo, hi, landed, done))
    return pending, bboxes0, bboxes1, conf_mode


def _bboxes_collect(enqueued):
    """Host half: wait for each batch's read-back, block displacements -> matched points (matcher.py:840-849)."""
    pending, bboxes0, bboxes1, conf_mode = enqueued
    if not pending:
        return np.empty((0, 2)), np.empty((0, 2)), np.empty(0)
    xy0, xy1, conf = [], [], []
    for lo, hi, landed, done in pending:
        done.synchronize()
        res = landed.numpy()
        p0, p1 = _blk.block_points(bboxes0[lo:hi], bboxes1[lo:hi], res[0], res[1])
        xy0.append(p0)
        xy1.append(p1)
        conf.append(res[2].astype(np.float64 if conf_mode == 1 else np.float32))
    return np.concatenate(xy0, axis=0), np.concatenate(xy1, axis=0), np.concatenate(conf, axis=0)


_JOB_STREAMS = {}


def _job_streams(device, count):
    """``count`` side streams per device for the job-list entry point (created once)."""
    have = _JOB_STREAMS.setdefault(device.index, [])
    while len(have) < count:
        have.append(torch.cuda.Stream(device=device))
    return have[:count]


def bboxes_mesh_renderer_matcher_many(jobs, depth=2, streams=2, **kwargs):
    """``bboxes_mesh_renderer_matcher`` over MANY independent section pairs (the aligner's job list, feabas/aligner.py:
    one pair per job; ``align_main.py`` fans them out to worker processes) with the device kept busy: the block passes
    of up to ``depth`` jobs are enqueued before the results of the oldest one are read back, so uploads, renders and
    correlations of the next pair run under the host's share of the previous one.  With ``streams`` > 1 consecutive
    jobs alternate between that many CUDA streams, so the image kernels of one pair (issue bound) and the FFT kernels
    of another (FP32 / shared-memory / HBM bound) share the SMs.  ``jobs``: iterable of
    ``(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1)`` -- consumed lazily, a generator may build the
    loaders (and start their uploads) on demand.  Returns the list of ``(xy0, xy1, conf)``, identical to one call per job."""
    from collections import deque
    queue, results = deque(), []
    side, home, count = None, None, 0
    for job in jobs:
        mesh0, mesh1 = job[0], job[1]
        if not (hasattr(mesh0, 'render_map') and hasattr(mesh1, 'render_map')):
            while queue:
                results.append(_bboxes_collect(queue.popleft()))
            results.append(_reference_block_pass(*job, **kwargs))
            continue
        if streams > 1:
            if side is None:
                # (the device without touching .tensor: a pending upload must be waited for -- and its memory recorded --
                # on the stream that consumes it, not on the caller's)
                dev = _as_loader(job[2], kwargs.get('device', None))._tensor.device
                home = torch.cuda.current_stream(dev)
                side = _job_streams(dev, int(streams))
                for st in side:
                    st.wait_stream(home)                  # whatever produced the inputs was ordered on the caller's stream
            with torch.cuda.stream(side[count % len(side)]):
                queue.append(_bboxes_enqueue(*job, **kwargs))
            count += 1
        else:
            queue.append(_bboxes_enqueue(*job, **kwargs))
        while len(queue) > max(int(depth), 0):
            results.append(_bboxes_collect(queue.popleft()))
    while queue:
        results.append(_bboxes_collect(queue.popleft()))
    if side is not None:
        for st in side:
            home.wait_stream(st)                          # later work on the caller's stream sees the side streams drained
    return results


_MERGE_BYTES = 1 << 30          # blocks of one merged launch: at most this many bytes per stack


def _mergeable(req):
    """Block passes that may share launches with other overlaps / section pairs: affine meshes, no per-batch
    band-pass (``sigma`` == 0: the stitching and thumbnail paths filter whole images up front)."""
    mesh0, mesh1, loader0, loader1 = req.args[:4]
    kw = req.kwargs
    return (hasattr(mesh0, 'render_map') and hasattr(mesh1, 'render_map') and kw.get('sigma', 0.0) == 0 and
            isinstance(loader0, ArrayLoader) and isinstance(loader1, ArrayLoader))


def _block_pass_many(requests):
    """``bboxes_mesh_renderer_matcher`` for a list of ``_BlockPass`` requests -> list of ``(xy0, xy1, conf)``.

    Every request is cut into batches exactly as the reference cuts it (feabas/matcher.py:804-822: new batch where the
    block size changes, at most ``batch_size`` blocks); batches of EQUAL block shapes, padding and sub-pixel flags are
    then concatenated across requests: one ``fb_crop_blocks_multi`` per side, one ``xcorr_fft`` launch sequence and, for
    the whole round, one device -> host copy.  A block keeps the source-crop origin of its own batch, and block pairs
    are independent in the xcorr kernels, so the numbers equal those of the one-request-at-a-time path bit for bit."""
    answers = [None] * len(requests)
    groups = {}
    for i, req in enumerate(requests):
        if not _mergeable(req):
            answers[i] = bboxes_mesh_renderer_matcher(*req.args, **req.kwargs)
            continue
        mesh0, mesh1, loader0, loader1, boxes0, boxes1 = req.args
        kw = req.kwargs
        if boxes0 is None or len(boxes0) == 0:
            answers[i] = (np.empty((0, 2)), np.empty((0, 2)), np.empty(0))
            continue
        boxes0, boxes1 = np.asarray(boxes0), np.asarray(boxes1)
        conf_mode = kw.get('conf_mode', FFT_CONF_MIRROR)
        edges = _blk.split_batches(boxes0, boxes1, kw.get('batch_size', None))
        for lo, hi in zip(edges[:-1], edges[1:]):
            if hi <= lo:
                continue
            rows0, shape0 = _block_rows(mesh0, loader0, boxes0[lo:hi])
            rows1, shape1 = _block_rows(mesh1, loader1, boxes1[lo:hi])
            t0, t1 = loader0.tensor, loader1.tensor
            key = (shape0, shape1, bool(kw.get('pad', True)), bool(kw.get('subpixel', False)), conf_mode, t0.dtype, t1.dtype,
                   t0.device.index, float(loader0.default_fillval), float(loader1.default_fillval))
            groups.setdefault(key, []).append((i, lo, hi, rows0, rows1, t0, t1))
    pending = []
    for key, items in groups.items():
        shape0, shape1, pad, subpixel, conf_mode, dt0, dt1, dev, fill0, fill1 = key
        per_block = max(shape0[0] * shape0[1] * _itemsize(dt0), shape1[0] * shape1[1] * _itemsize(dt1))
        limit = max(1, _MERGE_BYTES // per_block)
        start = 0
        while start < len(items):                              # whole batches per launch, bounded memory
            stop, count = start, 0
            while stop < len(items) and (count == 0 or count + items[stop][2] - items[stop][1] <= limit):
                count += items[stop][2] - items[stop][1]
                stop += 1
            part = items[start:stop]
            stack0 = _img.crop_blocks_multi([(it[5], it[3]) for it in part], shape0, fillval=fill0)
            stack1 = _img.crop_blocks_multi([(it[6], it[4]) for it in part], shape1, fillval=fill1)
            pending.append((part, conf_mode, xcorr_fft_device(stack0, stack1, conf_mode=conf_mode, pad=pad, subpixel=subpixel)))
            start = stop
    if pending:
        flat = torch.cat([res for _, _, res in pending], dim=1).cpu().numpy()      # the round's only synchronising read
        pieces = {}
        col = 0
        for part, conf_mode, _ in pending:
            for i, lo, hi, *_ in part:
                res = flat[:, col:col + hi - lo]
                col += hi - lo
                boxes0, boxes1 = np.asarray(requests[i].args[4]), np.asarray(requests[i].args[5])
                p0, p1 = _blk.block_points(boxes0[lo:hi], boxes1[lo:hi], res[0], res[1])
                pieces.setdefault(i, []).append((lo, p0, p1, res[2].astype(np.float64 if conf_mode == 1 else np.float32)))
        for i, parts in pieces.items():
            parts.sort(key=lambda x: x[0])
            answers[i] = (np.concatenate([x[1] for x in parts], axis=0), np.concatenate([x[2] for x in parts], axis=0),
                          np.concatenate([x[3] for x in parts], axis=0))
    return answers


def _itemsize(dtype):
    return torch.empty((), dtype=dtype).element_size()


def _reference_block_pass(mesh0, mesh1, image_loader0, image_loader1, bboxes0, bboxes1, **kwargs):
    """Reference ``Mesh`` objects (FEABAS installed): the reference's own ``MeshRenderer`` is built for both sections
    (feabas/matcher.py:792-829) and handed to ``bboxes_renderer_matcher``."""
    try:
        import feabas.matcher as ref
        import feabas.renderer as ref_renderer
    except Exception as exc:            # pragma: no cover - FEABAS is not installed in the build container
        raise TypeError('meshes without render_map() need the FEABAS package for rendering') from exc
    if isinstance(mesh0, dict):         # pragma: no cover
        mesh0 = ref.Mesh(**mesh0)
    elif isinstance(mesh0, str):        # pragma: no cover
        mesh0 = ref.Mesh.from_h5(mesh0)
    if isinstance(mesh1, dict):         # pragma: no cover
        mesh1 = ref.Mesh(**mesh1)
    elif isinstance(mesh1, str):        # pragma: no cover
        mesh1 = ref.Mesh.from_h5(mesh1)
    if isinstance(image_loader0, (str, dict)):      # pragma: no cover
        image_loader0 = ref.dal.get_loader_from_json(image_loader0)
    if isinstance(image_loader1, (str, dict)):      # pragma: no cover
        image_loader1 = ref.dal.get_loader_from_json(image_loader1)
    make = dict(geodesic_mask=kwargs.get('geodesic_mask', False), render_weight_threshold=kwargs.get('render_weight_threshold', 0),
                affine_approx_tol=kwargs.get('affine_approx_tol', 0.0))
    render0 = ref.MeshRenderer.from_mesh(mesh0, image_loader=image_loader0, **make)
    render1 = ref.MeshRenderer.from_mesh(mesh1, image_loader=image_loader1, **make)
    if render0 is None or render1 is None:
        return np.empty((0, 2)), np.empty((0, 2)), np.empty(0)
    return bboxes_renderer_matcher(render0, render1, image_loader0, image_loader1, bboxes0, bboxes1,
                                   renderer_module=ref_renderer, **kwargs)


def renderer_block_rows(render, bboxes, log_sigma, loader_resolution, renderer_module):
    """``fb_crop_blocks`` rows (source coordinates in the loader's global pixel frame) for the blocks ``bboxes`` of a
    REFERENCE ``MeshRenderer`` -- or None when the batch cannot go through the affine gather.

    ``MeshRenderer.crop_field`` (feabas/renderer.py:497-512) first tries an affine map per block: the global fit when its
    residue is below the tolerance, else a fit over the triangles that touch the block (``bbox_affine_tform``,
    renderer.py:395-416); only when that misses the tolerance too does it interpolate the piecewise-linear field.  The
    matcher asks for this with ``affine_approximated_render=True`` (the default, feabas/matcher.py:507,586-603: 0.1 px at
    the finest level, max(1, 2 % of the spacing) above).  This function follows exactly those decisions, block by block,
    and evaluates the covered-region rule of ``crop_field_affine`` (renderer.py:436-449) with the reference module's own
    geometry calls; a block that is only partly covered, a geodesic mask, or a block whose fit misses the tolerance
    makes the whole batch fall back to the reference's host renderer (the band-pass sees the batch as one array,
    feabas/common.py:369, so a batch is not split)."""
    approx = getattr(render, '_affine_approximator', None)
    tol = float(getattr(render, '_affine_approx_tol', 0) or 0)
    if approx is None or not tol > 0 or getattr(render, '_geodesic_mask', False):
        return None
    offset = np.tile(np.asarray(render._offset, dtype=np.float64).ravel(), 2)
    covered = getattr(render, '_covered_region', None)
    b = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    rows = np.empty((b.shape[0], 10), dtype=np.float64)
    shape = None
    for i in range(b.shape[0]):
        bbox0 = b[i] - offset
        if approx['global_residue'] < tol:
            full = np.asarray(approx['global_affine'], dtype=np.float64)
            a, t = full[:2, :2], full[-1, :2]
        else:
            a, t, res = render.bbox_affine_tform(bbox0, offsetting=False)
            if a is None or not res < tol:
                return None
        wd, ht = round(bbox0[2] - bbox0[0]), round(bbox0[3] - bbox0[1])
        if shape is None:
            shape = (int(ht), int(wd))
        elif shape != (int(ht), int(wd)):
            return None
        if log_sigma > 0 and covered is not None:       # precise_mask=log_sigma>0
            box0 = renderer_module.shpgeo.box(*(bbox0 - 0.5))
            box1 = renderer_module.shapely.affinity.affine_transform(box0, np.concatenate((a.T, t), axis=None))
            if not (box1.area - covered.intersection(box1).area) < 1:
                return None
        rows[i] = (bbox0[0], bbox0[1], (bbox0[2] - bbox0[0]) / wd, (bbox0[3] - bbox0[1]) / ht,
                   a[0, 0], a[1, 0], t[0], a[0, 1], a[1, 1], t[1])
    if loader_resolution != render.resolution:          # crop_multiple, renderer.py:621-624 (pixel centres kept)
        scale = render.resolution / loader_resolution
        rows[:, [4, 5, 7, 8]] *= scale
        rows[:, [6, 9]] = (rows[:, [6, 9]] + 0.5) * scale - 0.5
    return rows, shape


def _rows_footprint(rows, shape):
    """Integer bounding box (xmin, ymin, xmax, ymax) of the source pixels the rows can touch, with OpenCV's margin."""
    bh, bw = shape
    xe = np.stack((rows[:, 0], rows[:, 0] + (bw - 1) * rows[:, 2]), axis=-1)[:, :, None]
    ye = np.stack((rows[:, 1], rows[:, 1] + (bh - 1) * rows[:, 3]), axis=-1)[:, None, :]
    xs = xe * rows[:, 4, None, None] + ye * rows[:, 5, None, None] + rows[:, 6, None, None]
    ys = xe * rows[:, 7, None, None] + ye * rows[:, 8, None, None] + rows[:, 9, None, None]
    return (int(np.floor(xs.min())) - 6, int(np.floor(ys.min())) - 6, int(np.ceil(xs.max())) + 7, int(np.ceil(ys.max())) + 7)


def _pixels_for_rows(image_loader, rows, shape, device=None):
    """An ``ArrayLoader`` that holds every source pixel the rows touch: the whole image of an in-RAM loader (uploaded
    once and cached on it), else the region the batch needs, cut with the loader's own ``crop`` (tiles read and
    assembled by FEABAS's IO layer, feabas/dal.py) and uploaded."""
    if isinstance(image_loader, ArrayLoader) or hasattr(image_loader, '_img'):
        return _as_loader(image_loader, device)
    region = _rows_footprint(rows, shape)
    img = image_loader.crop(region, return_empty=True)
    return ArrayLoader(np.ascontiguousarray(img), fillval=image_loader.default_fillval, resolution=image_loader.resolution,
                       x0=region[0], y0=region[1], device=device)


def bboxes_renderer_matcher(render0, render1, image_loader0, image_loader1, bboxes0, bboxes1, renderer_module=None, **kwargs):
    """The block pass of ``bboxes_mesh_renderer_matcher`` (feabas/matcher.py:830-861) over two REFERENCE ``MeshRenderer``
    objects.  Batches whose blocks are affine within the renderer's tolerance and fully covered (``renderer_block_rows``)
    are cut from the section on the device (``fb_crop_blocks``), band-passed and correlated there; any other batch is
    rendered by the renderer's own ``crop_multiple`` on the host and correlated by the CUDA ``xcorr_fft``."""
    if renderer_module is None:
        import feabas.renderer as renderer_module       # pragma: no cover - needs FEABAS
    batch_size = kwargs.get('batch_size', None)
    sigma = kwargs.get('sigma', 0.0)
    conf_mode = kwargs.get('conf_mode', FFT_CONF_MIRROR)
    pad = kwargs.get('pad', True)
    subpixel = kwargs.get('subpixel', False)
    mask_range = kwargs.get('mask_range', None)
    render_mode = kwargs.get('render_mode', RENDER_FULL)
    empty = (np.empty((0, 2)), np.empty((0, 2)), np.empty(0))
    if bboxes0 is None or len(bboxes0) == 0:
        return empty
    bboxes0, bboxes1 = np.asarray(bboxes0), np.asarray(bboxes1)
    edges = _blk.split_batches(bboxes0, bboxes1, batch_size)
    conf_dtype = np.float64 if conf_mode == 1 else np.float32
    xy0, xy1, conf = [], [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        if hi <= lo:
            continue
        b0, b1 = bboxes0[lo:hi], bboxes1[lo:hi]
        on_device = render_mode == RENDER_FULL          # (RENDER_CONTIGEOUS / LOCAL_* modes: the reference's own code)
        r0 = renderer_block_rows(render0, b0, sigma, image_loader0.resolution, renderer_module) if on_device else None
        r1 = renderer_block_rows(render1, b1, sigma, image_loader1.resolution, renderer_module) if on_device and r0 is not None else None
        if r0 is not None and r1 is not None:
            pix0 = _pixels_for_rows(image_loader0, *r0, device=kwargs.get('device', None))
            pix1 = _pixels_for_rows(image_loader1, *r1, device=pix0.tensor.device.index)
            rows0, rows1 = r0[0].copy(), r1[0].copy()
            rows0[:, 6] -= pix0.x0; rows0[:, 9] -= pix0.y0
            rows1[:, 6] -= pix1.x0; rows1[:, 9] -= pix1.y0
            stack0 = _render_rows(pix0, rows0, r0[1], sigma, None, None, mask_range)
            stack1 = None if stack0 is None else _render_rows(pix1, rows1, r1[1], sigma, None, None, mask_range)
            if stack0 is None or stack1 is None:
                continue
            res = xcorr_fft_device(stack0, stack1, conf_mode=conf_mode, pad=pad, subpixel=subpixel).cpu().numpy()
            dx, dy, cf = res[0], res[1], res[2].astype(conf_dtype)
        else:
            stack0 = render0.crop_multiple(b0, mode=render_mode, log_sigma=sigma, remap_interp=1, mask_range=mask_range)   # 1 = cv2.INTER_LINEAR
            if stack0 is None:
                continue
            stack1 = render1.crop_multiple(b1, mode=render_mode, log_sigma=sigma, remap_interp=1, mask_range=mask_range)
            if stack1 is None:
                continue
            dx, dy, cf = xcorr_fft(stack0, stack1, conf_mode=conf_mode, pad=pad, subpixel=subpixel)
        p0, p1 = _blk.block_points(b0, b1, dx, dy)
        xy0.append(p0)
        xy1.append(p1)
        conf.append(cf)
    if not xy0:
        return empty
    return np.concatenate(xy0, axis=0), np.concatenate(xy1, axis=0), np.concatenate(conf, axis=0)


# --------------------------------------------------------------------------------------------
# coarse-to-fine loop
# --------------------------------------------------------------------------------------------
def _make_optimizer(mesh0, mesh1, stiffness_lambda, **kwargs):
    if hasattr(mesh0, 'render_map'):
        return AffineSLM([mesh0, mesh1], stiffness_lambda=stiffness_lambda, **kwargs)
    from feabas import optimizer        # pragma: no cover - reference meshes bring the reference solver
    return optimizer.SLM([mesh0, mesh1], stiffness_lambda=stiffness_lambda, **kwargs)   # pragma: no cover


def _relax(opt, linear, tol, steps, opt_kwargs, **extra):
    if linear:
        opt.optimize_linear(tol=tol, **extra, **opt_kwargs)
    else:
        opt.optimize_Newton_Raphson(max_newtonstep=steps, tol=tol, **extra, **opt_kwargs)


def _refine_mode_code(refine_mode):
    if not isinstance(refine_mode, str):
        return refine_mode
    name = refine_mode.lower()
    return 0 if name == 'none' else (1 if 'only' in name else 2)


def _absolute_spacings(spacings, mesh0, mesh1):
    """Spacings below 1 are fractions of the long side of the meshes' common bounding box."""
    spacings = np.array(spacings, dtype=np.float64, copy=True).ravel()
    if np.any(spacings < 1):
        box, valid = _blk.intersect_bbox(mesh0.bbox(gear=MESH_GEAR_MOVING), mesh1.bbox(gear=MESH_GEAR_MOVING))
        if not valid:
            return None
        spacings[spacings < 1] *= max(box[2] - box[0], box[3] - box[1])
    return spacings


def iterative_xcorr_matcher_w_mesh(mesh0, mesh1, image_loader0, image_loader1, spacings, **kwargs):
    """Alternate block matching and mesh relaxation from the coarsest spacing to the finest.

    Returns ``(xy0, xy1, weight, strain)``; ``(None, None, 0, strain)`` when nothing could be matched.
    kwargs and their defaults are the reference's (feabas/matcher.py:485-507).
    """
    return _drive(_coarse_to_fine(mesh0, mesh1, image_loader0, image_loader1, spacings, **kwargs))


class _BlockPass:
    """What the coarse-to-fine loop asks for at every level: one ``bboxes_mesh_renderer_matcher`` call."""
    __slots__ = ('args', 'kwargs')

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs


def _drive(loop):
    """Run one coarse-to-fine loop (a generator that yields ``_BlockPass`` requests) to completion."""
    try:
        request = next(loop)
        while True:
            request = loop.send(bboxes_mesh_renderer_matcher(*request.args, **request.kwargs))
    except StopIteration as stop:
        return stop.value


def _drive_many(loops):
    """Run many independent coarse-to-fine loops in lockstep: at every round the block passes all the loops are
    waiting for are executed TOGETHER (``_block_pass_many``: one gather + one xcorr launch sequence + one read-back
    per block shape, whatever the number of overlaps / section pairs), then every loop relaxes its own meshes on the
    host and asks for its next level.  Same results as driving the loops one after the other."""
    results = [None] * len(loops)
    waiting = {}

    def advance(i, answer):
        try:
            waiting[i] = loops[i].send(answer) if answer is not None else next(loops[i])
        except StopIteration as stop:
            results[i] = stop.value

    for i in range(len(loops)):
        advance(i, None)
    while waiting:
        order = sorted(waiting)
        requests = [waiting.pop(i) for i in order]
        for i, answer in zip(order, _block_pass_many(requests)):
            advance(i, answer)
    return results


def _coarse_to_fine(mesh0, mesh1, image_loader0, image_loader1, spacings, **kwargs):
    """The loop of ``iterative_xcorr_matcher_w_mesh`` as a generator: it yields a ``_BlockPass`` wherever the
    reference calls ``bboxes_mesh_renderer_matcher`` (feabas/matcher.py:626,668) and is sent ``(xy0, xy1, conf)``."""
    num_workers = kwargs.get('num_workers', 1)            # accepted; the GPU path batches instead of forking
    conf_thresh = kwargs.get('conf_thresh', 0.3)
    residue_mode = kwargs.get('residue_mode', 'huber')
    residue_len = kwargs.get('residue_len', 0)
    opt_tol = kwargs.get('opt_tol', None)
    distributor = kwargs.get('distributor', 'cartesian_bbox')
    min_num_blocks = kwargs.get('min_num_blocks', 2)
    shrink_factor = kwargs.get('shrink_factor', 1)
    allow_dwell = kwargs.get('allow_dwell', 0)
    allow_enlarge = kwargs.get('allow_enlarge', False)
    link_weight_decay = kwargs.get('link_weight_decay', 0.0)
    compute_strain = kwargs.get('compute_strain', True)
    batch_size = kwargs.pop('batch_size', None)
    initial_matches = kwargs.get('initial_matches', None)
    pad_request = kwargs.pop('pad', None)
    refine_mode = _refine_mode_code(kwargs.get('refine_mode', 2))
    subpixel_request = kwargs.pop('subpixel', None)
    max_spacing_skip = kwargs.get('max_spacing_skip', 0)
    callback_settings = kwargs.get('callback_settings', {'early_stop_thresh': 0.1, 'chances': 10, 'eval_step': 5})
    render_weight_threshold = kwargs.get('render_weight_threshold', 0)
    stiffness_lambda = kwargs.pop('stiffness_lambda', 1)
    affine_render = kwargs.pop('affine_approximated_render', True)
    del num_workers
    strain = DEFAULT_AVG_DEFORM
    nothing = (None, None, 0, strain)
    if residue_len < 0:
        # negative = in units of section thickness (feabas/matcher.py:523-525)
        residue_len = max(1, abs(residue_len) * _section_thickness() / mesh0.resolution)
    spacings = _absolute_spacings(spacings, mesh0, mesh1)
    if spacings is None:
        return nothing
    opt_kwargs = {'callback_settings': callback_settings, 'check_converge': True}
    linear = mesh0.is_linear and mesh1.is_linear
    one_locked = mesh0.locked or mesh1.locked
    if compute_strain:
        pristine0, pristine1 = mesh0.copy(), mesh1.copy()
    opt = _make_optimizer(mesh0, mesh1, stiffness_lambda)
    if initial_matches is not None:
        opt_kwargs['tolerated_perturbation'] = 0.1
        opt.add_link_from_coordinates(mesh0.uid, mesh1.uid, initial_matches.xy0, initial_matches.xy1,
                                      gear=(MESH_GEAR_INITIAL, MESH_GEAR_INITIAL), weight=initial_matches.weight,
                                      check_duplicates=False, render_weight_threshold=render_weight_threshold)
        opt.optimize_affine_cascade(start_gear=MESH_GEAR_FIXED, target_gear=MESH_GEAR_FIXED, svd_clip=None)
        opt.anneal(gear=(MESH_GEAR_FIXED, MESH_GEAR_MOVING), mode=_ANNEAL_CONNECTED_RIGID)
        if linear:
            opt.optimize_linear(tol=1e-6, precondition='smoothed_aggregation', **opt_kwargs)
        else:
            opt.optimize_Newton_Raphson(max_newtonstep=5, tol=1e-4, precondition='smoothed_aggregation', **opt_kwargs)
    else:
        mesh0.anneal(gear=(MESH_GEAR_MOVING, MESH_GEAR_FIXED), mode=_ANNEAL_COPY_EXACT)
        mesh1.anneal(gear=(MESH_GEAR_MOVING, MESH_GEAR_FIXED), mode=_ANNEAL_COPY_EXACT)
    opt_kwargs['tolerated_perturbation'] = 0.5
    spacings = np.sort(spacings)[::-1]
    finest = spacings[-1]
    spacing = spacings[0]
    level = 0
    started = False
    may_enlarge = bool(allow_enlarge)
    dwelled = 0
    pad = True if pad_request is None else pad_request
    while level < spacings.size:
        at_finest = spacing == finest
        subpixel = at_finest if subpixel_request is None else subpixel_request
        if affine_render:
            affine_tol = 0.1 if at_finest else max(1, 0.02 * spacing)
        else:
            affine_tol = 0
        if distributor == 'cartesian_bbox':
            boxes0, boxes1 = _blk.distributor_cartesian_bbox(mesh0, mesh1, spacing, min_num_blocks=min_num_blocks if at_finest else 1,
                                                             shrink_factor=shrink_factor, zorder=True)
        else:
            boxes0, boxes1 = _region_blocks(mesh0, mesh1, spacing, distributor, at_finest, refine_mode, **kwargs)
        if boxes0 is None:
            return nothing
        xy0, xy1, conf = yield _BlockPass(mesh0, mesh1, image_loader0, image_loader1, boxes0, boxes1,
                                          batch_size=batch_size, pad=pad, subpixel=subpixel,
                                          affine_approx_tol=affine_tol, **kwargs)
        good = conf > conf_thresh
        if not np.any(good):
            if not started:
                return nothing
            break
        if link_weight_decay == 0:
            opt.clear_links()
        else:
            for link in opt.links:
                link._weight = link._weight * link_weight_decay
        xy0, xy1, weight = xy0[good], xy1[good], conf[good]
        max_dis = np.max(np.sum((xy0 - xy1) ** 2, axis=-1)) ** 0.5
        tol = 0.01 / max(1, max_dis) if opt_tol is None else opt_tol
        # blocks of the next level must be at least 4x the largest displacement still unexplained
        need = 4 * max_dis
        target = np.searchsorted(-spacings, -need) - 1
        if may_enlarge and target < 0:
            may_enlarge = False
            level = -1
            spacing = np.ceil(need)
            if pad_request is None:
                pad = True
            continue
        may_enlarge = False
        if target > level:
            target = min(target, level + 1 + max_spacing_skip)
            if pad_request is None:
                pad = target > level + 1              # adjacent level: displacements are small, circular xcorr suffices
            level, dwelled = target, 0
        elif dwelled >= allow_dwell:
            if pad_request is None:
                pad = True
            level, dwelled = level + 1, 0
        else:
            if pad_request is None:
                pad = True
            dwelled += 1
        opt.add_link_from_coordinates(mesh0.uid, mesh1.uid, xy0, xy1, gear=(MESH_GEAR_MOVING, MESH_GEAR_MOVING),
                                      weight=weight, check_duplicates=False, render_weight_threshold=render_weight_threshold)
        if len(opt.links) == 0:
            if not started:
                return nothing
            break
        if max_dis > 0.1:
            _relax(opt, linear, tol, 3, opt_kwargs)
            if residue_len > 0:
                if residue_mode == 'huber':
                    opt.set_link_residue_huber(residue_len)
                elif residue_mode == 'threshold':
                    opt.set_link_residue_threshold(residue_len)
                else:
                    raise ValueError
                changed, _ = opt.adjust_link_weight_by_residue(relax_first=True)
                if changed and level < spacings.size:
                    _relax(opt, linear, tol, 3, opt_kwargs)
        started = True
        if 0 <= level < spacings.size:
            spacing = spacings[level]
    if len(opt.links) == 0:
        return nothing
    last = opt.links[-1]
    xy0 = last.xy0(gear=MESH_GEAR_INITIAL, use_mask=True, combine=True)
    xy1 = last.xy1(gear=MESH_GEAR_INITIAL, use_mask=True, combine=True)
    weight = last.weight(use_mask=True)
    if compute_strain:
        strain = _strain_ratio(pristine0, pristine1, xy0, xy1, weight, stiffness_lambda, one_locked, linear,
                               render_weight_threshold, opt_kwargs)
    return xy0, xy1, weight, strain


def _section_thickness():
    try:
        from feabas.config import section_thickness     # pragma: no cover
        return section_thickness()                      # pragma: no cover
    except Exception:
        return DEFAULT_THICKNESS


def _region_blocks(mesh0, mesh1, spacing, distributor, at_finest, refine_mode, **kwargs):
    """Region based distributors (shapely polygons, feabas/matcher.py:894-1058) belong to the reference's
    geometry layer; with reference meshes they are called as is, the affine stand-in has no holes and uses the
    cartesian grid."""
    if hasattr(mesh0, 'render_map'):
        return _blk.distributor_cartesian_bbox(mesh0, mesh1, spacing, min_num_blocks=kwargs.get('min_num_blocks', 2) if at_finest else 1,
                                               shrink_factor=kwargs.get('shrink_factor', 1), zorder=True)
    from feabas.matcher import distribute_matching_blocks                    # pragma: no cover
    mode = refine_mode if (at_finest or refine_mode != 2) else 0             # pragma: no cover
    return distribute_matching_blocks(mesh0, mesh1, spacing, dfunc=distributor, refine_mode=mode,   # pragma: no cover
                                      min_boundary_distance=kwargs.get('min_boundary_distance', 0),
                                      shrink_factor=kwargs.get('shrink_factor', 1), zorder=True,
                                      render_weight_threshold=kwargs.get('render_weight_threshold', 0))


def _strain_ratio(mesh0, mesh1, xy0, xy1, weight, stiffness_lambda, one_locked, linear, render_weight_threshold, opt_kwargs):
    """Elastic energy of the final matches relative to that of the rigid shape (feabas/matcher.py:752-777)."""
    opt = _make_optimizer(mesh0, mesh1, stiffness_lambda, assert_dominance=(not one_locked))
    opt.add_link_from_coordinates(mesh0.uid, mesh1.uid, xy0, xy1, gear=(MESH_GEAR_INITIAL, MESH_GEAR_INITIAL), weight=weight,
                                  check_duplicates=False, render_weight_threshold=render_weight_threshold)
    opt.optimize_affine_cascade(start_gear=MESH_GEAR_INITIAL, target_gear=MESH_GEAR_FIXED, svd_clip=(1, 1))
    opt.anneal(gear=(MESH_GEAR_FIXED, MESH_GEAR_MOVING), mode=_ANNEAL_COPY_EXACT)
    if linear:
        opt.optimize_linear(tol=1e-6, **opt_kwargs)
    else:
        opt.optimize_Newton_Raphson(max_newtonstep=5, tol=1e-4, **opt_kwargs)
    soft_avg = np.mean([m.soft_factor for m in opt.meshes])
    deformed = rigid = 0
    for m in opt.meshes:
        if (one_locked and not m.locked) or (not one_locked and m.soft_factor <= soft_avg):
            v_fixed = m.vertices(gear=MESH_GEAR_FIXED)
            move = m.vertices(gear=MESH_GEAR_MOVING) - v_fixed
            v_fixed = v_fixed - np.mean(v_fixed, axis=0, keepdims=True)
            move = move - np.mean(move, axis=0, keepdims=True)
            stiff, _ = m.stiffness_matrix()
            deformed += max(0, stiff.dot(move.ravel()).dot(move.ravel()))
            rigid += max(0, stiff.dot(v_fixed.ravel()).dot(v_fixed.ravel()))
    return (deformed / rigid) ** 0.5


# --------------------------------------------------------------------------------------------
# section and stitching entry points
# --------------------------------------------------------------------------------------------
def section_matcher(mesh0, mesh1, image_loader0, image_loader1, **kwargs):
    """Match two sections (feabas/matcher.py:370-427).  Returns ``(xy0, xy1, weight, strain)``."""
    initial_matches = kwargs.pop('initial_matches', None)
    spacings = kwargs.pop('spacings', [100])
    kwargs.setdefault('sigma', 2.5)
    kwargs.setdefault('batch_size', 100)
    kwargs.setdefault('distributor', 'cartesian_region')
    kwargs.setdefault('link_weight_decay', 0.0)
    compute_strain = kwargs.pop('compute_strain', False)
    stiff_thresh = kwargs.get('stiffness_multiplier_threshold', 0.1)
    kwargs.setdefault('render_weight_threshold', 0.1)
    stiffness_lambda = kwargs.setdefault('stiffness_lambda', 0.5)
    if stiff_thresh > 0 and hasattr(mesh0, 'triangle_mask_for_stiffness'):
        mesh0 = mesh0.submesh(mesh0.triangle_mask_for_stiffness(stiffness_multiplier_threshold=stiff_thresh))
        mesh1 = mesh1.submesh(mesh1.triangle_mask_for_stiffness(stiffness_multiplier_threshold=stiff_thresh))
    single = initial_matches is None or (mesh0.connected_triangles()[0] == 1 and mesh1.connected_triangles()[0] == 1)
    if single:
        return iterative_xcorr_matcher_w_mesh(mesh0, mesh1, image_loader0, image_loader1, spacings=spacings,
                                              initial_matches=initial_matches, compute_strain=compute_strain, **kwargs)
    # disconnected pieces are matched one pair at a time
    opt = _make_optimizer(mesh0, mesh1, stiffness_lambda)
    opt.add_link_from_coordinates(mesh0.uid, mesh1.uid, initial_matches.xy0, initial_matches.xy1,
                                  gear=(MESH_GEAR_INITIAL, MESH_GEAR_INITIAL), weight=initial_matches.weight, check_duplicates=False)
    opt.divide_disconnected_submeshes(prune_links=True)
    parts0, parts1, weights = [], [], []
    strain = DEFAULT_AVG_DEFORM
    for link in opt.links:
        sub0, sub1 = link.meshes
        seed = Match(link.xy0(gear=MESH_GEAR_INITIAL, use_mask=False, combine=True),
                     link.xy1(gear=MESH_GEAR_INITIAL, use_mask=False, combine=True), link.weight(use_mask=False))
        p0, p1, wt, strain = iterative_xcorr_matcher_w_mesh(sub0.copy(), sub1.copy(), image_loader0, image_loader1,
                                                            spacings=spacings, compute_strain=compute_strain,
                                                            initial_matches=seed, **kwargs)
        if p0 is None:
            continue
        same_order = (sub0.uid - sub1.uid) * (mesh0.uid - mesh1.uid) > 0
        parts0.append(p0 if same_order else p1)
        parts1.append(p1 if same_order else p0)
        weights.append(wt)
    if not parts0:
        return None, None, 0, DEFAULT_AVG_DEFORM
    return np.concatenate(parts0, axis=0), np.concatenate(parts1, axis=0), np.concatenate(weights, axis=0), strain


def _downsample(img, mask, factor, dev):
    t = _img.to_device(img, dev)
    if factor == 1:
        return t, (None if mask is None else _img.to_device(mask, t.device.index))
    small = _img.resize_area(t, factor)
    small_mask = None if mask is None else _img.resize_mask(_img.to_device(mask, t.device.index), factor)
    return small, small_mask


def _photometric(raw0, raw1, dog0, dog1, mask0, mask1, tx, ty, filtered):
    """Mean / spread of the two images inside their overlap (feabas/matcher.py:279-314); small reductions done
    with torch on the device tensors."""
    sx, sy = int(tx), int(ty)
    h0, w0 = dog0.shape
    h1, w1 = dog1.shape
    (x_lo, y_lo, x_hi, y_hi), _ = _blk.intersect_bbox((sx, sy, w0 + sx, h0 + sy), (0, 0, w1, h1))
    win0 = (slice(y_lo - sy, y_hi - sy), slice(x_lo - sx, x_hi - sx))
    win1 = (slice(y_lo, y_hi), slice(x_lo, x_hi))
    ones = torch.ones((max(y_hi - y_lo, 0), max(x_hi - x_lo, 0)), dtype=torch.bool, device=dog0.device)
    m0 = ones if mask0 is None else mask0[win0].to(torch.bool)
    m1 = ones if mask1 is None else mask1[win1].to(torch.bool)
    both = m0 & m1
    if int(m0.sum()) <= 3:
        return None
    if filtered:
        av0 = raw0[win0][both].to(torch.float64).mean().item()
        av1 = raw1[win1][both].to(torch.float64).mean().item()
        sd0 = dog0[win0][both].abs().mean().item()
        sd1 = dog1[win1][both].abs().mean().item()
    else:
        v0, v1 = dog0[win0][both].to(torch.float64), dog1[win1][both].to(torch.float64)
        av0, av1 = v0.mean().item(), v1.mean().item()
        sd0, sd1 = v0.std(unbiased=False).item(), v1.std(unbiased=False).item()
    return av0, av1, sd0, sd1


def stitching_matcher(img0, img1, **kwargs):
    """Displacement samples between two overlapping tile strips (feabas/matcher.py:224-367).

    Returns ``(xy0, xy1, weight, strain, phtm)`` -- points in each strip's own pixel frame -- or
    ``(None, None, conf_thresh, None, None)`` when the coarse translation is not trustworthy.
    The strips are uploaded once; resize, band-pass, block extraction and correlation run on the GPU.
    """
    sigma = kwargs.pop('sigma', 2.5)
    mask0 = kwargs.pop('mask0', None)
    mask1 = kwargs.pop('mask1', None)
    compute_photometric = kwargs.pop('compute_photometric', False)
    coarse = kwargs.pop('coarse_downsample', 1)
    fine = kwargs.pop('fine_downsample', 1)
    spacings = kwargs.pop('spacings', None)
    residue_len = kwargs.pop('residue_len', 5)
    dev = kwargs.pop('device', None)
    conf_mode = kwargs.get('conf_mode', FFT_CONF_MIRROR)
    conf_thresh = kwargs.get('conf_thresh', 0.3)
    min_num_blocks = kwargs.get('min_num_blocks', 2)
    kwargs.setdefault('residue_mode', 'huber')
    kwargs.setdefault('opt_tol', None)
    shape0 = tuple(img0.shape)
    shape1 = tuple(img1.shape)
    if spacings is None:
        spacings = _blk.auto_spacings(shape0, shape1)
    else:
        spacings = np.array(spacings, dtype=np.float64, copy=True).ravel()
    raw0, cmask0 = _downsample(img0, mask0, coarse, dev)
    dev = raw0.device.index
    raw1, cmask1 = _downsample(img1, mask1, coarse, dev)
    if sigma > 0:
        g0 = _masked_dog_any(raw0, sigma * coarse, cmask0)
        g1 = _masked_dog_any(raw1, sigma * coarse, cmask1)
    else:
        g0, g1 = raw0, raw1
    tx, ty, conf = global_translation_matcher(g0, g1, conf_mode=conf_mode, conf_thresh=conf_thresh)
    if conf < conf_thresh:
        return None, None, conf_thresh, None, None
    phtm = None
    if compute_photometric:
        phtm = _photometric(raw0, raw1, g0, g1, cmask0, cmask1, tx, ty, sigma > 0)
    if fine == coarse:
        f0, f1 = g0, g1
    else:
        f0, fmask0 = _downsample(img0, mask0, fine, dev)
        f1, fmask1 = _downsample(img1, mask1, fine, dev)
        if sigma > 0:
            f0 = _masked_dog_any(f0, sigma * fine, fmask0)
            f1 = _masked_dog_any(f1, sigma * fine, fmask1)
    tx = tx * fine / coarse
    ty = ty * fine / coarse
    resolution = _data_resolution() / fine
    residue_len = residue_len * fine
    loader0 = ArrayLoader(f0, fillval=0, resolution=resolution)
    loader1 = ArrayLoader(f1, fillval=0, resolution=resolution)
    if np.any(spacings < 1):
        box, _ = _blk.intersect_bbox(np.array(loader0.bounds) + np.tile((tx, ty), 2), loader1.bounds)
        spacings[spacings < 1] *= max(box[2] - box[0], box[3] - box[1])
    spacings = spacings * fine
    mesh0, mesh1 = _stitch_meshes(loader0, loader1, np.min(spacings), min_num_blocks)
    mesh0.apply_translation((tx, ty), MESH_GEAR_FIXED)
    mesh0.lock()
    xy0, xy1, weight, strain = iterative_xcorr_matcher_w_mesh(mesh0, mesh1, loader0, loader1, spacings=spacings,
                                                              distributor='cartesian_bbox', residue_len=residue_len, **kwargs)
    if fine != 1 and xy0 is not None:
        xy0 = _scale_coordinates(xy0, 1 / fine)
        xy1 = _scale_coordinates(xy1, 1 / fine)
    return xy0, xy1, weight, strain, phtm


def stitching_matcher_many(pairs, masks=None, chunk=256, **kwargs):
    """``stitching_matcher`` for MANY overlaps at once (the overlaps of a section: feabas/stitcher.py:385-394 hands
    them to worker processes one by one; here they advance together).

    ``pairs``: sequence of ``(img0, img1)`` strips; ``masks``: optional sequence of ``(mask0, mask1)`` (entries may be
    None); the keyword arguments are ``stitching_matcher``'s and apply to every overlap.  Returns the list of
    ``stitching_matcher`` results, in order, with the same numbers as one call per overlap would give:

    * strips of equal shape are uploaded, resized and band-passed as stacks and get their coarse translation from one
      batched cross-correlation;
    * the coarse-to-fine loops of all overlaps run in lockstep (``_drive_many``): per round, the blocks of every
      overlap that share a shape are gathered, correlated and read back together.
    """
    sigma = kwargs.pop('sigma', 2.5)
    compute_photometric = kwargs.pop('compute_photometric', False)
    coarse = kwargs.pop('coarse_downsample', 1)
    fine = kwargs.pop('fine_downsample', 1)
    spacings_arg = kwargs.pop('spacings', None)
    residue_len = kwargs.pop('residue_len', 5)
    dev = kwargs.pop('device', None)
    kwargs.pop('mask0', None), kwargs.pop('mask1', None)
    conf_mode = kwargs.get('conf_mode', FFT_CONF_MIRROR)
    conf_thresh = kwargs.get('conf_thresh', 0.3)
    min_num_blocks = kwargs.get('min_num_blocks', 2)
    kwargs.setdefault('residue_mode', 'huber')
    kwargs.setdefault('opt_tol', None)
    pairs = list(pairs)
    results = [None] * len(pairs)
    single_kw = dict(kwargs, sigma=sigma, compute_photometric=compute_photometric, coarse_downsample=coarse, fine_downsample=fine,
                     spacings=spacings_arg, residue_len=residue_len, device=dev)
    groups = {}
    for k, (img0, img1) in enumerate(pairs):
        m0, m1 = (None, None) if masks is None or masks[k] is None else masks[k]
        if m0 is not None or m1 is not None or img0.dtype != img1.dtype:
            # masked band-pass takes np.ptp per image (feabas/common.py:369): not stackable, one call
            results[k] = stitching_matcher(img0, img1, mask0=m0, mask1=m1, **dict(single_kw))
            continue
        groups.setdefault((tuple(img0.shape), tuple(img1.shape), str(img0.dtype)), []).append(k)
    loops, owners, finish = [], [], {}
    for (shape0, shape1, _), members in groups.items():
        for at in range(0, len(members), chunk):
            part = members[at:at + chunk]
            raw0 = _stack_to_device([pairs[k][0] for k in part], dev)
            dev = raw0.device.index
            raw1 = _stack_to_device([pairs[k][1] for k in part], dev)
            c0 = raw0 if coarse == 1 else _img.resize_area(raw0, coarse)
            c1 = raw1 if coarse == 1 else _img.resize_area(raw1, coarse)
            g0 = _img.masked_dog_device(c0, sigma * coarse) if sigma > 0 else c0
            g1 = _img.masked_dog_device(c1, sigma * coarse) if sigma > 0 else c1
            coarse_txy = _whole_image_translations(g0, g1, conf_mode)
            if fine == coarse:
                f0, f1 = g0, g1
            else:
                f0 = raw0 if fine == 1 else _img.resize_area(raw0, fine)
                f1 = raw1 if fine == 1 else _img.resize_area(raw1, fine)
                if sigma > 0:
                    f0, f1 = _img.masked_dog_device(f0, sigma * fine), _img.masked_dog_device(f1, sigma * fine)
            for j, k in enumerate(part):
                tx, ty, conf = coarse_txy[j]
                if not conf > conf_thresh:
                    tx, ty, conf = _translation_retry(g0[j], g1[j], tx, ty, conf, conf_mode, 6)
                if conf < conf_thresh:
                    results[k] = (None, None, conf_thresh, None, None)
                    continue
                phtm = _photometric(c0[j], c1[j], g0[j], g1[j], None, None, tx, ty, sigma > 0) if compute_photometric else None
                tx, ty = tx * fine / coarse, ty * fine / coarse
                resolution = _data_resolution() / fine
                loader0 = ArrayLoader(f0[j], fillval=0, resolution=resolution)
                loader1 = ArrayLoader(f1[j], fillval=0, resolution=resolution)
                if spacings_arg is None:
                    spacings = _blk.auto_spacings(shape0, shape1)
                else:
                    spacings = np.array(spacings_arg, dtype=np.float64, copy=True).ravel()
                if np.any(spacings < 1):
                    box, _ = _blk.intersect_bbox(np.array(loader0.bounds) + np.tile((tx, ty), 2), loader1.bounds)
                    spacings[spacings < 1] *= max(box[2] - box[0], box[3] - box[1])
                spacings = spacings * fine
                mesh0, mesh1 = _stitch_meshes(loader0, loader1, np.min(spacings), min_num_blocks)
                mesh0.apply_translation((tx, ty), MESH_GEAR_FIXED)
                mesh0.lock()
                loops.append(_coarse_to_fine(mesh0, mesh1, loader0, loader1, spacings=spacings, distributor='cartesian_bbox',
                                             residue_len=residue_len * fine, **kwargs))
                owners.append(k)
                finish[k] = phtm
    for k, (xy0, xy1, weight, strain) in zip(owners, _drive_many(loops)):
        if fine != 1 and xy0 is not None:
            xy0, xy1 = _scale_coordinates(xy0, 1 / fine), _scale_coordinates(xy1, 1 / fine)
        results[k] = (xy0, xy1, weight, strain, finish[k])
    return results


def section_matcher_many(jobs, **kwargs):
    """``section_matcher`` for many section pairs at once: ``jobs`` is a sequence of
    ``(mesh0, mesh1, image_loader0, image_loader1)`` (optionally a 5th entry: the pair's ``initial_matches``); the keyword
    arguments are ``section_matcher``'s.  The pairs' coarse-to-fine loops run in lockstep (``_drive_many``).  Returns the
    list of ``(xy0, xy1, weight, strain)``."""
    spacings = kwargs.pop('spacings', [100])
    kwargs.pop('initial_matches', None)
    kwargs.setdefault('sigma', 2.5)
    kwargs.setdefault('batch_size', 100)
    kwargs.setdefault('distributor', 'cartesian_region')
    kwargs.setdefault('link_weight_decay', 0.0)
    compute_strain = kwargs.pop('compute_strain', False)
    stiff_thresh = kwargs.get('stiffness_multiplier_threshold', 0.1)
    kwargs.setdefault('render_weight_threshold', 0.1)
    kwargs.setdefault('stiffness_lambda', 0.5)
    results, loops, owners = [None] * len(jobs), [], []
    for k, job in enumerate(jobs):
        mesh0, mesh1, loader0, loader1 = job[:4]
        initial = job[4] if len(job) > 4 else None
        if stiff_thresh > 0 and hasattr(mesh0, 'triangle_mask_for_stiffness'):
            mesh0 = mesh0.submesh(mesh0.triangle_mask_for_stiffness(stiffness_multiplier_threshold=stiff_thresh))
            mesh1 = mesh1.submesh(mesh1.triangle_mask_for_stiffness(stiffness_multiplier_threshold=stiff_thresh))
        if initial is not None and not (mesh0.connected_triangles()[0] == 1 and mesh1.connected_triangles()[0] == 1):
            results[k] = section_matcher(mesh0, mesh1, loader0, loader1, spacings=spacings, initial_matches=initial,
                                         compute_strain=compute_strain, **dict(kwargs, stiffness_multiplier_threshold=0))
            continue
        loops.append(_coarse_to_fine(mesh0, mesh1, loader0, loader1, spacings=spacings, initial_matches=initial,
                                     compute_strain=compute_strain, **kwargs))
        owners.append(k)
    for k, out in zip(owners, _drive_many(loops)):
        results[k] = out
    return results


def _stack_to_device(images, device):
    """Equal-shaped 2-D images (numpy arrays or tensors) -> one contiguous ``N x H x W`` CUDA tensor."""
    if all(isinstance(x, np.ndarray) for x in images):
        return _img.to_device(np.stack(images, axis=0), device)
    return torch.stack([_img.to_device(x, device) for x in images], dim=0).contiguous()


def _scale_coordinates(xy, scale):
    """Pixel-centre preserving rescale (feabas/spatial.py:77-89)."""
    return (np.asarray(xy) + 0.5) * scale - 0.5


_DATA_RESOLUTION_FN = False                             # False: not looked up yet; None: FEABAS is not importable


def _data_resolution():
    """``feabas.config.data_resolution()`` when FEABAS is installed (looked up once: a failing import costs
    ~0.2 ms per call, as much as a small xcorr batch)."""
    global _DATA_RESOLUTION_FN
    if _DATA_RESOLUTION_FN is False:
        try:
            from feabas.config import data_resolution   # pragma: no cover
            _DATA_RESOLUTION_FN = data_resolution       # pragma: no cover
        except Exception:
            _DATA_RESOLUTION_FN = None
    return _DATA_RESOLUTION_FN() if _DATA_RESOLUTION_FN is not None else DEFAULT_RESOLUTION


_MESH_FACTORY = None


def set_mesh_factory(factory):
    """``factory(bounds, mesh_size, min_num_blocks, uid, resolution) -> mesh``.  ``None`` restores the default, the
    affine stand-in ``AffineMesh.from_bbox``.  A FEABAS installation plugs its elastic meshes in with
    ``set_mesh_factory(lambda b, s, n, uid, res: Mesh.from_bbox(b, cartesian=True, mesh_size=s, min_num_blocks=n,
    uid=uid, resolution=res))`` (feabas/matcher.py:354-359)."""
    global _MESH_FACTORY
    _MESH_FACTORY = factory


def _stitch_meshes(loader0, loader1, mesh_size, min_num_blocks):
    factory = _MESH_FACTORY
    if factory is None:
        factory = _default_mesh_factory()
    return (factory(loader0.bounds, mesh_size, min_num_blocks, 0, loader0.resolution),
            factory(loader1.bounds, mesh_size, min_num_blocks, 1, loader1.resolution))


def _default_mesh_factory():
    def affine(bounds, mesh_size, min_num_blocks, uid, resolution):
        return AffineMesh.from_bbox(bounds, cartesian=True, mesh_size=mesh_size, min_num_blocks=min_num_blocks, uid=uid,
                                    resolution=resolution)
    return affine
