"""CPU oracle for the matcher layers around ``xcorr_fft``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates, with citations:

* ``common.masked_dog_filter``        reference ``feabas/common.py:353-377``
* ``common.divide_bbox``              reference ``feabas/common.py:380-409``
* ``common.intersect_bbox``           reference ``feabas/common.py:412-417``
* ``common.z_order``                  reference ``feabas/common.py:196-215``
* ``common.bbox_centers/bbox_sizes``  reference ``feabas/common.py:687-696``
* ``global_translation_matcher``      reference ``feabas/matcher.py:138-221``
* ``distributor_cartesian_bbox``      reference ``feabas/matcher.py:865-891``
* block -> point-pair conversion of ``bboxes_mesh_renderer_matcher``
  (``feabas/matcher.py:781-861``) for translation-only meshes over in-RAM images
  (``dal.StreamLoader.crop``, ``feabas/dal.py:1045-1050``): blocks are integer
  shifted crops with ``fillval`` outside the image.
* INTER_AREA 0.5x resize of ``stitching_matcher`` (``feabas/matcher.py:254-256``)
  via ``cv2.resize`` itself (third-party, same call).

Pinned by ``tests/golden/matcher_*.npz`` (outputs of the unmodified reference).
"""
import numpy as np
from scipy.ndimage import gaussian_filter1d

from .xcorr_oracle import xcorr_oracle, FFT_CONF_MIRROR


# --------------------------------------------------------------------------- #
# band-pass filter
# --------------------------------------------------------------------------- #
def _gauss2(x, s):
    x = gaussian_filter1d(x, s, axis=-1, mode='nearest')
    return gaussian_filter1d(x, s, axis=-2, mode='nearest')


def masked_dog_oracle(img, sigma, mask=None, signed=True, ptp=None):
    """common.py:353-377.  ``ptp`` overrides the stack-global ``np.ptp(img)``
    (the reference always uses the ptp of the whole array it is handed)."""
    img = np.asarray(img)
    if not np.issubdtype(img.dtype, np.floating):
        img = img.astype(np.float32)
    g1 = _gauss2(img, sigma)
    g2 = _gauss2(g1, sigma)
    out = g1 - g2
    if mask is not None and not np.all(mask):
        span = np.ptp(img) if ptp is None else ptp
        outside = span * (np.asarray(mask) == 0)
        s_c = (2.0 * sigma * sigma) ** 0.5
        bleed = _gauss2(outside, s_c) * (s_c ** 2) / (sigma ** 2)
        out = (np.abs(out) - bleed).clip(0, None) * np.sign(out)
    if not signed:
        out = np.abs(out)
    return out


# --------------------------------------------------------------------------- #
# bbox helpers
# --------------------------------------------------------------------------- #
def divide_bbox_oracle(bbox, block_size=None, min_num_blocks=1, round_output=True, shrink_factor=1):
    """common.py:380-409 -> (xmin, ymin, xmax, ymax) flat arrays, row-major over (y, x)."""
    x0, y0, x1, y1 = bbox
    ht, wd = y1 - y0, x1 - x0
    if block_size is None:
        block_size = max(ht, wd)
    bs = block_size if hasattr(block_size, '__len__') else (block_size, block_size)
    mn = min_num_blocks if hasattr(min_num_blocks, '__len__') else (min_num_blocks, min_num_blocks)
    ncol = max(np.ceil(wd / bs[1]), mn[1])
    nrow = max(np.ceil(ht / bs[0]), mn[0])
    bw = int(np.ceil(wd / ncol))
    bh = int(np.ceil(ht / nrow))
    xs = np.linspace(x0, x1 - bw, num=int(ncol), endpoint=True)
    ys = np.linspace(y0, y1 - bh, num=int(nrow), endpoint=True)
    if shrink_factor != 1:
        bw2, bh2 = bw * shrink_factor, bh * shrink_factor
        xs = xs + (bw - bw2) / 2
        ys = ys + (bh - bh2) / 2
        bw, bh = int(np.ceil(bw2)), int(np.ceil(bh2))
    if round_output:
        xs = np.round(xs).astype(np.int32)
        ys = np.round(ys).astype(np.int32)
    gx, gy = np.meshgrid(xs, ys)
    gx, gy = gx.ravel(), gy.ravel()
    return gx, gy, gx + bw, gy + bh


def intersect_bbox_oracle(b0, b1):
    lo_x, lo_y = max(b0[0], b1[0]), max(b0[1], b1[1])
    hi_x, hi_y = min(b0[2], b1[2]), min(b0[3], b1[3])
    return (lo_x, lo_y, hi_x, hi_y), (lo_x < hi_x) and (lo_y < hi_y)


def z_order_oracle(indices, base=2):
    """common.py:196-215 -- permutation sorting integer grid indices in Morton order."""
    idx = np.asarray(indices)
    nd = idx.shape[-1]
    idx = idx - idx.min(axis=0)
    code = np.zeros_like(idx)
    level = 0
    while np.any(idx > 0):
        code = code + (idx % base) * (base ** (nd * level))
        idx = np.floor(idx / base)
        level += 1
    score = np.sum(code * (base ** np.arange(nd)), axis=-1)
    return np.argsort(score, kind='stable')


def bbox_centers_oracle(bboxes):
    b = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    return np.stack(((b[:, 0] + b[:, 2]) * 0.5 - 0.5, (b[:, 1] + b[:, 3]) * 0.5 - 0.5), axis=-1)


def bbox_sizes_oracle(bboxes):
    """(height, width) per bbox, clipped at zero (common.py:693-696)."""
    b = np.asarray(bboxes, dtype=np.float64).reshape(-1, 4)
    return np.stack((b[:, 3] - b[:, 1], b[:, 2] - b[:, 0]), axis=-1).clip(0, None)


def cartesian_blocks_oracle(bbox0, bbox1, spacing, min_num_blocks=1, shrink_factor=1, zorder=False):
    """matcher.py:865-891 with the two mesh bboxes (MOVING gear) given directly."""
    sf = shrink_factor if hasattr(shrink_factor, '__len__') else (shrink_factor, shrink_factor)
    box, ok = intersect_bbox_oracle(bbox0, bbox1)
    if not ok:
        return None, None
    a = np.stack(divide_bbox_oracle(box, block_size=spacing, min_num_blocks=min_num_blocks,
                                    shrink_factor=sf[0]), axis=-1)
    b = np.stack(divide_bbox_oracle(box, block_size=spacing, min_num_blocks=min_num_blocks,
                                    shrink_factor=sf[1]), axis=-1)
    if zorder:
        gx = np.round((a[:, 0] - a[:, 0].min()) / spacing)
        gy = np.round((a[:, 1] - a[:, 1].min()) / spacing)
        order = z_order_oracle(np.stack((gx, gy), axis=-1))
        a, b = a[order], b[order]
    return a, b


# --------------------------------------------------------------------------- #
# global translation
# --------------------------------------------------------------------------- #
def _balanced_division(shape_hw, factor):
    # matcher.py:162-177
    if hasattr(factor, '__len__'):
        return tuple(factor[:2])
    r_img = shape_hw[0] / shape_hw[1]
    best, pick = np.inf, None
    for r in range(1, int(factor ** 0.5) + 1):
        if factor % r:
            continue
        for cand, score in (((int(factor / r), int(r)), abs(np.log(r_img * (r * r / factor)))),
                            ((int(r), int(factor / r)), abs(np.log(r_img / (r * r / factor))))):
            if score < best:
                best, pick = score, cand
    return pick


def _grown_window(lo, hi, want, limit):
    # matcher.py:189-194: symmetric growth to `want`, then pushed back inside [0, limit]
    grow = int(np.ceil((want - (hi - lo)) / 2))
    a, b = lo - grow, hi + grow
    shift = -min(a, 0) - max(b - limit, 0)
    return int(np.clip(a + shift, 0, limit)), int(np.clip(b + shift, 0, limit))


def global_translation_oracle(img0, img1, conf_mode=FFT_CONF_MIRROR, conf_thresh=0.3, divide_factor=6,
                              xcorr=xcorr_oracle):
    """matcher.py:138-221 with ``sigma == 0``.  ``xcorr`` is injectable so a
    test can drive the same control flow with the device kernel."""
    ht0, wd0 = img0.shape[-2:]
    ht1, wd1 = img1.shape[-2:]
    tx, ty, cf = xcorr(img0[None], img1[None], conf_mode=conf_mode, pad=True)
    tx, ty, cf = tx.item(), ty.item(), cf.item()
    tx += (wd1 - wd0) / 2
    ty += (ht1 - ht0) / 2
    if cf > conf_thresh:
        return tx, ty, cf
    div = _balanced_division(np.minimum((ht0, wd0), (ht1, wd1)), divide_factor)
    ax0, ay0, bx0, by0 = divide_bbox_oracle((0, 0, wd0, ht0), min_num_blocks=div)
    ax1, ay1, bx1, by1 = divide_bbox_oracle((0, 0, wd1, ht1), min_num_blocks=div)
    s0, s1, offx, offy = [], [], [], []
    for k in range(ax0.size):
        bw = max(bx0[k] - ax0[k], bx1[k] - ax1[k])
        bh = max(by0[k] - ay0[k], by1[k] - ay1[k])
        ya0, yb0 = _grown_window(ay0[k], by0[k], bh, ht0)
        xa0, xb0 = _grown_window(ax0[k], bx0[k], bw, wd0)
        blk0 = img0[ya0:yb0, xa0:xb0]
        if np.ptp(blk0) == 0:
            continue
        ya1, yb1 = _grown_window(ay1[k], by1[k], bh, ht1)
        xa1, xb1 = _grown_window(ax1[k], bx1[k], bw, wd1)
        blk1 = img1[ya1:yb1, xa1:xb1]
        if np.ptp(blk1) == 0:
            continue
        s0.append(blk0)
        s1.append(blk1)
        offx.append(((xb1 - xa1) - (xb0 - xa0)) / 2 + xa1 - xa0)
        offy.append(((yb1 - ya1) - (yb0 - ya0)) / 2 + ya1 - ya0)
    if not s0:
        return tx, ty, cf
    bx, by, bc = xcorr(np.stack(s0, 0), np.stack(s1, 0), conf_mode=conf_mode, pad=True)
    bx = bx + np.array(offx)
    by = by + np.array(offy)
    kb = int(np.argmax(bc))
    if bc[kb] >= cf:
        tx, ty, cf = bx[kb], by[kb], bc[kb]
    return tx, ty, cf


# --------------------------------------------------------------------------- #
# block grid over in-RAM images with translation-only meshes
# --------------------------------------------------------------------------- #
def crop_with_fill(img, bbox, fillval=0):
    """dal.StreamLoader.crop -> common.crop_image_from_bbox: bbox = (xmin, ymin,
    xmax, ymax) in image pixel coordinates, right/bottom exclusive, ``fillval``
    outside the image."""
    x0, y0, x1, y1 = (int(v) for v in bbox)
    out = np.full((y1 - y0, x1 - x0), fillval, dtype=img.dtype)
    sy0, sy1 = max(y0, 0), min(y1, img.shape[0])
    sx0, sx1 = max(x0, 0), min(x1, img.shape[1])
    if sy1 > sy0 and sx1 > sx0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = img[sy0:sy1, sx0:sx1]
    return out


def block_points_oracle(bboxes0, bboxes1, dx, dy):
    """matcher.py:840-849 (Appendix C of SURVEY.md)."""
    c0 = bbox_centers_oracle(bboxes0)
    c1 = bbox_centers_oracle(bboxes1)
    z0 = bbox_sizes_oracle(bboxes0)
    z1 = bbox_sizes_oracle(bboxes1)
    w = (z0 / (z0 + z1))[:, ::-1]
    d = np.stack((dx, dy), axis=-1)
    return c0 - d * w, c1 + d * (1 - w)


def block_grid_match_oracle(img0, img1, bboxes0, bboxes1, shift0=(0, 0), shift1=(0, 0), batch_size=None,
                            conf_mode=FFT_CONF_MIRROR, pad=True, subpixel=False, xcorr=xcorr_oracle):
    """``bboxes_mesh_renderer_matcher`` (matcher.py:781-861) when mesh ``k`` is
    the identity mesh translated by the INTEGER vector ``shift_k`` (so a block
    with bbox ``b`` in the common frame samples image ``k`` at ``b - shift_k``;
    that is the state of ``stitching_matcher``'s first pass at every level-0,
    matcher.py:354-362) and images are served by ``StreamLoader(fillval=0)``.
    Batch splitting follows matcher.py:804-822."""
    b0 = np.asarray(bboxes0)
    b1 = np.asarray(bboxes1)
    n = b0.shape[0]
    z0 = np.round(bbox_sizes_oracle(b0))
    z1 = np.round(bbox_sizes_oracle(b1))
    brk = np.nonzero(np.any(np.diff(z0, axis=0), axis=-1) | np.any(np.diff(z1, axis=0), axis=-1))[0]
    edges = np.concatenate(([0], brk + 1, [n]), axis=None)
    if batch_size is not None and batch_size < n:
        parts = []
        for lo, hi in zip(edges[:-1], edges[1:]):
            nb = max(1, int(np.ceil((hi - lo) / batch_size)))
            parts.append(np.linspace(lo, hi, num=nb + 1, endpoint=True))
        edges = np.unique(np.round(np.concatenate(parts, axis=-1)).astype(np.int32))
    p0, p1, cf = [], [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        if hi <= lo:
            continue
        st0 = np.stack([crop_with_fill(img0, b - np.tile(shift0, 2)) for b in b0[lo:hi]], 0)
        st1 = np.stack([crop_with_fill(img1, b - np.tile(shift1, 2)) for b in b1[lo:hi]], 0)
        dx, dy, c = xcorr(st0, st1, conf_mode=conf_mode, pad=pad, subpixel=subpixel)
        q0, q1 = block_points_oracle(b0[lo:hi], b1[lo:hi], dx, dy)
        p0.append(q0)
        p1.append(q1)
        cf.append(c)
    if not p0:
        return np.empty((0, 2)), np.empty((0, 2)), np.empty(0)
    return np.concatenate(p0, 0), np.concatenate(p1, 0), np.concatenate(cf, 0)


def auto_spacings_oracle(shape0, shape1):
    """stitching_matcher's default block spacings (matcher.py:243-251)."""
    shp = np.minimum(shape0, shape1)
    smx = max(shp) * 0.25
    smn = max(min(75, min(shp) / 3), 25)
    if smn > smx:
        return np.array([smn])
    nsp = max(1, round(np.log(smx / smn) / np.log(4)))
    return np.exp(np.linspace(np.log(smn), np.log(smx), num=nsp, endpoint=True))
