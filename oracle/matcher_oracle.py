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


# --------------------------------------------------------------------------- #
# block rendering through an affine map (cv2.remap, as the reference does)
# --------------------------------------------------------------------------- #
def resize_area_oracle(img, factor):
    """matcher.py:254-256 -- the very cv2 call of the reference."""
    import cv2
    return cv2.resize(img, None, fx=factor, fy=factor, interpolation=cv2.INTER_AREA)


def resize_mask_oracle(mask, factor):
    """matcher.py:257-264."""
    import cv2
    return cv2.resize(mask.astype(np.uint8), None, fx=factor, fy=factor, interpolation=cv2.INTER_NEAREST).astype(bool)


def render_blocks_oracle(img, bboxes, ainv, tinv, fillval=0, origin_xy=(0, 0), cover=None):
    """``MeshRenderer.crop_multiple`` for ONE affine map (renderer.py:601-648 with ``crop_field_affine``
    renderer.py:419-450 for every block) over an in-RAM image whose pixel (0, 0) sits at ``origin_xy``
    (``dal.StreamLoader``), rendered by ``common.render_by_subregions`` (common.py:256-350): fields of all
    blocks concatenated, source crop = floor(min) - 4 .. ceil(max) + 4, ``cv2.remap(INTER_LINEAR,
    BORDER_CONSTANT, fillval)``.  ``cover``: (xmin, ymin, xmax, ymax), the renderer's covered region (the mesh
    shrunk by half a pixel, renderer.py:98-101) in source pixels, used as ``crop_field_affine(precise_mask=True)``
    does (renderer.py:436-449): a block whose footprint (bbox - 0.5, mapped) sticks out by less than one square
    pixel is rendered whole, otherwise only pixels strictly inside the region.  Returns (stack N x H x W, mask)."""
    import cv2
    from . import convex
    fx, fy, masks = [], [], []
    for b in np.asarray(bboxes, dtype=np.float64).reshape(-1, 4):
        wd, ht = round(b[2] - b[0]), round(b[3] - b[1])
        xs = np.linspace(b[0], b[2], num=wd, endpoint=False, dtype=float)
        ys = np.linspace(b[1], b[3], num=ht, endpoint=False, dtype=float)
        xx, yy = np.meshgrid(xs, ys)
        bx = xx * ainv[0, 0] + yy * ainv[1, 0] + tinv[0] - origin_xy[0]
        by = xx * ainv[0, 1] + yy * ainv[1, 1] + tinv[1] - origin_xy[1]
        fx.append(bx)
        fy.append(by)
        if cover is None:
            masks.append(np.ones(bx.shape, dtype=bool))
            continue
        foot = convex.affine_transform(convex.box(*(b - 0.5)), (ainv[0, 0], ainv[1, 0], ainv[0, 1], ainv[1, 1],
                                                               tinv[0] - origin_xy[0], tinv[1] - origin_xy[1]))
        part = convex.box(*cover).intersection(foot)
        if foot.area - part.area < 1:
            masks.append(np.ones(bx.shape, dtype=bool))
        else:
            masks.append(convex.contains_xy(part, bx, by))
    nblk = len(fx)
    map_x, map_y, mask = np.concatenate(fx, 0), np.concatenate(fy, 0), np.concatenate(masks, 0)
    out = np.full(map_x.shape, fillval, dtype=img.dtype)
    if mask.any():
        x_lo, x_hi = np.floor(map_x[mask].min()) - 4, np.ceil(map_x[mask].max()) + 4
        y_lo, y_hi = np.floor(map_y[mask].min()) - 4, np.ceil(map_y[mask].max()) + 4
        src = crop_with_fill(img, (int(x_lo), int(y_lo), int(x_hi), int(y_hi)), fillval)
        warped = cv2.remap(src, (map_x - x_lo).astype(np.float32), (map_y - y_lo).astype(np.float32),
                           interpolation=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=fillval)
        out[mask] = warped[mask]
    return out.reshape(nblk, -1, out.shape[-1]), mask.reshape(nblk, -1, out.shape[-1])


# --------------------------------------------------------------------------- #
# coarse-to-fine loop with an affine section model (surrogate relaxation)
# --------------------------------------------------------------------------- #
class _Section:
    """One affine map instead of an elastic mesh: moving = initial @ a + t."""

    def __init__(self, bounds, locked=False):
        self.bounds = bounds
        self.a, self.t = np.eye(2), np.zeros(2)
        self.locked = locked

    def bbox(self):
        x0, y0, x1, y1 = self.bounds
        c = np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], dtype=np.float64) @ self.a + self.t
        return np.array([c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()])

    def fwd(self, p):
        return p @ self.a + self.t

    def inv(self, p):
        return (p - self.t) @ np.linalg.inv(self.a)

    def sampler(self):
        ai = np.linalg.inv(self.a)
        return ai, -self.t @ ai


def _wls_affine(src, dst, w):
    if src.shape[0] < 3 or np.linalg.matrix_rank(src - src.mean(0)) < 2:
        return np.eye(2), (np.average(dst - src, axis=0, weights=w) if w.sum() > 0 else np.zeros(2))
    sw = np.sqrt(w)[:, None]
    design = np.concatenate((src, np.ones((src.shape[0], 1))), axis=1)
    sol = np.linalg.lstsq(design * sw, dst * sw, rcond=None)[0]
    return sol[:2], sol[2]


def surrogate_loop_oracle(sec0, sec1, img0, img1, spacings, conf_thresh=0.3, residue_mode='huber', residue_len=0,
                          min_num_blocks=2, pad=None, subpixel=None, batch_size=None, conf_mode=FFT_CONF_MIRROR,
                          xcorr=xcorr_oracle, trace=None):
    """Control flow of ``iterative_xcorr_matcher_w_mesh`` (matcher.py:567-751) for distributor
    'cartesian_bbox', sec0 locked, ``allow_enlarge=False``, ``allow_dwell=0``, ``max_spacing_skip=0``,
    ``link_weight_decay=0``, sigma=0, with the affine relaxation model of feabas_b200.cuda.surrogate."""
    spacings = np.sort(np.asarray(spacings, dtype=np.float64))[::-1]
    sp, idx, started = spacings[0], 0, False
    use_pad = True if pad is None else pad
    link = None
    while idx < spacings.size:
        finest = sp == spacings[-1]
        sub = finest if subpixel is None else subpixel
        box, ok = intersect_bbox_oracle(sec0.bbox(), sec1.bbox())
        if not ok:
            return None, None, 0
        b0, b1 = cartesian_blocks_oracle(sec0.bbox(), sec1.bbox(), sp, min_num_blocks=min_num_blocks if finest else 1, zorder=True)
        n = b0.shape[0]
        z0 = np.round(bbox_sizes_oracle(b0))
        brk = np.nonzero(np.any(np.diff(z0, axis=0), axis=-1))[0]
        edges = np.concatenate(([0], brk + 1, [n]), axis=None)
        if batch_size is not None and batch_size < n:
            parts = [np.linspace(lo, hi, num=max(1, int(np.ceil((hi - lo) / batch_size))) + 1, endpoint=True)
                     for lo, hi in zip(edges[:-1], edges[1:])]
            edges = np.unique(np.round(np.concatenate(parts, axis=-1)).astype(np.int32))
        p0, p1, cf = [], [], []
        for lo, hi in zip(edges[:-1], edges[1:]):
            st0, _ = render_blocks_oracle(img0, b0[lo:hi], *sec0.sampler())
            st1, _ = render_blocks_oracle(img1, b1[lo:hi], *sec1.sampler())
            dx, dy, c = xcorr(st0, st1, conf_mode=conf_mode, pad=use_pad, subpixel=sub)
            q0, q1 = block_points_oracle(b0[lo:hi], b1[lo:hi], dx, dy)
            p0.append(q0), p1.append(q1), cf.append(c)
        p0, p1, cf = np.concatenate(p0, 0), np.concatenate(p1, 0), np.concatenate(cf, 0)
        if trace is not None:
            trace.append(dict(spacing=float(sp), pad=bool(use_pad), subpixel=bool(sub), nblocks=int(n), conf=cf.copy(),
                              xy0=p0.copy(), xy1=p1.copy()))
        if np.all(cf <= conf_thresh):
            if not started:
                return None, None, 0
            break
        good = cf > conf_thresh
        p0, p1, wt = p0[good], p1[good], cf[good].astype(np.float64)
        max_dis = np.max(np.sum((p0 - p1) ** 2, axis=-1)) ** 0.5
        nxt = np.searchsorted(-spacings, -4 * max_dis) - 1
        if nxt > idx:
            nxt = min(nxt, idx + 1)
            if pad is None:
                use_pad = False          # matcher.py:701-705: adjacent level
            idx = nxt
        else:
            if pad is None:
                use_pad = True
            idx += 1
        link = dict(i0=sec0.inv(p0), i1=sec1.inv(p1), w=wt, rw=np.ones_like(wt))
        if max_dis > 0.1:
            def relax():
                a, t = _wls_affine(sec1.fwd(link['i1']), sec0.fwd(link['i0']), link['w'] * link['rw'])
                sec1.a, sec1.t = sec1.a @ a, sec1.t @ a + t
            relax()
            if residue_len > 0:
                r = np.sum((sec1.fwd(link['i1']) - sec0.fwd(link['i0'])) ** 2, axis=-1) ** 0.5
                if residue_mode == 'huber':
                    new = np.where(r > residue_len, residue_len / np.maximum(r, 1e-30), 1.0)
                else:
                    new = (r <= residue_len).astype(np.float64)
                changed = np.any(np.abs(new - link['rw']) > 1e-3)
                link['rw'] = new
                if changed and idx < spacings.size:
                    relax()
        started = True
        if idx < spacings.size:
            sp = spacings[idx]
    if link is None:
        return None, None, 0
    keep = (link['w'] * link['rw']) > 0
    return link['i0'][keep], link['i1'][keep], (link['w'] * link['rw'])[keep]


def stitching_oracle(img0, img1, sigma=2.5, coarse_downsample=1, fine_downsample=1, spacings=None, residue_len=5,
                     conf_thresh=0.3, min_num_blocks=2, pad=None, residue_mode='huber', xcorr=xcorr_oracle, trace=None):
    """``stitching_matcher`` (matcher.py:224-367) without masks / photometric output, on the affine section model."""
    if spacings is None:
        spacings = auto_spacings_oracle(img0.shape, img1.shape)
    spacings = np.array(spacings, dtype=np.float64)
    g0 = resize_area_oracle(img0, coarse_downsample) if coarse_downsample != 1 else img0
    g1 = resize_area_oracle(img1, coarse_downsample) if coarse_downsample != 1 else img1
    if sigma > 0:
        g0 = masked_dog_oracle(g0, sigma * coarse_downsample)
        g1 = masked_dog_oracle(g1, sigma * coarse_downsample)
    tx, ty, cf = global_translation_oracle(g0, g1, conf_thresh=conf_thresh, xcorr=xcorr)
    if trace is not None:
        trace.append(dict(coarse=(tx, ty, cf)))
    if cf < conf_thresh:
        return None, None, conf_thresh
    if fine_downsample == coarse_downsample:
        f0, f1 = g0, g1
    else:
        f0 = resize_area_oracle(img0, fine_downsample) if fine_downsample != 1 else img0
        f1 = resize_area_oracle(img1, fine_downsample) if fine_downsample != 1 else img1
        if sigma > 0:
            f0 = masked_dog_oracle(f0, sigma * fine_downsample)
            f1 = masked_dog_oracle(f1, sigma * fine_downsample)
    tx, ty = tx * fine_downsample / coarse_downsample, ty * fine_downsample / coarse_downsample
    # Mesh.from_bbox(cartesian=True) puts the outer vertices at bounds - 0.5 (mesh.py:426-427)
    sec0 = _Section((-0.5, -0.5, f0.shape[1] - 0.5, f0.shape[0] - 0.5), locked=True)
    sec1 = _Section((-0.5, -0.5, f1.shape[1] - 0.5, f1.shape[0] - 0.5))
    sec0.t = np.array([tx, ty], dtype=np.float64)
    xy0, xy1, wt = surrogate_loop_oracle(sec0, sec1, f0, f1, spacings * fine_downsample, conf_thresh=conf_thresh,
                                         residue_mode=residue_mode, residue_len=residue_len * fine_downsample,
                                         min_num_blocks=min_num_blocks, pad=pad, xcorr=xcorr, trace=trace)
    if xy0 is not None and fine_downsample != 1:
        xy0 = (xy0 + 0.5) / fine_downsample - 0.5       # spatial.scale_coordinates, spatial.py:77-89
        xy1 = (xy1 + 0.5) / fine_downsample - 0.5
    return xy0, xy1, wt
