"""Device image operators around the matcher, with the reference's call conventions.

* ``masked_dog_filter``   feabas/common.py:353-377
* ``resize_area``         ``cv2.resize(img, None, fx=f, fy=f, interpolation=cv2.INTER_AREA)``, f = 1/k
                          (feabas/matcher.py:254-256)
* ``resize_mask``         ``cv2.resize(mask.astype(uint8), ..., INTER_NEAREST).astype(bool)`` (matcher.py:257-264)
* ``crop_blocks``         affine block gather (feabas/renderer.py:419-450,601-648; feabas/common.py:256-350)

NumPy in -> NumPy out (host arrays are uploaded, results downloaded); CUDA tensor in -> CUDA
tensor out on the current stream, nothing synchronised.  torch only provides memory and streams.
"""
import math

import numpy as np

from . import _lib

try:
    import torch
except Exception:                       # pragma: no cover
    torch = None


def _need_torch():
    if torch is None:
        raise _lib.FeabasCudaError('feabas_b200.cuda needs torch for device memory')


def is_cuda_tensor(x):
    return torch is not None and isinstance(x, torch.Tensor) and x.is_cuda


def to_device(x, device=None, dtype=None):
    """numpy array / CPU tensor / CUDA tensor -> contiguous CUDA tensor (no copy if already there)."""
    _need_torch()
    if not torch.cuda.is_available():
        _lib.lib()                                             # a missing .so is reported first
        raise _lib.FeabasCudaError('no CUDA device available (feabas_b200 has no CPU fallback)')
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if not t.is_cuda:
        dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
        t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def upload_small(arr, device):
    """Small host array (block rows, index lists; a few KB) -> CUDA tensor on the current stream, straight from pageable
    memory: the driver embeds such copies in the command stream, so they do not queue up on the copy engine behind
    the whole-section uploads of ``ArrayLoader`` (staging them through page-locked memory was measured: the same GPU
    idle gaps, and 18 % less end-to-end throughput on the config-4 block pass, because a pinned copy is a DMA)."""
    return torch.from_numpy(np.ascontiguousarray(arr)).to(device, non_blocking=True)


_UPLOAD_STREAMS = {}


def upload_stream(dev):
    """One side stream per device for host -> device copies of whole images (ArrayLoader)."""
    key = dev.index
    if key not in _UPLOAD_STREAMS:
        _UPLOAD_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _UPLOAD_STREAMS[key]


def _code(t):
    if t.dtype == torch.float32:
        return _lib.FB_F32
    if t.dtype == torch.uint8:
        return _lib.FB_U8
    raise TypeError(f'unsupported image dtype {t.dtype} (float32 / uint8)')


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def stack_minmax(stack):
    """Per-image (min, max) of an ``N x ...`` CUDA tensor -> ``N x 2`` float32 CUDA tensor."""
    n = stack.shape[0]
    out = torch.empty((n, 2), dtype=torch.float32, device=stack.device)
    if n:
        _lib.check(_lib.lib().fb_stack_minmax(stack.data_ptr(), n, stack[0].numel(), _code(stack), out.data_ptr(),
                                              stack.device.index, _stream(stack)))
    return out


def masked_dog_device(img, sigma, mask=None, signed=True, ptp=None, exact=False, out=None, mask_images=None):
    """``img``: CUDA tensor ``(..., H, W)`` float32 or uint8; ``mask``: None or CUDA tensor broadcastable to
    ``img`` (nonzero = keep) that is known NOT to be all true.  ``mask_images``: int32 CUDA tensor, one entry per image of
    ``mask``: the image of the stack that mask belongs to (``mask`` then only covers the images that have masked pixels;
    ``fb_masked_dog_sparse``).  Returns float32, same shape."""
    shape = img.shape
    h, w = shape[-2:]
    n = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    img = img.contiguous()
    if img.dtype == torch.float64:
        # common.py:363-364 converts only non-floating dtypes: a float64 stack stays float64 throughout
        if mask_images is not None:
            raise ValueError('mask_images is a float32 / uint8 feature')
        return _masked_dog_f64(img, n, h, w, sigma, mask, signed, ptp, out)
    if img.dtype not in (torch.float32, torch.uint8):
        # every non-floating dtype is converted to float32; float16 follows scipy, which filters it as float64 and
        # hands back float16 -- no caller does that, refuse it rather than guess
        if img.dtype.is_floating_point:
            raise TypeError('masked_dog_filter: %s images are not supported (float32 / float64 / integer)' % img.dtype)
        img = img.to(torch.float32)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=img.device)
    mask_n, mptr, iptr = 1, None, None
    if mask is not None and mask_images is not None:
        m = mask if mask.dtype == torch.uint8 else (mask != 0).to(torch.uint8)
        m = m.reshape(-1, h, w).contiguous()
        mask_n, mptr, iptr = m.shape[0], m.data_ptr(), mask_images.data_ptr()
    elif mask is not None:
        m = mask
        if m.dtype != torch.uint8:
            m = (m != 0).to(torch.uint8)
        if m.dim() > 2 and int(np.prod(m.shape[:-2])) > 1:
            m = m.expand(shape).contiguous()
            mask_n = n
        else:
            m = m.reshape(h, w).contiguous()
        mptr = m.data_ptr()
    L = _lib.lib()
    wb = L.fb_masked_dog_workspace(n, h, w)
    work = torch.empty(wb, dtype=torch.uint8, device=img.device)
    flags = (0 if signed else _lib.FB_DOG_UNSIGNED) | (_lib.FB_DOG_EXACT if exact else 0)
    _lib.check(L.fb_masked_dog_sparse(img.data_ptr(), mptr, iptr, n, h, w, _code(img), mask_n, float(sigma),
                                      float('nan') if ptp is None else float(ptp), flags, out.data_ptr(), work.data_ptr(), wb,
                                      img.device.index, _stream(img)))
    return out


def _masked_dog_f64(img, n, h, w, sigma, mask, signed, ptp, out):
    """float64 branch of ``masked_dog_device`` (``fb_masked_dog_f64``): float64 in, float64 out."""
    if out is None:
        out = torch.empty(img.shape, dtype=torch.float64, device=img.device)
    mask_n, mptr = 1, None
    if mask is not None:
        m = mask if mask.dtype == torch.uint8 else (mask != 0).to(torch.uint8)
        if m.dim() > 2 and int(np.prod(m.shape[:-2])) > 1:
            m = m.expand(img.shape).contiguous()
            mask_n = n
        else:
            m = m.reshape(h, w).contiguous()
        mptr = m.data_ptr()
    L = _lib.lib()
    wb = L.fb_masked_dog_f64_workspace(n, h, w)
    work = torch.empty(wb, dtype=torch.uint8, device=img.device)
    _lib.check(L.fb_masked_dog_f64(img.data_ptr(), mptr, n, h, w, mask_n, float(sigma), float('nan') if ptp is None else float(ptp),
                                   0 if signed else _lib.FB_DOG_UNSIGNED, out.data_ptr(), work.data_ptr(), wb,
                                   img.device.index, _stream(img)))
    return out


def masked_dog_filter(img, sigma, mask=None, signed=True, **kwargs):
    """Drop-in for ``feabas.common.masked_dog_filter`` (common.py:353-377).

    Extra keyword arguments (not in the reference): ``ptp`` -- the value of ``np.ptp`` the mask term
    should use (the reference takes it over whatever array it is handed, so a caller that splits a
    stack across GPUs passes the stack-global value); ``exact`` -- float64 accumulation in scipy's
    order (reference rounding) instead of float32; ``device``.
    """
    ptp = kwargs.get('ptp', None)
    exact = kwargs.get('exact', False)
    on_gpu = is_cuda_tensor(img)
    if mask is not None:
        # common.py:368: an all-true mask is the same as no mask
        if bool(mask.all()):
            mask = None
    t = to_device(img, kwargs.get('device', None))
    m = None if mask is None else to_device(mask, t.device.index)
    out = masked_dog_device(t, sigma, m, signed=signed, ptp=ptp, exact=exact)
    return out if on_gpu else out.cpu().numpy()


def _round_half_even(v):
    return int(np.rint(v))


_DBL_EPSILON = 2.220446049250313e-16


def resize_area(img, factor, **kwargs):
    """``cv2.resize(img, None, fx=factor, fy=factor, interpolation=cv2.INTER_AREA)`` for any ``factor <= 1``
    (matcher.py:254-256,321-322); ``img``: ``(..., H, W)`` uint8 or float32, bit-exact for both.  OpenCV has two
    code paths and so does this: k x k cell means when ``1/factor`` is an integer to within DBL_EPSILON
    (``fb_resize_area``), per-axis coverage tables otherwise (``fb_resize_area_frac``)."""
    if not factor > 0:
        raise ValueError(f'resize factor {factor}')
    scale = 1.0 / float(factor)
    if scale < 1:
        # enlarging: OpenCV switches INTER_AREA to a bilinear variant; no FEABAS configuration asks for it
        raise NotImplementedError(f'INTER_AREA resize only shrinks (got factor {factor})')
    k = int(np.rint(scale))
    fast = abs(scale - k) < _DBL_EPSILON
    on_gpu = is_cuda_tensor(img)
    t = to_device(img, kwargs.get('device', None))
    if fast and k == 1:
        return t.clone() if on_gpu else np.array(img, copy=True)
    shape = t.shape
    h, w = shape[-2:]
    n = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    oh, ow = _round_half_even(h * factor), _round_half_even(w * factor)   # cv2: saturate_cast<int>(size * fx)
    if oh < 1 or ow < 1:
        raise ValueError(f'resize of {h}x{w} by {factor} leaves no pixel')      # (cv2 raises as well)
    if (oh, ow) == (h, w):
        return t.clone() if on_gpu else np.array(img, copy=True)               # cv2::resize: equal sizes are a plain copy
    out = torch.empty(tuple(shape[:-2]) + (oh, ow), dtype=t.dtype, device=t.device)
    if fast:
        _lib.check(_lib.lib().fb_resize_area(t.data_ptr(), n, h, w, _code(t), k, out.data_ptr(), oh, ow, t.device.index, _stream(t)))
    else:
        _lib.check(_lib.lib().fb_resize_area_frac(t.data_ptr(), n, h, w, _code(t), scale, scale, out.data_ptr(), oh, ow,
                                                  t.device.index, _stream(t)))
    return out if on_gpu else out.cpu().numpy()


def resize_mask(mask, factor, **kwargs):
    """Nearest-neighbour resize of a boolean mask, OpenCV's index rule."""
    on_gpu = is_cuda_tensor(mask)
    t = to_device(mask, kwargs.get('device', None))
    t = (t != 0).to(torch.uint8).contiguous()
    shape = t.shape
    h, w = shape[-2:]
    n = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    oh, ow = _round_half_even(h * factor), _round_half_even(w * factor)
    out = torch.empty(tuple(shape[:-2]) + (oh, ow), dtype=torch.uint8, device=t.device)
    _lib.check(_lib.lib().fb_resize_nearest(t.data_ptr(), n, h, w, 1.0 / factor, 1.0 / factor, out.data_ptr(), oh, ow,
                                            t.device.index, _stream(t)))
    out = out.to(torch.bool)
    return out if on_gpu else out.cpu().numpy()


def crop_blocks_masked(img, blocks, block_shape, origin=None, fillval=0, cover=None, out=None, compact=False):
    """Gather ``len(blocks)`` blocks of ``block_shape = (bh, bw)`` from the 2-D CUDA tensor ``img``.

    ``blocks``: ``N x 10`` float64 (numpy or CUDA): x0, y0, step_x, step_y, A00, A10, t0, A01, A11, t1
    (see ``fb_crop_blocks`` in include/feabas_cuda.h).  ``origin``: the integer (x, y) origin OpenCV's
    fixed-point coordinates are taken from; default = what the reference computes for the batch,
    ``floor(min field) - 4`` (common.py:300-304).  ``cover``: (xmin, ymin, xmax, ymax) of the source region
    the mesh covers (``AffineMesh.covered_rect`` in image pixel coordinates), or None.  The reference's rule
    (feabas/renderer.py:436-449): a block whose footprint sticks out of the covered region by less than one square
    pixel is rendered whole, otherwise only the pixels whose source position is strictly inside the region.
    Returns ``(stack, mask)``; ``mask`` (uint8, 1 = rendered) is None when every block is rendered whole.  With
    ``compact=True`` the third value ``mask_images`` (int32 CUDA tensor) lists the blocks that are only partly covered and
    ``mask`` holds one image per entry of it (the layout ``masked_dog_device(..., mask_images=...)`` takes); the fourth,
    ``any_covered``, is False when no pixel of the whole batch can be covered."""
    bh, bw = int(block_shape[0]), int(block_shape[1])
    blk_host = None
    if not is_cuda_tensor(blocks):
        blk_host = np.ascontiguousarray(blocks, dtype=np.float64).reshape(-1, 10)
        blocks = upload_small(blk_host, img.device)
    n = blocks.shape[0]
    if out is None:
        out = torch.empty((n, bh, bw), dtype=img.dtype, device=img.device)
    mask, cptr, mptr, fptr = None, None, None, None
    whole, mask_images, any_covered = None, None, True
    if cover is not None and n:
        if blk_host is None:
            blk_host = blocks.cpu().numpy()
        uncovered = footprint_uncovered_area(blk_host, bh, bw, cover)
        whole = uncovered < 1
        full_area = np.abs(blk_host[:, 4] * blk_host[:, 8] - blk_host[:, 5] * blk_host[:, 7]) * (bw * blk_host[:, 2]) * (bh * blk_host[:, 3])
        any_covered = bool(np.any(uncovered < full_area * (1 - 1e-9)))
    if origin is None:
        if blk_host is None:
            blk_host = blocks.cpu().numpy()
        # the reference takes the origin of its source crop from the RENDERED pixels (common.py:300-304): every pixel
        # of a block rendered whole, only the covered ones otherwise
        origin = batch_origin(blk_host, bh, bw, None if whole is None else ~whole, cover) if n else (0, 0)
    if whole is not None and whole.all():
        cover = None            # every block is rendered whole: same stack, no mask to write / reduce / read back
    if cover is not None:
        import ctypes
        carr = (ctypes.c_double * 4)(*[float(v) for v in cover])
        cptr = ctypes.cast(carr, ctypes.c_void_p)
        if compact:
            part = np.nonzero(~whole)[0].astype(np.int32)
            slot = np.full(n, -1, dtype=np.int32)
            slot[part] = np.arange(part.size, dtype=np.int32)
            both = upload_small(np.concatenate((slot, part)), img.device)
            slots, mask_images = both[:n], both[n:]
            mask = torch.empty((part.size, bh, bw), dtype=torch.uint8, device=img.device)
            fptr = slots.data_ptr()
        else:
            mask = torch.empty((n, bh, bw), dtype=torch.uint8, device=img.device)
            if n and whole.any():
                slots = upload_small(np.where(whole, -1, np.arange(n)).astype(np.int32), img.device)
                fptr = slots.data_ptr()
        mptr = mask.data_ptr()
    if n:
        ih, iw = img.shape
        _lib.check(_lib.lib().fb_crop_blocks(img.data_ptr(), ih, iw, _code(img), blocks.data_ptr(), n, bh, bw,
                                             float(origin[0]), float(origin[1]), float(fillval), out.data_ptr(),
                                             cptr, fptr, mptr, img.device.index, _stream(img)))
    if not compact and mask is not None and whole is not None and whole.any():
        mask[upload_small(np.nonzero(whole)[0], img.device)] = 1          # (whole blocks: the kernel wrote no mask)
    if compact:
        return out, mask, mask_images, any_covered
    return out, mask


def _clip_many(pts, cnt, axis, bound, keep_greater):
    """One Sutherland-Hodgman step for MANY convex polygons at once: ``pts`` is ``B x V x 2`` with the first ``cnt[b]``
    vertices of polygon b valid; keeps the part with coordinate ``axis`` on the kept side of ``bound``.  Returns the
    clipped polygons in the same representation (``V`` grows by one per step at most)."""
    nb, nv, _ = pts.shape
    idx = np.arange(nv)[None, :]
    valid = idx < cnt[:, None]
    nxt_i = np.where(idx + 1 < cnt[:, None], idx + 1, 0)
    nxt = np.take_along_axis(pts, nxt_i[:, :, None].repeat(2, axis=2), axis=1)
    d = (pts[:, :, axis] - bound) if keep_greater else (bound - pts[:, :, axis])
    dn = np.take_along_axis(d, nxt_i, axis=1)
    keep = valid & (d >= 0)
    cross = valid & ((d >= 0) != (dn >= 0))
    with np.errstate(divide='ignore', invalid='ignore'):
        frac = np.where(cross, d / (d - dn), 0.0)
    inter = pts + (nxt - pts) * frac[:, :, None]
    # candidates in polygon order: vertex i (if kept), then the crossing on edge i -> i + 1
    cand = np.stack((pts, inter), axis=2).reshape(nb, 2 * nv, 2)
    flag = np.stack((keep, cross), axis=2).reshape(nb, 2 * nv)
    order = np.argsort(~flag, axis=1, kind='stable')
    out_n = flag.sum(axis=1)
    width = min(2 * nv, nv + 1)
    sel = order[:, :width]
    out = np.take_along_axis(cand, sel[:, :, None].repeat(2, axis=2), axis=1)
    return out, out_n


def _poly_areas(pts, cnt):
    """Shoelace areas of many polygons in the ``_clip_many`` representation."""
    nb, nv, _ = pts.shape
    idx = np.arange(nv)[None, :]
    valid = idx < cnt[:, None]
    nxt_i = np.where(idx + 1 < cnt[:, None], idx + 1, 0)
    nxt = np.take_along_axis(pts, nxt_i[:, :, None].repeat(2, axis=2), axis=1)
    cross = np.where(valid, pts[:, :, 0] * nxt[:, :, 1] - nxt[:, :, 0] * pts[:, :, 1], 0.0)
    return np.where(cnt >= 3, 0.5 * np.abs(cross.sum(axis=1)), 0.0)


def footprint_uncovered_area(blocks, bh, bw, cover):
    """Per block: area of its footprint in the source image that lies outside ``cover``.

    The footprint is the block's bounding box grown to pixel edges, ``bbox - 0.5``, mapped through the block's
    affine map (feabas/renderer.py:437-442); the reference renders a block whole when less than one square pixel
    of it is uncovered (:443-444).  Blocks entirely inside report exactly 0 without any clipping; the others are
    clipped against the four sides of ``cover`` together (vectorised Sutherland-Hodgman)."""
    b = np.asarray(blocks, dtype=np.float64).reshape(-1, 10)
    x_lo, y_lo = b[:, 0] - 0.5, b[:, 1] - 0.5
    x_hi, y_hi = x_lo + bw * b[:, 2], y_lo + bh * b[:, 3]
    cx = np.stack((x_lo, x_hi, x_hi, x_lo), axis=-1)
    cy = np.stack((y_lo, y_lo, y_hi, y_hi), axis=-1)
    qx = cx * b[:, 4, None] + cy * b[:, 5, None] + b[:, 6, None]
    qy = cx * b[:, 7, None] + cy * b[:, 8, None] + b[:, 9, None]
    inside = (qx.min(axis=1) >= cover[0]) & (qx.max(axis=1) <= cover[2]) & (qy.min(axis=1) >= cover[1]) & (qy.max(axis=1) <= cover[3])
    out = np.zeros(b.shape[0], dtype=np.float64)
    todo = np.nonzero(~inside)[0]
    if todo.size:
        pts = np.stack((qx[todo], qy[todo]), axis=-1)                      # K x 4 x 2
        cnt = np.full(todo.size, 4, dtype=np.int64)
        whole = _poly_areas(pts, cnt)
        for axis, bound, greater in ((0, cover[0], True), (0, cover[2], False), (1, cover[1], True), (1, cover[3], False)):
            pts, cnt = _clip_many(pts, cnt, axis, bound, greater)
        out[todo] = whole - _poly_areas(pts, cnt)
    return out


_SRC_DTYPE = np.dtype([('img', np.uint64), ('ih', np.int32), ('iw', np.int32), ('ox', np.float64), ('oy', np.float64)])   # fb_crop_src


def crop_blocks_multi(parts, block_shape, fillval=0):
    """Blocks of MANY source images in one launch (``fb_crop_blocks_multi``).  ``parts``: list of ``(img, blocks)``
    -- a 2-D CUDA tensor and the ``K x 10`` float64 rows (numpy) of the blocks to cut from it; every part is one batch
    of the reference (its source-crop origin, ``floor(min field) - 4``, is taken over the part).  All images share a
    dtype and a device.  Returns the stack ``sum(K) x bh x bw``, parts in order."""
    bh, bw = int(block_shape[0]), int(block_shape[1])
    img0 = parts[0][0]
    row_list = [np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 10) for _, b in parts]
    counts = np.array([r.shape[0] for r in row_list], dtype=np.int64)
    rows = np.concatenate(row_list, axis=0)
    n = rows.shape[0]
    src = np.empty(n, dtype=_SRC_DTYPE)
    for img, _ in parts:
        if img.dtype != img0.dtype or img.device != img0.device or img.dim() != 2 or not img.is_contiguous():
            raise ValueError('crop_blocks_multi: the source images must be contiguous 2-D tensors of one dtype on one device')
    if n:
        live = counts > 0
        starts = (np.cumsum(counts) - counts)[live]
        ox, oy = batch_origins(rows, bh, bw, starts)                      # one origin per part, all parts at once
        reps = counts[live]
        src['img'] = np.repeat(np.array([img.data_ptr() for (img, _), ok in zip(parts, live) if ok], dtype=np.uint64), reps)
        src['ih'] = np.repeat(np.array([img.shape[0] for (img, _), ok in zip(parts, live) if ok], dtype=np.int32), reps)
        src['iw'] = np.repeat(np.array([img.shape[1] for (img, _), ok in zip(parts, live) if ok], dtype=np.int32), reps)
        src['ox'], src['oy'] = np.repeat(ox, reps), np.repeat(oy, reps)
    out = torch.empty((n, bh, bw), dtype=img0.dtype, device=img0.device)
    if n:
        packed = np.concatenate((rows.view(np.uint8).reshape(-1), src.view(np.uint8).reshape(-1)))    # one upload
        dev_buf = upload_small(packed, img0.device)
        rows_ptr = dev_buf.data_ptr()
        _lib.check(_lib.lib().fb_crop_blocks_multi(rows_ptr + rows.nbytes, _code(img0), rows_ptr, n, bh, bw, float(fillval),
                                                  out.data_ptr(), img0.device.index, _stream(img0)))
    return out


def crop_blocks(img, blocks, block_shape, origin=None, fillval=0, out=None):
    """``crop_blocks_masked`` without a coverage region; returns the stack only."""
    return crop_blocks_masked(img, blocks, block_shape, origin=origin, fillval=fillval, out=out)[0]


def _block_field_minima(b, bh, bw):
    """Per block: minimum of the (affine) coordinate field over the block -- it sits at a corner."""
    xe = np.stack((b[:, 0], b[:, 0] + (bw - 1) * b[:, 2]), axis=-1)[:, :, None]     # N x 2 x 1
    ye = np.stack((b[:, 1], b[:, 1] + (bh - 1) * b[:, 3]), axis=-1)[:, None, :]     # N x 1 x 2
    xs = xe * b[:, 4, None, None] + ye * b[:, 5, None, None] + b[:, 6, None, None]
    ys = xe * b[:, 7, None, None] + ye * b[:, 8, None, None] + b[:, 9, None, None]
    return xs.reshape(b.shape[0], -1).min(axis=1), ys.reshape(b.shape[0], -1).min(axis=1)


def batch_origins(rows, bh, bw, starts):
    """``batch_origin`` of many batches at once: ``rows`` holds the batches one after the other, batch k starts at row
    ``starts[k]`` (increasing, first 0).  Returns two float64 arrays (x, y), one entry per batch."""
    x_min, y_min = _block_field_minima(np.asarray(rows, dtype=np.float64).reshape(-1, 10), bh, bw)
    starts = np.asarray(starts, dtype=np.int64)
    return np.floor(np.minimum.reduceat(x_min, starts)) - 4, np.floor(np.minimum.reduceat(y_min, starts)) - 4


def batch_origin(blocks, bh, bw, partial=None, cover=None):
    """floor(min of the batch's coordinate field) - 4 per axis (common.py:300-304): the field is affine, so
    its extrema sit at block corners.  ``partial`` (bool per block) marks blocks of which only the pixels inside
    ``cover`` are rendered: their contribution to the minimum cannot lie below the cover's lower edge."""
    b = np.asarray(blocks, dtype=np.float64).reshape(-1, 10)
    x_min, y_min = _block_field_minima(b, bh, bw)
    if partial is not None and cover is not None and np.any(partial):
        x_min = np.where(partial, np.maximum(x_min, cover[0]), x_min)
        y_min = np.where(partial, np.maximum(y_min, cover[1]), y_min)
    return math.floor(x_min.min()) - 4, math.floor(y_min.min()) - 4
