"""Block geometry of the matcher (host side, NumPy; microseconds per call).

Mirrors, with the reference's argument meaning and return conventions:

* ``divide_bbox``                  feabas/common.py:380-409
* ``intersect_bbox``               feabas/common.py:412-417
* ``z_order``                      feabas/common.py:196-215
* ``bbox_centers``/``bbox_sizes``  feabas/common.py:687-696
* ``distributor_cartesian_bbox``   feabas/matcher.py:865-891
* ``split_batches``                the batch partition of feabas/matcher.py:804-822
* ``block_points``                 the (dx, dy) -> point-pair conversion of feabas/matcher.py:840-849
* ``auto_spacings``                feabas/matcher.py:243-251

Bounding boxes are ``[xmin, ymin, xmax, ymax]`` with right / bottom exclusive.
"""
import numpy as np

from .constant import MESH_GEAR_MOVING


def _pair(v):
    """scalar -> (v, v); sequence -> its first two entries, (row / y, column / x) order."""
    if hasattr(v, '__len__'):
        return v[0], v[1]
    return v, v


def divide_bbox(bbox, **kwargs):
    """Tile ``bbox`` with equal blocks of at most ``block_size`` (at least ``min_num_blocks`` per
    axis), evenly spread so that the first and last block touch the borders.

    Returns four flat arrays ``(xmin, ymin, xmax, ymax)``, blocks ordered row-major (y outer).
    """
    block_size = kwargs.get('block_size', None)
    min_num_blocks = kwargs.get('min_num_blocks', 1)
    round_output = kwargs.get('round_output', True)
    shrink_factor = kwargs.get('shrink_factor', 1)
    x_lo, y_lo, x_hi, y_hi = bbox
    height, width = y_hi - y_lo, x_hi - x_lo
    if block_size is None:
        block_size = max(height, width)
    size_y, size_x = _pair(block_size)
    least_y, least_x = _pair(min_num_blocks)
    count_y = max(np.ceil(height / size_y), least_y)
    count_x = max(np.ceil(width / size_x), least_x)
    step_y = int(np.ceil(height / count_y))
    step_x = int(np.ceil(width / count_x))
    starts_x = _linspace(x_lo, x_hi - step_x, int(count_x))
    starts_y = _linspace(y_lo, y_hi - step_y, int(count_y))
    if shrink_factor != 1:
        small_x, small_y = step_x * shrink_factor, step_y * shrink_factor
        starts_x = starts_x + (step_x - small_x) / 2
        starts_y = starts_y + (step_y - small_y) / 2
        step_x, step_y = int(np.ceil(small_x)), int(np.ceil(small_y))
    if round_output:
        starts_x = np.round(starts_x).astype(np.int32)
        starts_y = np.round(starts_y).astype(np.int32)
    grid_x, grid_y = np.tile(starts_x, starts_y.size), np.repeat(starts_y, starts_x.size)      # meshgrid(...).ravel()
    return grid_x, grid_y, grid_x + step_x, grid_y + step_y


def _linspace(start, stop, num):
    """``np.linspace(start, stop, num, endpoint=True)`` (float64) without its argument handling -- the same operations in
    the same order: ``arange(num) * step + start`` with ``step = (stop - start) / (num - 1)``, the last sample set to
    ``stop``; the block grid is built thousands of times per section."""
    start, stop = float(start), float(stop)
    if num == 1:
        return np.array([start])
    delta = stop - start
    step = delta / (num - 1)
    y = np.arange(num, dtype=np.float64)
    y = y * step if step != 0 else y / (num - 1) * delta
    y += start
    y[-1] = stop
    return y


def intersect_bbox(bbox0, bbox1):
    """-> ((xmin, ymin, xmax, ymax), valid)."""
    lo_x, lo_y = max(bbox0[0], bbox1[0]), max(bbox0[1], bbox1[1])
    hi_x, hi_y = min(bbox0[2], bbox1[2]), min(bbox0[3], bbox1[3])
    return (lo_x, lo_y, hi_x, hi_y), bool(lo_x < hi_x and lo_y < hi_y)


def z_order(indices, base=2):
    """Permutation that sorts integer-valued grid indices (N x d) along a Morton curve."""
    grid = np.asarray(indices)
    ndim = grid.shape[-1]
    grid = grid - grid.min(axis=0)
    if base == 2 and ndim == 2 and grid.size and float(grid.max()) < 2 ** 31 and np.all(grid == np.floor(grid)):
        # the digit loop below for the usual case, as bit interleaving on integers: the same integer score
        # (digit d of axis a lands on bit 2 d + a), hence the same stable order
        g = grid.astype(np.int64)
        for shift, maskbits in ((16, 0x0000FFFF0000FFFF), (8, 0x00FF00FF00FF00FF), (4, 0x0F0F0F0F0F0F0F0F),
                                (2, 0x3333333333333333), (1, 0x5555555555555555)):
            g = (g | (g << shift)) & maskbits
        return np.argsort(g[:, 0] + 2 * g[:, 1], kind='stable')
    key = np.zeros_like(grid)
    digit = 0
    while np.any(grid > 0):
        key = key + (grid % base) * (base ** (ndim * digit))
        grid = np.floor(grid / base)
        digit += 1
    score = np.sum(key * (base ** np.arange(ndim)), axis=-1)
    return np.argsort(score, kind='stable')


def bbox_centers(bboxes):
    b = np.asarray(bboxes).reshape(-1, 4)
    return np.stack(((b[:, 0] + b[:, 2]) / 2 - 0.5, (b[:, 1] + b[:, 3]) / 2 - 0.5), axis=-1)


def bbox_sizes(bboxes):
    """(height, width) per box, negative extents clipped to zero."""
    b = np.asarray(bboxes).reshape(-1, 4)
    return np.stack((b[:, 3] - b[:, 1], b[:, 2] - b[:, 0]), axis=-1).clip(0, None)


def distributor_cartesian_bbox(mesh0, mesh1, spacing, **kwargs):
    """Cartesian block grid on the intersection of the two meshes' bounding boxes.
    ``mesh.bbox(gear=...)`` is the only thing asked of the mesh objects."""
    gear = kwargs.get('gear', MESH_GEAR_MOVING)
    min_num_blocks = kwargs.get('min_num_blocks', 1)
    shrink0, shrink1 = _pair(kwargs.get('shrink_factor', 1))
    zorder = kwargs.get('zorder', False)
    common_box, valid = intersect_bbox(mesh0.bbox(gear=gear), mesh1.bbox(gear=gear))
    if not valid:
        return None, None
    boxes0 = np.stack(divide_bbox(common_box, block_size=spacing, min_num_blocks=min_num_blocks, shrink_factor=shrink0), axis=-1)
    if shrink1 == shrink0:                               # (the usual call: one grid for both sides)
        boxes1 = boxes0.copy()
    else:
        boxes1 = np.stack(divide_bbox(common_box, block_size=spacing, min_num_blocks=min_num_blocks, shrink_factor=shrink1), axis=-1)
    if zorder:
        col = np.round((boxes0[:, 0] - boxes0[:, 0].min()) / spacing)
        row = np.round((boxes0[:, 1] - boxes0[:, 1].min()) / spacing)
        order = z_order(np.stack((col, row), axis=-1))
        boxes0, boxes1 = boxes0[order], boxes1[order]
    return boxes0, boxes1


def split_batches(bboxes0, bboxes1, batch_size=None):
    """Index edges of the xcorr batches: a new batch starts wherever the (rounded) block size of
    either side changes, and runs longer than ``batch_size`` are cut into near-equal pieces."""
    count = bboxes0.shape[0]
    if count and (batch_size is None or batch_size >= count):
        # the usual call: one grid of equal blocks, no batch limit -- the general code below returns [0, count] for it
        s0 = np.round(bboxes0[:, 2:4] - bboxes0[:, 0:2])
        s1 = np.round(bboxes1[:, 2:4] - bboxes1[:, 0:2])
        if (s0 == s0[0]).all() and (s1 == s1[0]).all() and (s0[0] > 0).all() and (s1[0] > 0).all():
            return np.array([0, count])
    size0 = np.round(bbox_sizes(bboxes0))
    size1 = np.round(bbox_sizes(bboxes1))
    change = np.any(np.diff(size0, axis=0), axis=-1) | np.any(np.diff(size1, axis=0), axis=-1)
    edges = np.concatenate(([0], np.nonzero(change)[0] + 1, [count]), axis=None)
    if batch_size is None or batch_size >= count:
        return edges
    pieces = []
    for lo, hi in zip(edges[:-1], edges[1:]):
        cuts = max(1, int(np.ceil((hi - lo) / batch_size)))
        pieces.append(np.linspace(lo, hi, num=cuts + 1, endpoint=True))
    return np.unique(np.round(np.concatenate(pieces, axis=-1)).astype(np.int32))


def block_points(bboxes0, bboxes1, dx, dy):
    """Block displacement -> matched point pair: the displacement is shared between the two block
    centres in proportion to the block sizes."""
    size0, size1 = bbox_sizes(bboxes0), bbox_sizes(bboxes1)
    share = (size0 / (size0 + size1))[:, ::-1]            # (h, w) -> (x, y)
    shift = np.stack((dx, dy), axis=-1)
    return bbox_centers(bboxes0) - shift * share, bbox_centers(bboxes1) + shift * (1 - share)


def auto_spacings(shape0, shape1):
    """Default pyramid of ``stitching_matcher``: geometric from ~75 px to a quarter of the long side."""
    shape = np.minimum(shape0, shape1)
    coarse = max(shape) * 0.25
    fine = max(min(75, min(shape) / 3), 25)
    if fine > coarse:
        return np.array([fine])
    levels = max(1, round(np.log(coarse / fine) / np.log(4)))
    return np.exp(np.linspace(np.log(fine), np.log(coarse), num=levels, endpoint=True))


def balanced_division(shape_hw, divide_factor):
    """(rows, cols) with rows * cols == divide_factor whose blocks are closest to square
    (feabas/matcher.py:162-177); a sequence is taken as given."""
    if hasattr(divide_factor, '__len__'):
        return tuple(divide_factor[:2])
    aspect = shape_hw[0] / shape_hw[1]
    best, choice = np.inf, None
    for r in range(1, int(divide_factor ** 0.5) + 1):
        if divide_factor % r:
            continue
        q = r * r / divide_factor
        for cand, score in (((int(divide_factor / r), int(r)), abs(np.log(aspect * q))),
                            ((int(r), int(divide_factor / r)), abs(np.log(aspect / q)))):
            if score < best:
                best, choice = score, cand
    return choice
