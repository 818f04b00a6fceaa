"""CPU tier: host-side block geometry of feabas_b200.cuda.blocks against golden vectors produced by the
unmodified reference (tests/golden/matcher_host.npz) and against the oracle on random input."""
import numpy as np
import pytest

from feabas_b200.cuda import blocks as bk
from feabas_b200.cuda.constant import MESH_GEAR_MOVING
from oracle import matcher_oracle as mo


def _same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    np.testing.assert_array_equal(a, b)


class _Box:
    def __init__(self, b):
        self._b = b

    def bbox(self, gear=None):
        assert gear == MESH_GEAR_MOVING
        return self._b


def test_divide_bbox_golden(golden_host):
    i = 0
    while f'divide/{i}/bbox' in golden_host:
        kw = {k.split('/')[-1]: (v if v.ndim else v.item()) for k, v in golden_host.items() if k.startswith(f'divide/{i}/kw/')}
        if 'min_num_blocks' in kw and np.ndim(kw['min_num_blocks']):
            kw['min_num_blocks'] = tuple(kw['min_num_blocks'])
        _same(np.stack(bk.divide_bbox(tuple(golden_host[f'divide/{i}/bbox']), **kw), 0), golden_host[f'divide/{i}/out'])
        i += 1
    assert i == 5


def test_zorder_and_bbox_helpers_golden(golden_host):
    _same(bk.z_order(golden_host['zorder/in']), golden_host['zorder/out'])
    _same(bk.bbox_centers(golden_host['bbox/in']), golden_host['bbox/centers'])
    _same(bk.bbox_sizes(golden_host['bbox/in']), golden_host['bbox/sizes'])


def test_cartesian_distributor_golden(golden_host):
    i = 0
    while f'cart/{i}/bbox0' in golden_host:
        kw = {k.split('/')[-1]: v.item() for k, v in golden_host.items() if k.startswith(f'cart/{i}/kw/')}
        o0, o1 = bk.distributor_cartesian_bbox(_Box(golden_host[f'cart/{i}/bbox0']), _Box(golden_host[f'cart/{i}/bbox1']),
                                               golden_host[f'cart/{i}/spacing'].item(), **kw)
        _same(o0, golden_host[f'cart/{i}/out0']), _same(o1, golden_host[f'cart/{i}/out1'])
        i += 1
    assert i == 3
    assert bk.distributor_cartesian_bbox(_Box((0, 0, 10, 10)), _Box((20, 20, 30, 30)), 5) == (None, None)


def test_random_against_oracle():
    rng = np.random.default_rng(5)
    for _ in range(50):
        x0, y0 = rng.integers(-500, 500, 2)
        w, h = rng.integers(30, 4000, 2)
        sp = float(rng.uniform(20, 900))
        kw = dict(block_size=sp, min_num_blocks=int(rng.integers(1, 4)), shrink_factor=float(rng.choice([1, 0.7, 0.5])))
        _same(np.stack(bk.divide_bbox((x0, y0, x0 + w, y0 + h), **kw)), np.stack(mo.divide_bbox_oracle((x0, y0, x0 + w, y0 + h), **kw)))
        b0 = (x0, y0, x0 + w, y0 + h)
        b1 = (x0 + int(rng.integers(-20, 20)), y0 + int(rng.integers(-20, 20)), x0 + w, y0 + h)
        got = bk.distributor_cartesian_bbox(_Box(b0), _Box(b1), sp, min_num_blocks=2, zorder=True)
        want = mo.cartesian_blocks_oracle(b0, b1, sp, min_num_blocks=2, zorder=True)
        _same(got[0], want[0]), _same(got[1], want[1])
        dx, dy = rng.standard_normal((2, got[0].shape[0]))
        p = bk.block_points(got[0], got[1], dx, dy)
        q = mo.block_points_oracle(got[0], got[1], dx, dy)
        _same(p[0], q[0]), _same(p[1], q[1])


def test_fast_paths_of_the_block_grid_are_the_reference_arithmetic():
    """The block grid is built thousands of times per section: ``_linspace`` is ``np.linspace`` operation for operation,
    ``divide_bbox`` / ``z_order`` on float-valued boxes (what the matcher passes: mesh bounding boxes after a translation)
    equal the oracle bit for bit, ``split_batches``' shortcut equals its general code."""
    rng = np.random.default_rng(0)
    for it in range(3000):
        a, b = rng.uniform(-1e4, 1e4, 2)
        if it % 50 == 0:
            b = a
        n = int(rng.integers(1, 60))
        _same(bk._linspace(a, b, n), np.linspace(a, b, num=n, endpoint=True))
    for it in range(1500):
        x0, y0 = rng.uniform(-500, 500, 2) if it % 2 else rng.integers(-500, 500, 2)
        w, h = rng.uniform(30, 4000, 2) if it % 3 else rng.integers(30, 4000, 2)
        box = (x0, y0, x0 + w, y0 + h)
        kw = dict(block_size=float(rng.uniform(20, 900)), min_num_blocks=int(rng.integers(1, 4)), shrink_factor=float(rng.choice([1, 0.7, 0.5])),
                  round_output=bool(it % 4))
        _same(np.stack(bk.divide_bbox(box, **kw)), np.stack(mo.divide_bbox_oracle(box, **kw)))
        g = rng.integers(0, int(rng.integers(1, 70)), (int(rng.integers(1, 400)), 2)).astype(np.float64)
        _same(bk.z_order(g), mo.z_order_oracle(g))
    _same(bk.z_order(np.array([[0.5, 1.0], [2.0, 0.0], [1.0, 1.0]])), mo.z_order_oracle(np.array([[0.5, 1.0], [2.0, 0.0], [1.0, 1.0]])))   # general loop
    boxes = np.stack(mo.divide_bbox_oracle((3.2, -7.9, 3003.2, 392.1), block_size=75, min_num_blocks=2), -1)
    _same(bk.split_batches(boxes, boxes, None), np.array([0, boxes.shape[0]]))
    _same(bk.split_batches(boxes, boxes, 10 ** 6), np.array([0, boxes.shape[0]]))
    assert bk.split_batches(boxes, boxes, 7).size > 2


def test_auto_spacings_and_division():
    # 3000 x 500 strips -> [75, 750] (float fuzz kept), 400 x 500 -> [75]
    for s0, s1 in [((3000, 500), (3000, 500)), ((400, 4000), (400, 4000)), ((400, 500), (400, 500)), ((90, 70), (80, 75))]:
        _same(bk.auto_spacings(s0, s1), mo.auto_spacings_oracle(s0, s1))
    for shape in [(200, 500), (500, 200), (300, 300), (1500, 250)]:
        for f in (6, 10, 20, (1, 4)):
            assert tuple(bk.balanced_division(shape, f)) == tuple(mo._balanced_division(shape, f))


def test_split_batches_matches_oracle_partition():
    rng = np.random.default_rng(2)
    b = np.stack(mo.divide_bbox_oracle((0, 0, 3000, 400), block_size=75, min_num_blocks=2), -1)
    big = np.stack(mo.divide_bbox_oracle((0, 0, 3000, 400), block_size=750), -1)
    boxes = np.concatenate((big, b), 0)
    for bs in (None, 7, 100, 1000):
        edges = bk.split_batches(boxes, boxes, bs)
        assert edges[0] == 0 and edges[-1] == boxes.shape[0] and np.all(np.diff(edges) > 0)
        assert big.shape[0] in edges                      # size change starts a new batch
        if bs is not None:
            assert np.diff(edges).max() <= max(bs, 1) + 1
