"""CPU tier: small host-side helpers of the matcher layers and of bench.py (no GPU, no library calls)."""
import numpy as np

import bench
from feabas_b200.cuda import image as _img


def _rows(x0, y0, a=(1.0, 0.0, 0.0, 1.0), t=(0.0, 0.0), step=1.0):
    # x0, y0, step_x, step_y, A00, A10, t0, A01, A11, t1
    return np.array([[x0, y0, step, step, a[0], a[1], t[0], a[2], a[3], t[1]]], dtype=np.float64)


def test_footprint_uncovered_area_follows_the_reference_rule():
    """feabas/renderer.py:436-449: footprint = bbox - 0.5 through the block's affine map; < 1 px^2 outside the covered
    region -> the block is rendered whole.  Checked against the oracle's convex-polygon restatement."""
    from oracle import convex
    cover = (0.0, 0.0, 1023.0, 767.0)                     # mesh on bounds (0, 0, 1024, 768): vertices - 0.5, shrunk by 0.5
    area = lambda rows: _img.footprint_uncovered_area(rows, 64, 64, cover)
    assert area(_rows(1, 1))[0] == 0                      # footprint [0.5, 64.5]^2: inside
    assert area(_rows(0, 0))[0] == 2 * 0.5 * 64 - 0.25    # half a pixel sticks out along two edges
    np.testing.assert_allclose(area(_rows(-3, 100))[0], 3.5 * 64)
    rot = (np.cos(0.1), -np.sin(0.1), np.sin(0.1), np.cos(0.1))
    rows = np.concatenate((_rows(300, 300, a=rot), _rows(0, 0, a=rot), _rows(2000, 0), _rows(1000.3, 740.6, a=rot, t=(3.0, -2.0))))
    got = area(rows)
    for r, g in zip(rows, got):
        foot = convex.affine_transform(convex.box(r[0] - 0.5, r[1] - 0.5, r[0] + 63.5, r[1] + 63.5), (r[4], r[5], r[7], r[8], r[6], r[9]))
        want = foot.area - convex.box(*cover).intersection(foot).area
        np.testing.assert_allclose(g, want, atol=1e-9)
    assert got[0] == 0 and got[2] == 64 * 64


def test_batch_origin_matches_the_field_minimum():
    b = np.concatenate((_rows(10.5, 20.25), _rows(300, 7, a=(0.9, 0.1, -0.1, 0.9), t=(5.0, -3.0))))
    ox, oy = _img.batch_origin(b, 32, 48)
    xs, ys = [], []
    for r in b:
        for row in (0, 31):
            for col in (0, 47):
                xx, yy = r[0] + col * r[2], r[1] + row * r[3]
                xs.append(xx * r[4] + yy * r[5] + r[6])
                ys.append(xx * r[7] + yy * r[8] + r[9])
    assert ox == np.floor(min(xs)) - 4 and oy == np.floor(min(ys)) - 4


def test_batch_origins_of_many_batches_equal_one_call_per_batch():
    """``crop_blocks_multi`` takes the source-crop origins of all its parts in one vectorised pass."""
    rng = np.random.default_rng(1)
    for _ in range(100):
        counts = rng.integers(1, 40, int(rng.integers(1, 6)))
        rows = rng.uniform(-50, 50, (counts.sum(), 10))
        rows[:, 2:4] = rng.uniform(0.5, 1.5, (counts.sum(), 2))
        rows[:, [4, 8]] = 1 + rng.normal(0, 0.01, (counts.sum(), 2))
        starts = np.cumsum(counts) - counts
        ox, oy = _img.batch_origins(rows, 64, 48, starts)
        for j, (lo, c) in enumerate(zip(starts, counts)):
            assert _img.batch_origin(rows[lo:lo + c], 64, 48) == (ox[j], oy[j])


def test_montage_overlaps_of_the_stitch_workload():
    wl = dict(bench.WORKLOADS['stitch2x3'], tile=(300, 400), margin=10)
    strips = bench.make_overlap_strips(wl, seed=1)
    assert len(strips) == 11                                                    # 4 horizontal + 3 vertical + 4 diagonal
    shapes = sorted({(a.shape, b.shape) for a, b in strips})
    assert all(a == b for a, b in shapes)
    assert ((300, 50), (300, 50)) in shapes and ((40, 400), (40, 400)) in shapes and ((40, 50), (40, 50)) in shapes


def test_numa_binding_never_raises():
    assert bench.bind_to_gpu_numa(0) >= 0
