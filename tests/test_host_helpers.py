"""CPU tier: small host-side helpers of the matcher layers and of bench.py (no GPU, no library calls)."""
import numpy as np

import bench
from feabas_b200.cuda import image as _img


def _rows(x0, y0, a=(1.0, 0.0, 0.0, 1.0), t=(0.0, 0.0), step=1.0):
    # x0, y0, step_x, step_y, A00, A10, t0, A01, A11, t1
    return np.array([[x0, y0, step, step, a[0], a[1], t[0], a[2], a[3], t[1]]], dtype=np.float64)


def test_blocks_inside_is_exact_at_the_cover_edges():
    cover = (0.0, 0.0, 1024.0, 768.0)
    assert _img._blocks_inside(_rows(0, 0), 64, 64, cover)                     # touches the lower edges: x = 0 is inside
    assert _img._blocks_inside(_rows(960, 704), 64, 64, cover)                 # last pixel 1023 / 767 < upper edges
    assert not _img._blocks_inside(_rows(961, 704), 64, 64, cover)             # last pixel 1024: outside
    assert not _img._blocks_inside(_rows(-1, 0), 64, 64, cover)
    rot = (np.cos(0.1), -np.sin(0.1), np.sin(0.1), np.cos(0.1))
    assert _img._blocks_inside(_rows(300, 300, a=rot), 64, 64, cover)
    assert not _img._blocks_inside(_rows(0, 0, a=rot), 64, 64, cover)          # the rotation pushes a corner below 0
    both = np.concatenate((_rows(0, 0), _rows(2000, 0)))
    assert not _img._blocks_inside(both, 64, 64, cover)


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


def test_montage_overlaps_of_the_stitch_workload():
    wl = dict(bench.WORKLOADS['stitch2x3'], tile=(300, 400), margin=10)
    strips = bench.make_overlap_strips(wl, seed=1)
    assert len(strips) == 11                                                    # 4 horizontal + 3 vertical + 4 diagonal
    shapes = sorted({(a.shape, b.shape) for a, b in strips})
    assert all(a == b for a, b in shapes)
    assert ((300, 50), (300, 50)) in shapes and ((40, 400), (40, 400)) in shapes and ((40, 50), (40, 50)) in shapes


def test_numa_binding_never_raises():
    assert bench.bind_to_gpu_numa(0) >= 0
