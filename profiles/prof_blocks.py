"""Where the time of the config-4 block pass goes (align512_blocks): host profile + device time per stage."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import feabas_b200.cuda as fc
from feabas_b200.cuda import image as im, matcher as pm

size, h, w, sigma = 8192, 512, 512, 3.5
dev = torch.device('cuda', 0)
a, b = bench.make_section_pair(size, 300, dev, (7, -5))
m0 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=0)
m1 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=1)
boxes = np.array([(x * w, y * h, x * w + w, y * h + h) for y in range(size // h) for x in range(size // w)], dtype=np.float64)
l0, l1 = fc.ArrayLoader(a), fc.ArrayLoader(b)
kw = dict(sigma=sigma, batch_size=len(boxes), pad=True, subpixel=True)


def step():
    return fc.bboxes_mesh_renderer_matcher(m0, m1, l0, l1, boxes, boxes, **kw)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
torch.cuda.synchronize()
print('step ms', (time.perf_counter() - t0) * 100)


def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n, out


rows, shape = pm._block_rows(m0, l0, boxes)
x_lo, y_lo, x_hi, y_hi = m0.covered_rect()
cover = (x_lo, y_lo, x_hi, y_hi)
print('footprint host ms', timed(lambda: im.footprint_uncovered_area(rows, 512, 512, cover))[1])
d, hst, (stack, mask) = timed(lambda: im.crop_blocks_masked(l0.tensor, rows, shape, fillval=0, cover=cover))
print('crop (masked) device ms %.3f wall %.3f mask %s' % (d, hst, None if mask is None else tuple(mask.shape)))
d, hst, (stack2, _) = timed(lambda: im.crop_blocks_masked(l0.tensor, rows, shape, fillval=0, cover=None))
print('crop (no cover) device ms %.3f wall %.3f' % (d, hst))
d, hst, _ = timed(lambda: im.masked_dog_device(stack, sigma, mask))
print('dog masked device ms %.3f wall %.3f' % (d, hst))
d, hst, f0 = timed(lambda: im.masked_dog_device(stack, sigma, None))
print('dog plain device ms %.3f wall %.3f' % (d, hst))
d, hst, _ = timed(lambda: fc.xcorr_fft_device(f0, f0, subpixel=True, pad=True))
print('xcorr device ms %.3f wall %.3f' % (d, hst))
d, hst, _ = timed(lambda: pm._render_stack(m0, l0, boxes, sigma))
print('render_stack device ms %.3f wall %.3f' % (d, hst))
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(30)
