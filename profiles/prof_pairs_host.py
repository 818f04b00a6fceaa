"""Host-side view of the config-4 job list: how long each enqueue call returns in while the GPU is busy (a call that
takes as long as the device work in front of it is synchronising somewhere)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import feabas_b200.cuda as fc
from feabas_b200.cuda import matcher as pm, image as im, xcorr as xc

n_pairs = 4
size, h, w, sigma = 8192, 512, 512, 3.5
dev = torch.device('cuda', 0)
secs = [bench.make_section_pair(size, 300 + k, dev, (7, -5)) for k in range(n_pairs)]
m0 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=0)
m1 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=1)
boxes = np.array([(x * w, y * h, x * w + w, y * h + h) for y in range(size // h) for x in range(size // w)], dtype=np.float64)
loaders = [(fc.ArrayLoader(a), fc.ArrayLoader(b)) for a, b in secs]
kw = dict(sigma=sigma, batch_size=len(boxes), pad=True, subpixel=True)
for _ in range(2):
    fc.bboxes_mesh_renderer_matcher_many(((m0, m1, l0, l1, boxes, boxes) for l0, l1 in loaders), **kw)
torch.cuda.synchronize()

# wrap the stages with host timers
log = []
def wrap(mod, name):
    fn = getattr(mod, name)
    def timed(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        log.append((name, (time.perf_counter() - t0) * 1e3))
        return out
    setattr(mod, name, timed)
wrap(im, 'crop_blocks_masked'); wrap(im, 'masked_dog_device'); wrap(pm, 'xcorr_fft_device'); wrap(im, 'footprint_uncovered_area')
t_all = time.perf_counter()
q = []
for l0, l1 in loaders:
    t0 = time.perf_counter()
    q.append(pm._bboxes_enqueue(m0, m1, l0, l1, boxes, boxes, **kw))
    log.append(('ENQUEUE job', (time.perf_counter() - t0) * 1e3))
for e in q:
    t0 = time.perf_counter()
    pm._bboxes_collect(e)
    log.append(('COLLECT job', (time.perf_counter() - t0) * 1e3))
print('total ms', (time.perf_counter() - t_all) * 1e3)
for name, ms in log:
    print(f'{name:28s} {ms:8.3f} ms')
