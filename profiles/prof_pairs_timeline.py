"""Kernel timeline of the config-4 job list (align512_pairs, 4 section pairs) from CUPTI (torch.profiler): per-kernel
device time in situ, GPU busy fraction, the gaps.  python profiles/prof_pairs_timeline.py [pairs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import bench
import feabas_b200.cuda as fc

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 2
size, h, w, sigma = 8192, 512, 512, 3.5
dev = torch.device('cuda', 0)
secs = [bench.make_section_pair(size, 300 + k, dev, (7, -5)) for k in range(n_pairs)]
m0 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=0)
m1 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=1)
boxes = np.array([(x * w, y * h, x * w + w, y * h + h) for y in range(size // h) for x in range(size // w)], dtype=np.float64)
loaders = [(fc.ArrayLoader(a), fc.ArrayLoader(b)) for a, b in secs]
kw = dict(sigma=sigma, batch_size=len(boxes), pad=True, subpixel=True)


def step():
    return fc.bboxes_mesh_renderer_matcher_many(((m0, m1, l0, l1, boxes, boxes) for l0, l1 in loaders), streams=n_streams, **kw)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
# union of busy intervals
busy, cur_s, cur_e = 0.0, None, None
for e in ev:
    s, t = e.time_range.start, e.time_range.end
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, t
    else:
        cur_e = max(cur_e, t)
busy += cur_e - cur_s
print(f'{n_pairs} pairs: span {1e-3 * (t1 - t0):.3f} ms = {1e-3 * (t1 - t0) / n_pairs:.3f} ms per pair, GPU busy {busy / (t1 - t0):.3f}')
agg = {}
for e in ev:
    k = e.name.replace('void ', '').replace('(anonymous namespace)::', '').split('(')[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.end - e.time_range.start
print(f'{"kernel":72s} {"calls":>6s} {"ms/pair":>9s} {"us/call":>9s}')
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:72s} {c:6d} {1e-3 * t / n_pairs:9.3f} {t / c:9.1f}')
# largest idle gaps
gaps = []
cur_e = None
for e in ev:
    if cur_e is not None and e.time_range.start > cur_e:
        gaps.append((e.time_range.start - cur_e, e.name[:60]))
    cur_e = e.time_range.end if cur_e is None else max(cur_e, e.time_range.end)
gaps.sort(reverse=True)
print('largest gaps (us, next kernel):', [(round(g, 1), n) for g, n in gaps[:8]], 'total gap ms/pair', 1e-3 * sum(g for g, _ in gaps) / n_pairs)
