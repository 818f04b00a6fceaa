"""Host-side profile of the lockstep stitching path (cProfile): python profiles/prof_stitch.py [n_overlaps]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import feabas_b200.cuda as fc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
wl = dict(bench.WORKLOADS['stitch20x20'])
jobs = bench.make_jobs(wl, 1, 0, n)
for _ in range(2):
    fc.stitching_matcher_many(jobs, **bench.STITCH_KW)
torch.cuda.synchronize()
t0 = time.perf_counter()
fc.stitching_matcher_many(jobs, **bench.STITCH_KW)
torch.cuda.synchronize()
print('overlaps/s', n / (time.perf_counter() - t0))
pr = cProfile.Profile()
pr.enable()
fc.stitching_matcher_many(jobs, **bench.STITCH_KW)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
