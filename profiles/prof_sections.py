"""cProfile of the section_matcher job workload (run on the GPU box): where the host time of the coarse-to-fine loop goes."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import bench, torch
import feabas_b200.cuda as fc
wl = bench.WORKLOADS['thumb_sections']
jobs = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in bench.make_jobs(dict(wl, pairs=3), 1)]
def run():
    for a, b in jobs:
        hh, ww = a.shape
        fc.section_matcher(fc.AffineMesh((0, 0, ww, hh), uid=0), fc.AffineMesh((0, 0, ww, hh), uid=1),
                           fc.ArrayLoader(a, device=0), fc.ArrayLoader(b, device=0), **bench.SECTION_KW)
for _ in range(3): run()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(28); print(s.getvalue()[:7000])
