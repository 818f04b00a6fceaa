import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import bench, torch
import feabas_b200.cuda as fc
wl = bench.WORKLOADS['stitch2x3']
strips = bench.make_overlap_strips(wl, 1)
def run():
    for a, b in strips:
        fc.stitching_matcher(a, b, device=0, **bench.STITCH_KW)
for _ in range(3): run()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): run()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
