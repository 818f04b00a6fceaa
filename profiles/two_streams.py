"""Experiment: the same batch as bench.py's default workload, split over S CUDA streams (separate workspaces per
stream inside the library), to see whether kernels of different kinds (HBM-bound rows, FP32-bound columns)
co-scheduled on the SMs beat the serial pipeline.  usage: python profiles/two_streams.py [streams] [offset]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
import feabas_b200.cuda as fc

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps, batch = int(os.environ.get("FB_STEPS", 100)), 256
dev = torch.device('cuda', 0)
fc._lib.set_option('ws_bytes', 3 << 30)
a, b, shifts = bench.make_pairs(batch, 512, 512, 100, dev)
parts = [(a[i::S].contiguous(), b[i::S].contiguous()) for i in range(S)]
outs = [torch.empty((5, p[0].shape[0]), dtype=torch.float64, device=dev) for p in parts]
streams = [torch.cuda.Stream() for _ in range(S)]

def step():
    for s, (pa, pb), o in zip(streams, parts, outs):
        with torch.cuda.stream(s):
            fc.xcorr_fft_device(pa, pb, subpixel=True, pad=True, out=o)

for _ in range(5):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in streams:
    s.wait_event(e0)
for _ in range(steps):
    step()
for s in streams:
    torch.cuda.current_stream().wait_stream(s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f'streams {S}: {batch * steps / ms * 1e3:.0f} matches/s, {ms / steps:.4f} ms per step of {batch} pairs')
