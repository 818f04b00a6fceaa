"""Run the golden cases one by one (for compute-sanitizer --tool initcheck: which case reads uninitialised memory)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import feabas_b200.cuda as fc
from conftest import load_golden
from conftest import case_kwargs
from feabas_b200 import synth

small = load_golden('xcorr_small.npz')
for force in (None, 'staged'):
    for name, rec in small.items():
        kw = case_kwargs(rec)
        try:
            fc.xcorr_fft(rec['img0'], rec['img1'], force=force, **kw)
        except Exception as e:
            print('CASE', force, name, 'skipped', str(e)[:60], flush=True)
            continue
        torch.cuda.synchronize()
        plan = None
        try:
            a, b = np.asarray(rec['img0']), np.asarray(rec['img1'])
            print('CASE', force, name, a.shape, b.shape, a.dtype, kw, flush=True)
        except Exception:
            pass
seeded = load_golden('xcorr_seeded.npz')
for name, rec in seeded.items():
    s0, s1, _ = synth.block_pairs(int(rec['n']), rec['size'].tolist(), int(rec['seed']), max_shift=int(rec['max_shift']))
    kw = case_kwargs(rec)
    fc.xcorr_fft(s0, s1, **kw)
    torch.cuda.synchronize()
    print('CASE seeded', name, s0.shape, kw, flush=True)
