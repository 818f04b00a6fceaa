#!/usr/bin/env python
"""Benchmark of the xcorr hot path (BASELINE.json metric: xcorr block-matches/sec + HBM GB/s vs
roofline, next to the host-CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step = one pass of ``xcorr_fft`` over one batch of synthetic EM-like block pairs.  Default
workload ``xcorr512``: 512x512 float32 blocks, pad=True (FFT 1024x1024), FFT_CONF_MIRROR,
subpixel=True -- the block shape of BASELINE.json configs[3]/[4] and of the north-star target.
N > 1 (torchrun, one process per GPU): every rank processes its own batch (independent pairs,
no data-path collective; weak scaling); time = max over ranks.

Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every key).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (block h, w, pad, pairs per step per GPU)
    'xcorr512': dict(h=512, w=512, pad=True, batch=256),
    'xcorr512_nopad': dict(h=512, w=512, pad=False, batch=512),
    'xcorr128': dict(h=128, w=128, pad=True, batch=4096),
    'xcorr256': dict(h=256, w=256, pad=True, batch=1024),
    'xcorr1024': dict(h=1024, w=1024, pad=True, batch=64),
    'xcorr2048': dict(h=2048, w=2048, pad=True, batch=16),
    'stitch_fine': dict(h=74, w=67, pad=True, batch=16384),       # config 1/2 finest level, FFT 150x135
    'thumb150': dict(h=150, w=150, pad=True, batch=4096),         # config 3, FFT 300x300
    'xcorr300': dict(h=300, w=300, pad=True, batch=1024),         # FFT 600x600 (60 points per lane)
    'align280': dict(h=280, w=280, pad=True, batch=1024),         # default fine alignment (spacing 400, shrink 0.7), FFT 576x576
    # config 4 through bboxes_mesh_renderer_matcher (matcher.py:781-861): a uint8 section pair, a dense grid of
    # 512x512 blocks gathered + band-passed (sigma 3.5) + matched on the device; e2e uploads the two sections
    'align512_blocks': dict(kind='blocks', h=512, w=512, pad=True, section=8192, sigma=3.5, batch=256),
    # config 1 (BASELINE.json configs[0]): stitching_matcher on every overlap of a 2x3 montage of 3000x4000 uint8 tiles,
    # 10 % overlap, shipped stitching YAML kwargs; the unit counted is the block match (one xcorr of one block pair)
    'stitch2x3': dict(kind='stitch', rows=2, cols=3, tile=(3000, 4000), overlap=0.1, margin=100, batch=0, h=74, w=67, pad=True),
    # config 3: section_matcher (coarse-to-fine block matching, spacings [150, 50]) on pairs of band-passed 2048^2 thumbnails,
    # kwargs of default_thumbnail_configs.yaml:43-53 as thumbnail.py:509 passes them (sigma applied upstream -> 0 here)
    'thumb_sections': dict(kind='sections', pairs=8, size=2048, batch=0, h=50, w=50, pad=True),
}


def algorithmic_bytes(h, w, ny, nx, mirror=True, fused=False):
    """B_min when the pair is resident on chip, else the 3-stage B_pass."""
    kp = nx // 2 + 1
    n_out = 2 if mirror else 1
    b_min = 2 * h * w * 4 + 20
    b_pass = 2 * h * w * 4 + 2 * (2 * h * kp * 8) + 2 * (n_out * ny * kp * 8)
    return (b_min if fused else b_pass), b_min, b_pass


def column_kernel_bytes(h, ny, nx, mirror=True):
    """Algorithmic bytes of the dominant kernel (column stage) per pair: read both row spectra,
    write the P (and Q) half surfaces."""
    kp = nx // 2 + 1
    return 2 * h * kp * 8 + (2 if mirror else 1) * ny * kp * 8


def make_pairs(n, h, w, seed, device, max_shift=32):
    """Seeded synthetic EM-like block pairs with known integer displacement (torch, any device):
    band-limited texture cut from one canvas per pair at two offsets + independent noise."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    m = max_shift + 8
    shifts = torch.randint(-max_shift, max_shift + 1, (n, 2), generator=g, device=device)

    def blur(x, sigma):
        r = int(4 * sigma + 0.5)
        t = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
        k = torch.exp(-0.5 * (t / sigma) ** 2)
        k = k / k.sum()
        x = F.conv2d(F.pad(x, (r, r, 0, 0), mode='replicate'), k.view(1, 1, 1, -1))
        return F.conv2d(F.pad(x, (0, 0, r, r), mode='replicate'), k.view(1, 1, -1, 1))

    a = torch.empty((n, h, w), dtype=torch.float32, device=device)
    b = torch.empty((n, h, w), dtype=torch.float32, device=device)
    step = max(1, min(n, (1 << 26) // ((h + 2 * m) * (w + 2 * m))))
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        c = torch.randn((hi - lo, 1, h + 2 * m, w + 2 * m), generator=g, device=device)
        c = blur(c, 2.0)
        c = c / c.std()
        c0 = c + 0.25 * torch.randn(c.shape, generator=g, device=device)
        c1 = c + 0.25 * torch.randn(c.shape, generator=g, device=device)
        c0 = blur(c0, 2.5) - blur(blur(c0, 2.5), 2.5)           # DoG-like band-pass, sigma 2.5
        c1 = blur(c1, 2.5) - blur(blur(c1, 2.5), 2.5)
        for i in range(lo, hi):
            dx, dy = int(shifts[i, 0]), int(shifts[i, 1])
            a[i] = c0[i - lo, 0, m:m + h, m:m + w]
            b[i] = c1[i - lo, 0, m - dy:m - dy + h, m - dx:m - dx + w]
    return a, b, shifts


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's xcorr_fft on the host cores
# ----------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(h, w, pad, per_worker, seed):
    os.environ['OMP_NUM_THREADS'] = '1'                       # reference: config.py:301-310
    import numpy as np
    rng = np.random.default_rng(seed + os.getpid())
    _W['a'] = rng.standard_normal((per_worker, h, w)).astype(np.float32)
    _W['b'] = np.roll(_W['a'], (3, -5), axis=(1, 2)) + 0.1 * rng.standard_normal((per_worker, h, w)).astype(np.float32)
    _W['pad'] = pad
    from oracle import xcorr_oracle as xo
    _W['f'] = xo.xcorr_oracle
    _W['f'](_W['a'][:1], _W['b'][:1], subpixel=True, pad=pad)  # warm


def _cpu_task(_):
    t = time.perf_counter()
    _W['f'](_W['a'], _W['b'], conf_mode=2, subpixel=True, pad=_W['pad'])
    return time.perf_counter() - t


class CpuArm:
    """Process pool, one single-threaded worker per physical core (the reference's own parallel
    model: feabas/concurrent.py:59-96), each running the oracle port on its own pairs."""

    def __init__(self, wl, per_worker):
        import psutil
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        self.cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        self.per_worker = per_worker
        self.pool = ProcessPoolExecutor(self.cores, mp_context=mp.get_context('spawn'),
                                        initializer=_cpu_init, initargs=(wl['h'], wl['w'], wl['pad'], per_worker, 1234))
        list(self.pool.map(_cpu_task, range(self.cores)))      # spawn + warm every worker

    def step(self):
        """One timed pass: every worker processes its pairs once.  Returns (pairs, seconds)."""
        t = time.perf_counter()
        list(self.pool.map(_cpu_task, range(self.cores)))
        return self.cores * self.per_worker, time.perf_counter() - t

    def close(self):
        self.pool.shutdown()


def cpu_pairs_per_worker(wl, target_cpu_seconds, cores):
    # ~48 ms per 512^2 padded pair per core; scale by FFT area
    est = 48e-3 * (wl['h'] * wl['w'] * (4 if wl['pad'] else 1)) / (512 * 512 * 4)
    return max(1, int(round(target_cpu_seconds / cores / max(est, 1e-5))))



# ----------------------------------------------------------------------------------------------
# matcher-level workload: block grid of a section pair through bboxes_mesh_renderer_matcher
# ----------------------------------------------------------------------------------------------
def make_section_pair(size, seed, device, shift=(7, -5)):
    """Two uint8 'sections' (size x size) cut from one band-limited canvas at a known integer offset
    + independent noise.  A feature at p in img0 sits at p + shift in img1."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    m = 32
    c = torch.randn((1, 1, size + 2 * m, size + 2 * m), generator=g, device=device)
    t = torch.arange(-12, 13, device=device, dtype=torch.float32)
    k = torch.exp(-0.5 * (t / 3.0) ** 2); k = k / k.sum()
    c = F.conv2d(F.pad(c, (12, 12, 0, 0), mode='replicate'), k.view(1, 1, 1, -1))
    c = F.conv2d(F.pad(c, (0, 0, 12, 12), mode='replicate'), k.view(1, 1, -1, 1))[0, 0]
    c = c / c.std()
    dx, dy = shift
    a = c[m:m + size, m:m + size]
    b = c[m - dy:m - dy + size, m - dx:m - dx + size]
    def u8(x):
        x = x + 0.2 * torch.randn(x.shape, generator=g, device=device)
        return (x * 40 + 128).clamp_(0, 255).to(torch.uint8).contiguous()
    return u8(a), u8(b)


def _cpu_blocks_init(h, w, pad, per_worker, sigma, seed):
    os.environ['OMP_NUM_THREADS'] = '1'
    import numpy as np
    rng = np.random.default_rng(seed + os.getpid())
    side = int(np.ceil(np.sqrt(per_worker)))
    img = rng.integers(0, 256, (side * h + 8, side * w + 8), dtype=np.uint8)
    _W['img0'] = img
    _W['img1'] = np.roll(img, (3, -5), axis=(0, 1))
    _W['boxes'] = [(x * w, y * h, x * w + w, y * h + h) for y in range(side) for x in range(side)][:per_worker]
    _W['pad'], _W['sigma'] = pad, sigma
    from oracle import xcorr_oracle as xo
    from oracle import matcher_oracle as mo
    _W['f'], _W['dog'] = xo.xcorr_oracle, mo.masked_dog_oracle


def _cpu_blocks_task(_):
    """crop + masked DoG + xcorr_fft of this worker's blocks (renderer.py:601-648 -> matcher.py:846)."""
    import numpy as np
    t = time.perf_counter()
    st0 = np.stack([_W['img0'][y0:y1, x0:x1] for x0, y0, x1, y1 in _W['boxes']])
    st1 = np.stack([_W['img1'][y0:y1, x0:x1] for x0, y0, x1, y1 in _W['boxes']])
    f0 = _W['dog'](st0, _W['sigma']).astype(np.float32)
    f1 = _W['dog'](st1, _W['sigma']).astype(np.float32)
    _W['f'](f0, f1, conf_mode=2, subpixel=True, pad=_W['pad'])
    return time.perf_counter() - t


def cpu_blocks_arm(wl, per_worker):
    import psutil
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    pool = ProcessPoolExecutor(cores, mp_context=mp.get_context('spawn'), initializer=_cpu_blocks_init,
                               initargs=(wl['h'], wl['w'], wl['pad'], per_worker, wl['sigma'], 4321))
    list(pool.map(_cpu_blocks_task, range(cores)))            # spawn + warm
    return pool, cores


def bench_blocks(args, wl, rank, world, local, warmup):
    import numpy as np
    h, w, pad, size, sigma = wl['h'], wl['w'], wl['pad'], wl['section'], wl['sigma']
    nby, nbx = size // h, size // w
    batch = nby * nbx
    from oracle import xcorr_oracle as xo
    ny, nx = xo.fft_shape((h, w), (h, w), pad)
    config = {'workload': f'{args.workload}: bboxes_mesh_renderer_matcher on a {size}x{size} uint8 section pair, {nby}x{nbx} grid of '
                          f'{h}x{w} blocks, sigma={sigma} (masked DoG), pad={pad} (FFT {ny}x{nx}), FFT_CONF_MIRROR, subpixel=True',
              'pairs_per_step_per_gpu': batch, 'fft': [ny, nx], 'l2': 'inputs larger than L2 (no flush needed)'}
    kw = dict(sigma=sigma, batch_size=batch, pad=pad, subpixel=True)

    if args.impl == 'reference':
        if rank != 0:
            return
        per_step_wall = min(3.0, max(0.3, 90.0 / (args.steps + warmup)))
        pool, cores = None, 0
        import psutil
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        per_worker = max(1, int(round(per_step_wall / 0.085)))   # ~85 ms per 512^2 block pair per core (crop + 2 DoG + xcorr)
        pool, cores = cpu_blocks_arm(wl, per_worker)
        for _ in range(warmup):
            list(pool.map(_cpu_blocks_task, range(cores)))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(_cpu_blocks_task, range(cores)))
        secs = time.perf_counter() - t0
        pool.shutdown()
        val = cores * per_worker * args.steps / secs
        sample = f'{per_worker} blocks per worker x {cores} single-thread workers per step (oracle port: crop + masked DoG + xcorr_fft)'
        print(json.dumps({
            'impl': 'reference', 'metric': 'xcorr_block_matches_per_sec', 'value': val, 'unit': 'matches/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': val, 'unit': 'matches/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    config['numa_bound_cpus'] = bind_to_gpu_numa(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    L = fc._lib
    L.set_option('ws_bytes', int(args.ws_gib * (1 << 30)))
    shift = (7, -5)
    a, b = make_section_pair(size, 300 + rank, dev, shift)
    m0, m1 = fc.AffineMesh((0, 0, size, size), uid=0), fc.AffineMesh((0, 0, size, size), uid=1)
    boxes = np.array([(x * w, y * h, x * w + w, y * h + h) for y in range(nby) for x in range(nbx)], dtype=np.float64)
    l0, l1 = fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        return fc.bboxes_mesh_renderer_matcher(m0, m1, l0, l1, boxes, boxes, **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        xy0, xy1, conf = step()
    d = xy1 - xy0                                               # every block must recover the section offset
    n_ok = int(np.sum((np.round(d[:, 0]) == shift[0]) & (np.round(d[:, 1]) == shift[1])))
    assert n_ok >= 0.99 * batch, f'only {n_ok}/{batch} blocks recovered the offset {shift}: {d[:4]}'
    config['ground_truth_recovered'] = n_ok / batch
    L.profile_read(local, stream, reset=True) if L.launch_count() else None
    L.set_option('profile', 1)
    mon, mon_path = clocks_monitor_start(local) if rank == 0 else (None, None)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = clocks_monitor_stop(mon, mon_path) if rank == 0 else None
    L.set_option('profile', 0)
    prof = L.profile_read(local, stream, reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * batch * args.steps / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        ah, bh = a.cpu().pin_memory(), b.cpu().pin_memory()
        e_steps = max(3, min(args.steps, 10))

        def e_step():
            # the call a user makes: host images in, matches out (upload of both sections inside)
            return fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(ah, device=local), fc.ArrayLoader(bh, device=local),
                                                   boxes, boxes, **kw)
        for _ in range(2):
            r = e_step()
        assert np.array_equal(r[0], xy0) and np.array_equal(r[2], conf), 'host path and device path disagree'
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            r = e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': world * batch * e_steps / dt, 'unit': 'matches/s',
               'h2d_bytes_per_step': int(ah.numel() + bh.numel()), 'd2h_bytes_per_step': int(5 * 8 * batch), 'steps': e_steps,
               'api': 'feabas_b200.cuda.bboxes_mesh_renderer_matcher(ArrayLoader(pinned uint8 host sections)) -> fb_crop_blocks + fb_masked_dog + fb_xcorr_batch_device'}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = peaks()
    b_alg, b_min, b_pass = algorithmic_bytes(h, w, ny, nx, True, False)
    kern = {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)} for k, v in prof.items() if v[1]}
    dom = max(kern, key=lambda k: kern[k]['ms_total'])
    pairs_per_launch = batch * args.steps / kern[dom]['launches']
    bytes_per_pair = column_kernel_bytes(h, ny, nx, True) if dom == 'columns' else b_min
    achieved = bytes_per_pair * pairs_per_launch / (kern[dom]['ms_per_launch'] * 1e-3) / 1e9
    xcorr_ms = sum(v['ms_total'] for v in kern.values())
    # image stage (fb_crop_blocks + fb_masked_dog, both sections): gather u8 -> f32 stack, two blur passes per stack
    img_bytes = 2 * batch * h * w * (1 + 4 + 3 * 8)              # read u8, write f32; DoG: 2 x (read + write) + final read/write
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': None, 'peak_kind': peak_kind, 'algorithmic_bytes_per_pair': bytes_per_pair,
                'pairs_per_launch': pairs_per_launch, 'ms_per_launch': kern[dom]['ms_per_launch'],
                'kernel_share_of_step': kern[dom]['ms_total'] / ms if world == 1 else None,
                'pipeline': {'bytes_per_pair': b_alg, 'b_min': b_min, 'b_pass': b_pass,
                             'achieved': value / world * b_alg / 1e9, 'frac': value / world * b_alg / 1e9 / peak},
                'kernel_timing': 'timed region; the library runs two half-batches on two streams, so the per-kernel event times overlap',
                'image_stage': {'ms_per_step': None,
                                'algorithmic_bytes_per_step': img_bytes,
                                'note': 'crop (u8 gather -> f32) + masked DoG of both stacks + host control flow = step - xcorr kernels'},
                'kernels': kern}
    cpu = None
    if not args.no_cpu_baseline and world == 1:          # N = 1 only (ranks of a multi-GPU run are bound to their GPU's CPUs)
        per_worker = 16
        pool, cores = cpu_blocks_arm(wl, per_worker)
        t0 = time.perf_counter()
        list(pool.map(_cpu_blocks_task, range(cores)))
        s_ = time.perf_counter() - t0
        pool.shutdown()
        cpu = {'value': cores * per_worker / s_, 'unit': 'matches/s', 'cores': cores, 'kind': 'port',
               'sample': f'{cores * per_worker} blocks ({per_worker} per single-thread worker, {cores} workers): crop + masked DoG + '
                         f'xcorr_fft (oracle port), {s_:.1f} s wall'}
    line = {'metric': 'xcorr_block_matches_per_sec', 'value': value, 'unit': 'matches/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



# ----------------------------------------------------------------------------------------------
# stitching workload: stitching_matcher over every overlap of a small montage (config 1)
# ----------------------------------------------------------------------------------------------
STITCH_KW = dict(spacings=None, conf_thresh=0.33, residue_mode='huber', residue_len=2, pad=True, sigma=2.5,
                 coarse_downsample=0.5, fine_downsample=1.0, compute_photometric=False)   # default_stitching_configs.yaml:11-20


def make_overlap_strips(wl, seed):
    """Overlap strips (+ margin) of every pair of touching tiles, cut the way stitcher.py:556-568 does."""
    import numpy as np
    from feabas_b200 import synth
    rows, cols, (th, tw), margin = wl['rows'], wl['cols'], wl['tile'], wl['margin']
    tiles, nominal, _ = synth.tile_grid(rows, cols, tile_hw=(th, tw), overlap=wl['overlap'], jitter=10, seed=seed)
    strips = []
    for i in range(len(tiles)):
        for j in range(i + 1, len(tiles)):
            (x0, y0), (x1, y1) = nominal[i], nominal[j]
            bx0, by0, bx1, by1 = max(x0, x1), max(y0, y1), min(x0, x1) + tw, min(y0, y1) + th
            if bx1 - bx0 < 25 or by1 - by0 < 25:
                continue
            bx0, by0, bx1, by1 = bx0 - margin, by0 - margin, bx1 + margin, by1 + margin
            def cut(t, ox, oy):
                xa, ya, xb, yb = max(bx0 - ox, 0), max(by0 - oy, 0), min(bx1 - ox, tw), min(by1 - oy, th)
                return np.ascontiguousarray(t[ya:yb, xa:xb])
            strips.append((cut(tiles[i], x0, y0), cut(tiles[j], x1, y1)))
    return strips


SECTION_KW = dict(sigma=0.0, spacings=[150, 50], conf_thresh=0.35, pad=True, distributor='cartesian_bbox',
                  residue_mode='huber', residue_len=3, batch_size=300)                 # default_thumbnail_configs.yaml:43-53, thumbnail.py:509


def make_section_thumbs(wl, seed):
    """Pairs of float32 band-passed thumbnails (size x size) of neighbouring sections: one canvas, a small offset."""
    import numpy as np
    from feabas_b200 import synth
    size, out = wl['size'], []
    rng = np.random.default_rng(seed)
    for k in range(wl['pairs']):
        canvas = synth.dog_f32(synth.em_canvas(size + 40, size + 40, seed=seed * 100 + k), 3.5)
        dx, dy = rng.integers(-9, 10, 2)
        a = np.ascontiguousarray(canvas[20:20 + size, 20:20 + size])
        b = np.ascontiguousarray(canvas[20 + dy:20 + dy + size, 20 + dx:20 + dx + size])
        out.append((a, b + 0.05 * rng.standard_normal(b.shape).astype(np.float32)))
    return out


def make_jobs(wl, seed):
    return make_overlap_strips(wl, seed) if wl['kind'] == 'stitch' else make_section_thumbs(wl, seed)


def _cpu_stitch_init(seed, wl):
    os.environ['OMP_NUM_THREADS'] = '1'
    _W['strips'] = make_jobs(wl, seed)
    _W['kind'] = wl['kind']
    from oracle import matcher_oracle as mo
    _W['f'] = mo.stitching_oracle
    _W['mo'] = mo


def _cpu_stitch_task(k):
    a, b = _W['strips'][k % len(_W['strips'])]
    trace = []
    if _W['kind'] == 'stitch':
        kw = {k_: v for k_, v in STITCH_KW.items() if k_ not in ('compute_photometric',)}
        out = _W['f'](a, b, trace=trace, **kw)
    else:
        mo = _W['mo']
        sec0 = mo._Section((0, 0, a.shape[1], a.shape[0]), locked=True)
        sec1 = mo._Section((0, 0, b.shape[1], b.shape[0]))
        out = mo.surrogate_loop_oracle(sec0, sec1, a, b, SECTION_KW['spacings'], conf_thresh=SECTION_KW['conf_thresh'],
                                       residue_mode='huber', residue_len=SECTION_KW['residue_len'], pad=True, batch_size=SECTION_KW['batch_size'],
                                       trace=trace)
    blocks = sum(t.get('nblocks', 0) + (1 if 'coarse' in t else 0) for t in trace)
    return 0 if out[0] is None else len(out[0]), blocks


def _gpu_job_worker(workload, seed, steps, local, barrier, queue):
    """One of P worker processes sharing a GPU (FEABAS's own parallel model: spawned workers, one overlap / section pair
    per task, feabas/concurrent.py:59-96): every worker has its own CUDA context and runs the job list `steps` times."""
    try:
        import torch
        torch.cuda.set_device(local)
        import feabas_b200.cuda as fc
        wl = dict(WORKLOADS[workload])
        stitch = wl['kind'] == 'stitch'
        jobs = make_jobs(wl, seed)
        lib = fc._lib.lib()

        def run():
            for a, b in jobs:
                if stitch:
                    fc.stitching_matcher(a, b, device=local, **STITCH_KW)
                else:
                    hh, ww = a.shape
                    fc.section_matcher(fc.AffineMesh((0, 0, ww, hh), uid=0), fc.AffineMesh((0, 0, ww, hh), uid=1),
                                       fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local), **SECTION_KW)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        p0 = lib.fb_pair_count()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            run()
        torch.cuda.synchronize()
        queue.put((lib.fb_pair_count() - p0, len(jobs) * steps, time.perf_counter() - t0))
    except Exception as exc:                                   # pragma: no cover
        queue.put(('error', repr(exc), 0.0))
        try:
            barrier.abort()
        except Exception:
            pass


def multi_process_throughput(workload, workers, steps, local):
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    barrier, queue = ctx.Barrier(workers + 1), ctx.Queue()
    procs = [ctx.Process(target=_gpu_job_worker, args=(workload, 11 + k, steps, local, barrier, queue)) for k in range(workers)]
    for pr in procs:
        pr.start()
    try:
        barrier.wait(timeout=300)
    except Exception:
        for pr in procs:
            pr.terminate()
        return {'workers': workers, 'error': 'workers did not reach the start barrier'}
    t0 = time.perf_counter()
    res = [queue.get(timeout=600) for _ in procs]
    dt = time.perf_counter() - t0
    for pr in procs:
        pr.join(timeout=60)
    if any(r[0] == 'error' for r in res):
        return {'workers': workers, 'error': str([r for r in res if r[0] == 'error'][0][1])}
    return {'workers': workers, 'value': sum(r[0] for r in res) / dt, 'unit': 'matches/s', 'jobs_per_s': sum(r[1] for r in res) / dt,
            'steps_per_worker': steps, 'seconds': dt,
            'note': 'wall clock over P spawned worker processes sharing this GPU (one CUDA context each, host strips / thumbnails in), '
                    'the way FEABAS fans overlaps / section pairs out to workers; supplementary to the single-process numbers'}


def bench_stitch(args, wl, rank, world, local, warmup):
    import numpy as np
    stitch = wl['kind'] == 'stitch'
    if stitch:
        desc = (f"stitching_matcher on every overlap of a {wl['rows']}x{wl['cols']} montage of {wl['tile'][0]}x{wl['tile'][1]} uint8 tiles, "
                f"{int(100 * wl['overlap'])} % overlap, margin {wl['margin']}, shipped YAML kwargs (sigma 2.5, coarse 0.5, pad, conf_thresh 0.33)")
    else:
        desc = (f"section_matcher on {wl['pairs']} pairs of {wl['size']}x{wl['size']} float32 band-passed thumbnails, spacings [150, 50], pad, "
                'conf_thresh 0.35, huber residue 3 (affine surrogate mesh)')
    config = {'workload': f'{args.workload}: {desc}; unit = block match (one xcorr of one block pair)',
              'l2': 'working set per call far below L2: latency bound, not HBM bound'}
    if args.impl == 'reference':
        if rank != 0:
            return
        import psutil
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        pool = ProcessPoolExecutor(cores, mp_context=mp.get_context('spawn'), initializer=_cpu_stitch_init, initargs=(1, wl))
        n_ov = len(make_overlap_strips(dict(wl, tile=(300, 400), margin=10), 1)) if stitch else wl['pairs']     # job count only
        list(pool.map(_cpu_stitch_task, range(cores)))
        steps = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        res = []
        n_tasks = max(n_ov, cores)                             # every worker busy: the job list is cycled
        for _ in range(steps):
            res += list(pool.map(_cpu_stitch_task, range(n_tasks)))
        secs = time.perf_counter() - t0
        pool.shutdown()
        blocks = sum(r[1] for r in res) or sum(r[0] for r in res)
        val = blocks / secs
        sample = (f'{steps} passes over {n_tasks} jobs (the {n_ov} of one step, cycled), one job per task on {cores} single-thread workers '
                  f"(oracle port of {'stitching_matcher' if stitch else 'the section_matcher loop'})")
        n_ov = n_tasks
        print(json.dumps({
            'impl': 'reference', 'metric': 'xcorr_block_matches_per_sec', 'value': val, 'unit': 'matches/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': dict(config, overlaps_per_s=n_ov * steps / secs),
            'cpu_baseline': {'value': val, 'unit': 'matches/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    L = fc._lib
    strips = make_jobs(wl, 1 + rank)
    config['overlaps_per_step_per_gpu' if stitch else 'section_pairs_per_step_per_gpu'] = len(strips)
    dstrips = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in strips]
    hstrips = [(torch.from_numpy(a).pin_memory().numpy(), torch.from_numpy(b).pin_memory().numpy()) for a, b in strips]
    stream = torch.cuda.current_stream().cuda_stream
    lib = L.lib()

    def run(pairs):
        n = 0
        x0 = lib.fb_pair_count() if hasattr(lib, 'fb_pair_count') else 0
        for a, b in pairs:
            if stitch:
                out = fc.stitching_matcher(a, b, device=local, **STITCH_KW)
            else:
                hh, ww = a.shape
                out = fc.section_matcher(fc.AffineMesh((0, 0, ww, hh), uid=0), fc.AffineMesh((0, 0, ww, hh), uid=1),
                                         fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local), **SECTION_KW)
            n += 0 if out[0] is None else len(out[0])
        return n, (lib.fb_pair_count() - x0 if hasattr(lib, 'fb_pair_count') else 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        pts, pairs = run(dstrips)
    assert pts > 0, 'no overlap produced matches'
    config['match_points_per_step'] = pts
    units = pairs or pts
    L.profile_read(local, stream, reset=True) if L.launch_count() else None
    L.set_option('profile', 1)
    mon, mon_path = clocks_monitor_start(local) if rank == 0 else (None, None)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run(dstrips)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = clocks_monitor_stop(mon, mon_path) if rank == 0 else None
    L.set_option('profile', 0)
    prof = L.profile_read(local, stream, reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * units * args.steps / (ms * 1e-3)
    e_steps = max(2, min(args.steps, 5))
    run(hstrips)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        run(hstrips)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {'value': world * units * e_steps / dt, 'unit': 'matches/s', 'h2d_bytes_per_step': int(sum(a.nbytes + b.nbytes for a, b in hstrips)),
           'd2h_bytes_per_step': int(5 * 8 * units), 'steps': e_steps, 'overlaps_per_s': world * len(strips) * e_steps / dt,
           'api': 'feabas_b200.cuda.stitching_matcher(uint8 host strips, shipped YAML kwargs), one call per overlap' if stitch else
                  'feabas_b200.cuda.section_matcher(AffineMesh, AffineMesh, ArrayLoader(host float32 thumbnail) x 2, ...), one call per section pair'}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    kern = {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)} for k, v in prof.items() if v[1]}
    dom = max(kern, key=lambda k: kern[k]['ms_total'])
    xcorr_ms = sum(v['ms_total'] for v in kern.values())
    b_min = 2 * wl['h'] * wl['w'] * 4 + 20
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': units * args.steps * b_min / (kern[dom]['ms_total'] * 1e-3) / 1e9, 'peak': peak,
                'unit': 'GB/s', 'frac': units * args.steps * b_min / (kern[dom]['ms_total'] * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_kind': peak_kind,
                'note': 'latency bound: one matcher call per overlap / section pair, batches of 5-300 small blocks; B_min of the finest-level block '
                        'used as algorithmic bytes; xcorr kernels are %.0f %% of the step, the rest is host control flow + image kernels' % (100 * xcorr_ms / ms),
                'kernel_share_of_step': kern[dom]['ms_total'] / ms if world == 1 else None, 'kernels': kern}
    cpu = None
    if not args.no_cpu_baseline and world == 1:          # N = 1 only (ranks of a multi-GPU run are bound to their GPU's CPUs)
        import psutil
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        pool = ProcessPoolExecutor(cores, mp_context=mp.get_context('spawn'), initializer=_cpu_stitch_init, initargs=(1, wl))
        list(pool.map(_cpu_stitch_task, range(cores)))
        t0 = time.perf_counter()
        n_tasks = max(len(strips), cores)                      # every worker busy: the job list is cycled
        res = list(pool.map(_cpu_stitch_task, range(n_tasks)))
        s_ = time.perf_counter() - t0
        pool.shutdown()
        blocks = sum(r[1] for r in res) or sum(r[0] for r in res)
        cpu = {'value': blocks / s_, 'unit': 'matches/s', 'cores': cores, 'kind': 'port', 'overlaps_per_s': n_tasks / s_,
               'sample': f'{n_tasks} jobs (the {len(strips)} of one step, cycled), one job per task on {cores} single-thread workers '
                         f"(oracle port of {'stitching_matcher' if stitch else 'the section_matcher loop'}), {s_:.1f} s wall"}
    multi = None
    if args.workers > 0 and world == 1:
        torch.cuda.empty_cache()
        multi = multi_process_throughput(args.workload, args.workers, max(20, args.steps), local)
    line = {'metric': 'xcorr_block_matches_per_sec', 'value': value, 'unit': 'matches/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': dict(config, overlaps_per_s=world * len(strips) * args.steps / (ms * 1e-3),
                                                                   unit_count='xcorr pairs (fb_pair_count)' if pairs else 'returned match points'),
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'multi_process': multi}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
def clocks_monitor_start(gpu_index):
    path = tempfile.mktemp(prefix='fb_clocks_', suffix='.csv')
    q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    try:
        proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(gpu_index)],
                                stdout=open(path, 'w'), stderr=subprocess.DEVNULL)
    except OSError:
        return None, path
    return proc, path


def clocks_monitor_stop(proc, path):
    out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
    if proc is None:
        return out
    proc.terminate()
    try:
        proc.wait(timeout=5)
    except Exception:
        proc.kill()
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(path)
    except OSError:
        pass
    if sm:
        out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}
    return out


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full
    capture of this workload (profiles/traffic.json, written by profiles/ncu_summary.py --traffic), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)[workload][kernel]
    except Exception:
        return None


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs next to its GPU (NVML affinity mask) BEFORE host buffers are allocated, so
    that the pinned staging memory is first-touched on the GPU's NUMA node: with N ranks the end-to-end path
    is bound by host memory / PCIe, not by the kernels."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = [64 * i + b for i, wd in enumerate(mask) for b in range(64) if (wd >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0



def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:
        return 6650.0, 'fallback'


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='xcorr512', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='pairs per step per GPU (default: workload table)')
    ap.add_argument('--ws-gib', type=float, default=6.0, help='HBM workspace budget (GiB)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fast-flags', type=int, default=0, help='kernel experiment switches (fb_set_option fast_flags)')
    ap.add_argument('--workers', type=int, default=0, help='job workloads: also measure P worker processes sharing the GPU')
    ap.add_argument('--force', default=None, choices=['generic', 'staged', 'fused'], help='force an execution shape (comparison runs)')
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl['batch'] = args.batch
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    warmup = max(args.warmup, 3)
    if wl.get('kind') == 'blocks':
        return bench_blocks(args, wl, rank, world, local, warmup)
    if wl.get('kind') in ('stitch', 'sections'):
        return bench_stitch(args, wl, rank, world, local, warmup)
    h, w, pad, batch = wl['h'], wl['w'], wl['pad'], wl['batch']

    from oracle import xcorr_oracle as xo           # bench may execute oracle/ only for the CPU legs
    ny, nx = xo.fft_shape((h, w), (h, w), pad)
    config = {'workload': f'{args.workload}: batched xcorr_fft, {h}x{w} float32 blocks, pad={pad} (FFT {ny}x{nx}), '
                          f'FFT_CONF_MIRROR, subpixel=True',
              'pairs_per_step_per_gpu': batch, 'fft': [ny, nx], 'l2': 'inputs larger than L2 (no flush needed)'}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == 'reference':
        if rank != 0:
            return
        import psutil
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        # bounded sample per step: the whole --steps/--warmup run ends in about 90 s of wall clock
        per_step_wall = min(3.0, max(0.2, 90.0 / (args.steps + warmup)))
        per_worker = cpu_pairs_per_worker(wl, per_step_wall * cores, cores)
        arm = CpuArm(wl, per_worker)
        for _ in range(warmup):
            arm.step()
        pairs = secs = 0
        for _ in range(args.steps):
            p_, s_ = arm.step()
            pairs += p_; secs += s_
        arm.close()
        val = pairs / secs
        sample = f'{per_worker} pairs per worker x {arm.cores} single-thread workers per step (oracle port of xcorr_fft, scipy pocketfft)'
        print(json.dumps({
            'impl': 'reference', 'metric': 'xcorr_block_matches_per_sec', 'value': val, 'unit': 'matches/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': warmup, 'ms_per_step': 1e3 * secs / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config,
            'cpu_baseline': {'value': val, 'unit': 'matches/s', 'cores': arm.cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    config['numa_bound_cpus'] = bind_to_gpu_numa(local) if world > 1 else 0
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    L = fc._lib
    L.set_option('ws_bytes', int(args.ws_gib * (1 << 30)))
    if os.environ.get('FB_PIPELINE_WAVES'):
        L.set_option('pipeline_waves', int(os.environ['FB_PIPELINE_WAVES']))
        config['pipeline_waves'] = int(os.environ['FB_PIPELINE_WAVES'])
    if os.environ.get('FB_FUSED_THREADS'):
        L.set_option('fused_threads', int(os.environ['FB_FUSED_THREADS']))
        config['fused_threads'] = int(os.environ['FB_FUSED_THREADS'])
    if os.environ.get('FB_HOST_CHUNK_MIB'):
        L.set_option('host_chunk_bytes', int(os.environ['FB_HOST_CHUNK_MIB']) << 20)
        config['host_chunk_mib'] = int(os.environ['FB_HOST_CHUNK_MIB'])
    if os.environ.get('FB_PIPELINE'):
        L.set_option('pipeline', int(os.environ['FB_PIPELINE']))
        config['pipeline'] = int(os.environ['FB_PIPELINE'])
    if os.environ.get('FB_MAX_RADIX'):
        L.set_option('max_radix', int(os.environ['FB_MAX_RADIX']))
        config['max_radix'] = int(os.environ['FB_MAX_RADIX'])
    if args.fast_flags:
        L.set_option('fast_flags', args.fast_flags)
        config['fast_flags'] = args.fast_flags
    flags = 0x2 | (2 << 2) | (1 if pad else 0)
    info = L.plan_info(h, w, h, w, L.FB_F32, ny, nx, flags)
    fused = info['path'] == 'fused'
    config['path'] = info['path'] if not args.force else 'forced ' + args.force

    a, b, shifts = make_pairs(batch, h, w, seed=100 + rank, device=dev, max_shift=min(32, min(h, w) // 8))
    out = torch.empty((5, batch), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        fc.xcorr_fft_device(a, b, subpixel=True, pad=pad, out=out, force=args.force)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    # correctness inside the bench: the synthetic ground truth must be recovered
    res = out.cpu().numpy()
    sh = shifts.cpu().numpy()
    n_ok = int(np.sum((np.round(res[0]) == sh[:, 0]) & (np.round(res[1]) == sh[:, 1])))
    # (tiny blocks cut from noisy canvases can legitimately lock onto a different peak: allow 0.5 %)
    assert n_ok >= 0.995 * batch or args.fast_flags >= 64, f'only {n_ok}/{batch} ground-truth displacements recovered'
    config['ground_truth_recovered'] = n_ok / batch

    L.profile_read(local, stream, reset=True) if L.launch_count() else None
    L.set_option('profile', 1)
    mon, mon_path = clocks_monitor_start(local) if rank == 0 else (None, None)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.launch_count() - launches0
    clocks = clocks_monitor_stop(mon, mon_path) if rank == 0 else None
    L.set_option('profile', 0)
    prof_timed = L.profile_read(local, stream, reset=True)
    ms_timed = ms
    # The library runs a chunk as two independent parts on two streams (their kernels' ramps and tails overlap), so
    # the per-kernel CUDA-event times of the timed region overlap each other.  The per-kernel roofline numbers come
    # from a serial pass (one stream, same batch, same kernels) right after the timed region; both sets are reported.
    serial_steps = max(3, min(args.steps, 50))
    L.set_option('pipeline', 1)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    L.profile_read(local, stream, reset=True)
    L.set_option('profile', 1)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(serial_steps):
        step()
    s1.record()
    torch.cuda.synchronize()
    ms_serial_step = s0.elapsed_time(s1) / serial_steps
    L.set_option('profile', 0)
    prof = L.profile_read(local, stream, reset=True)
    L.set_option('pipeline', int(os.environ.get('FB_PIPELINE', 2)))
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * batch * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        ah, bh = a.cpu().pin_memory(), b.cpu().pin_memory()
        an, bn = ah.numpy(), bh.numpy()
        e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            r = fc.xcorr_fft(an, bn, subpixel=True, pad=pad, device=local)
        assert np.array_equal(r[0], res[0]) and np.array_equal(r[1], res[1]), 'host path and device path disagree'
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            r = fc.xcorr_fft(an, bn, subpixel=True, pad=pad, device=local)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {'value': world * batch * e_steps / dt, 'unit': 'matches/s',
               'h2d_bytes_per_step': int(an.nbytes + bn.nbytes), 'd2h_bytes_per_step': int(5 * 8 * batch),
               'steps': e_steps, 'api': 'feabas_b200.cuda.xcorr_fft(numpy pinned host arrays) -> fb_xcorr_batch_host'}
        del ah, bh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, from the CUDA-event times taken inside the timed region
    peak, peak_kind = peaks()
    b_alg, b_min, b_pass = algorithmic_bytes(h, w, ny, nx, True, fused)
    kern = {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)} for k, v in prof.items() if v[1]}
    dom = max(kern, key=lambda k: kern[k]['ms_total'])
    pairs_per_launch = batch * serial_steps / kern[dom]['launches']
    if dom == 'columns':
        bytes_per_pair = column_kernel_bytes(h, ny, nx, True)
    elif dom == 'rows_forward':
        bytes_per_pair = 2 * h * w * 4 + 2 * h * (nx // 2 + 1) * 8
    elif dom == 'rows_inverse':
        bytes_per_pair = 2 * ny * (nx // 2 + 1) * 8
    else:
        bytes_per_pair = b_min
    achieved = bytes_per_pair * pairs_per_launch / (kern[dom]['ms_per_launch'] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': measured_traffic(args.workload, dom) if batch == WORKLOADS[args.workload]['batch'] else None,
                'peak_kind': peak_kind, 'algorithmic_bytes_per_pair': bytes_per_pair,
                'pairs_per_launch': pairs_per_launch, 'ms_per_launch': kern[dom]['ms_per_launch'],
                'kernel_share_of_step': kern[dom]['ms_total'] / (ms_serial_step * serial_steps),
                'pipeline': {'bytes_per_pair': b_alg, 'b_min': b_min, 'b_pass': b_pass,
                             'achieved': value / world * b_alg / 1e9, 'frac': value / world * b_alg / 1e9 / peak},
                'kernels': kern,
                'kernel_timing': f'serial pass of {serial_steps} steps on one stream right after the timed region '
                                 f'({ms_serial_step:.4f} ms per step = {batch / ms_serial_step * 1e3:.0f} matches/s); in the timed region the '
                                 'library runs two half-batches on two streams and the per-kernel event times overlap (kernels_timed_region)',
                'serial_ms_per_step': ms_serial_step,
                'kernels_timed_region': {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)}
                                         for k, v in prof_timed.items() if v[1]}}

    cpu = None
    if not args.no_cpu_baseline and world == 1:          # N = 1 only (ranks of a multi-GPU run are bound to their GPU's CPUs)
        import psutil
        cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        per_worker = cpu_pairs_per_worker(wl, 20.0, cores)
        arm = CpuArm(wl, per_worker)
        p_, s_ = arm.step()
        arm.close()
        cpu = {'value': p_ / s_, 'unit': 'matches/s', 'cores': arm.cores, 'kind': 'port',
               'sample': f'{p_} pairs of the same workload ({per_worker} per single-thread worker, {arm.cores} workers), '
                         f'oracle port of xcorr_fft (scipy pocketfft), {s_:.1f} s wall'}

    line = {'metric': 'xcorr_block_matches_per_sec', 'value': value, 'unit': 'matches/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
