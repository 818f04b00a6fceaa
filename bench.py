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
    # BASELINE.json configs[1]: every overlap of a 20x20 montage of 3000x4000 tiles (380 horizontal + 380 vertical + 722 diagonal
    # = 1482 overlaps); one fixed job list, sharded over the ranks by overlap index (strong scaling)
    'stitch20x20': dict(kind='stitch', rows=20, cols=20, tile=(3000, 4000), overlap=0.1, margin=100, batch=0, h=74, w=67, pad=True, synth='per_overlap'),
    # BASELINE.json configs[2]: 64 consecutive sections (2048^2 thumbnails), compare_distance 2 -> 125 section pairs, sharded by pair
    'thumb64': dict(kind='sections', sections=64, pairs=125, size=2048, batch=0, h=50, w=50, pad=True, synth='stack'),
    # BASELINE.json configs[3] at reduced count: 64 section pairs (8192^2 uint8 each side), a dense grid of 512^2 blocks per pair
    # (256 blocks, sigma 3.5), sharded by pair
    'align512_pairs': dict(kind='blocks', h=512, w=512, pad=True, section=8192, sigma=3.5, batch=256, pairs=64),
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


def _blur(x, sigma):
    """Separable Gaussian, replicate border (torch, any device); x: N x 1 x H x W."""
    import torch
    import torch.nn.functional as F
    r = int(4 * sigma + 0.5)
    t = torch.arange(-r, r + 1, device=x.device, dtype=torch.float32)
    k = torch.exp(-0.5 * (t / sigma) ** 2)
    k = k / k.sum()
    x = F.conv2d(F.pad(x, (r, r, 0, 0), mode='replicate'), k.view(1, 1, 1, -1))
    return F.conv2d(F.pad(x, (0, 0, r, r), mode='replicate'), k.view(1, 1, -1, 1))


def pairs_group(h, w, max_shift):
    """Pairs per generator chunk: a function of the block shape only, so that pair i is the same array whatever the
    batch size -- the CPU arm times a prefix of the GPU arm's batch."""
    area = (h + 2 * (max_shift + 8)) * (w + 2 * (max_shift + 8))
    g = 8
    while g < 1024 and 2 * g * area <= (1 << 22):
        g *= 2
    return g


def make_pairs(n, h, w, seed, device='cpu', max_shift=32):
    """Seeded synthetic EM-like block pairs with known integer displacement: band-limited texture cut from one canvas
    per pair at two offsets + independent noise, DoG-like band-pass.  One generator stream per chunk of ``pairs_group``
    pairs, drawn with the CUDA generator when a GPU is present (``gen_device()``), so that both arms of the benchmark
    -- run on the same box -- see byte-identical arrays; the result is moved to ``device``.
    Returns (stack0, stack1, shifts)."""
    import torch
    m = max_shift + 8
    grp = pairs_group(h, w, max_shift)
    gdev = gen_device()
    a = torch.empty((n, h, w), dtype=torch.float32, device=gdev)
    b = torch.empty((n, h, w), dtype=torch.float32, device=gdev)
    shifts = torch.empty((n, 2), dtype=torch.int64)
    for lo in range(0, n, grp):
        hi = min(n, lo + grp)
        g = torch.Generator(device=gdev)
        g.manual_seed(seed * 1000003 + lo // grp)
        sh = torch.randint(-max_shift, max_shift + 1, (grp, 2), generator=g, device=gdev).cpu()
        c = torch.randn((grp, 1, h + 2 * m, w + 2 * m), generator=g, device=gdev)
        n0 = torch.randn(c.shape, generator=g, device=gdev)
        n1 = torch.randn(c.shape, generator=g, device=gdev)
        c = _blur(c, 2.0)
        c = c / c.std()
        c0, c1 = c + 0.25 * n0, c + 0.25 * n1
        c0 = _blur(c0, 2.5) - _blur(_blur(c0, 2.5), 2.5)           # DoG-like band-pass, sigma 2.5
        c1 = _blur(c1, 2.5) - _blur(_blur(c1, 2.5), 2.5)
        for i in range(lo, hi):
            dx, dy = int(sh[i - lo, 0]), int(sh[i - lo, 1])
            a[i] = c0[i - lo, 0, m:m + h, m:m + w]
            b[i] = c1[i - lo, 0, m - dy:m - dy + h, m - dx:m - dx + w]
        shifts[lo:hi] = sh[:hi - lo]
    return a.to(device), b.to(device), shifts


def gen_device():
    """Where the synthetic inputs are drawn: the GPU's generator when there is one (both arms run on the GPU box; the CPU
    generator gives different -- equally valid -- numbers and is only used where no GPU exists)."""
    import torch
    return 'cuda' if torch.cuda.is_available() else 'cpu'


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's xcorr_fft on the host cores
# ----------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(path0, path1, pad, per_worker):
    os.environ['OMP_NUM_THREADS'] = '1'                       # reference: config.py:301-310
    import numpy as np
    _W['a'], _W['b'] = np.load(path0, mmap_mode='r'), np.load(path1, mmap_mode='r')
    _W['pad'], _W['per_worker'] = pad, per_worker
    from oracle import xcorr_oracle as xo
    _W['f'] = xo.xcorr_oracle
    _W['f'](np.asarray(_W['a'][:1]), np.asarray(_W['b'][:1]), subpixel=True, pad=pad)  # warm


def _cpu_task(k):
    import numpy as np
    n, per = _W['a'].shape[0], _W['per_worker']
    idx = (k * per + np.arange(per)) % n                      # worker k's slice of the shared sample (cycled)
    a, b = np.ascontiguousarray(_W['a'][idx]), np.ascontiguousarray(_W['b'][idx])
    t = time.perf_counter()
    _W['f'](a, b, conf_mode=2, subpixel=True, pad=_W['pad'])
    return time.perf_counter() - t


class CpuArm:
    """Process pool, one single-threaded worker per physical core (the reference's own parallel model:
    feabas/concurrent.py:59-96), each running the oracle port of ``xcorr_fft`` on its slice of a sample of the GPU arm's
    own arrays (``make_pairs`` with the same seed: the first pairs of rank 0's batch), shared through /dev/shm."""

    def __init__(self, wl, per_worker, seed):
        import numpy as np
        import psutil
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        self.cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
        self.per_worker = per_worker
        n = min(wl['batch'], self.cores * per_worker)
        a, b, _ = make_pairs(n, wl['h'], wl['w'], seed, 'cpu', max_shift=min(32, min(wl['h'], wl['w']) // 8))
        shm = '/dev/shm' if os.path.isdir('/dev/shm') else tempfile.gettempdir()
        self.paths = [os.path.join(shm, f'fb_bench_{os.getpid()}_{i}.npy') for i in range(2)]
        np.save(self.paths[0], a.cpu().numpy())
        np.save(self.paths[1], b.cpu().numpy())
        self.sample_pairs = n
        self.pool = ProcessPoolExecutor(self.cores, mp_context=mp.get_context('spawn'),
                                        initializer=_cpu_init, initargs=(self.paths[0], self.paths[1], wl['pad'], per_worker))
        list(self.pool.map(_cpu_task, range(self.cores)))      # spawn + warm every worker

    def step(self):
        """One timed pass: every worker processes its pairs once.  Returns (pairs, seconds)."""
        t = time.perf_counter()
        list(self.pool.map(_cpu_task, range(self.cores)))
        return self.cores * self.per_worker, time.perf_counter() - t

    def close(self):
        self.pool.shutdown()
        for p_ in self.paths:
            try:
                os.unlink(p_)
            except OSError:
                pass


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
    n_pairs = int(wl.get('pairs', 0))                          # > 0: ONE list of section pairs, sharded over the ranks (strong scaling)
    strong = n_pairs > 0
    config = {'workload': f'{args.workload}: bboxes_mesh_renderer_matcher on ' + (f'{n_pairs} ' if strong else 'a ') +
                          f'{size}x{size} uint8 section pair' + ('s' if strong else '') + f', {nby}x{nbx} grid of '
                          f'{h}x{w} blocks per pair, sigma={sigma} (masked DoG), pad={pad} (FFT {ny}x{nx}), FFT_CONF_MIRROR, subpixel=True',
              'block_pairs_per_step': batch * n_pairs if strong else batch, 'fft': [ny, nx], 'l2': 'inputs larger than L2 (no flush needed)'}
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
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': val, 'unit': 'matches/s', 'cores': cores, 'kind': 'port',
                             'sample': sample + '; white-noise uint8 sections of the same shapes (timing-neutral for FFT / filter work)'},
            'e2e': {'value': val, 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    run_info = {'numa_bound_cpus': bind_to_gpu_numa(local) if world > 1 else 0}
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    from feabas_b200.cuda import shard
    L = fc._lib
    L.set_option('ws_bytes', int(args.ws_gib * (1 << 30)))
    shift = (7, -5)
    if strong:
        lo_p, hi_p = shard.shard_range(n_pairs, world, rank)
        secs_dev = [make_section_pair(size, 300 + k, dev, shift) for k in range(lo_p, hi_p)]
    else:
        secs_dev = [make_section_pair(size, 300 + rank, dev, shift)]
    run_info['section_pairs_this_rank'] = len(secs_dev)
    a, b = secs_dev[0] if secs_dev else make_section_pair(size, 300, dev, shift)
    m0 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=0)
    m1 = fc.AffineMesh.from_bbox((0, 0, size, size), cartesian=True, uid=1)
    boxes = np.array([(x * w, y * h, x * w + w, y * h + h) for y in range(nby) for x in range(nbx)], dtype=np.float64)
    loaders = [(fc.ArrayLoader(x, device=local), fc.ArrayLoader(y, device=local)) for x, y in secs_dev]
    stream = torch.cuda.current_stream().cuda_stream
    batch_rank = batch * len(secs_dev)

    def step():
        if strong:       # the job list of this rank through the pipelined entry point (next pair enqueued before this one is read back)
            return fc.bboxes_mesh_renderer_matcher_many(((m0, m1, l0, l1, boxes, boxes) for l0, l1 in loaders), **kw)[-1]
        l0, l1 = loaders[0]
        return fc.bboxes_mesh_renderer_matcher(m0, m1, l0, l1, boxes, boxes, **kw)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return v

    for _ in range(warmup):
        xy0, xy1, conf = step()
    d = xy1 - xy0                                               # every block must recover the section offset
    n_ok = int(np.sum((np.round(d[:, 0]) == shift[0]) & (np.round(d[:, 1]) == shift[1])))
    assert n_ok >= 0.99 * batch, f'only {n_ok}/{batch} blocks recovered the offset {shift}: {d[:4]}'
    run_info['ground_truth_recovered'] = n_ok / batch
    units = allsum(float(batch_rank))                           # block pairs per step, all ranks
    L.profile_read(local, 'all', reset=True) if L.launch_count() else None
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
    prof = L.profile_read(local, 'all', reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = units * args.steps / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        host = [(x.cpu().pin_memory(), y.cpu().pin_memory()) for x, y in secs_dev]
        e_steps = max(2, min(args.steps, 10 if len(host) <= 1 else 3))

        def e_step():
            # the call a user makes: host images in, matches out (upload of both sections inside); with a list of pairs the
            # match lists of all ranks are gathered on the host (shard.gather_concat)
            if strong:
                outs = fc.bboxes_mesh_renderer_matcher_many(((m0, m1, fc.ArrayLoader(ah, device=local), fc.ArrayLoader(bh, device=local), boxes, boxes)
                                                             for ah, bh in host), **kw)
            else:
                outs = [fc.bboxes_mesh_renderer_matcher(m0, m1, fc.ArrayLoader(ah, device=local), fc.ArrayLoader(bh, device=local), boxes, boxes, **kw)
                        for ah, bh in host]
            if strong:
                cat = [np.concatenate([o[i] for o in outs], axis=0) if outs else None for i in range(3)]
                return shard.gather_concat(tuple(cat))
            return outs[-1]
        for _ in range(2):
            r = e_step()
        if strong:
            assert r[0].shape[0] == batch * n_pairs, 'gathered match list is incomplete'
            d = r[1] - r[0]
            assert np.mean((np.round(d[:, 0]) == shift[0]) & (np.round(d[:, 1]) == shift[1])) >= 0.99
        else:
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
        e2e = {'value': units * e_steps / dt, 'unit': 'matches/s',
               'h2d_bytes_per_step': int(allsum(float(sum(x.numel() + y.numel() for x, y in host)))), 'd2h_bytes_per_step': int(5 * 8 * units), 'steps': e_steps,
               'api': ('feabas_b200.cuda.bboxes_mesh_renderer_matcher_many(jobs of ArrayLoader(pinned uint8 host sections))' if strong else
                       'feabas_b200.cuda.bboxes_mesh_renderer_matcher(ArrayLoader(pinned uint8 host sections))')
                      + ' -> fb_crop_blocks + fb_masked_dog + fb_xcorr_batch_device'
                      + (' per section pair of this rank (two pairs in flight), + shard.gather_concat of the match lists' if strong else '')}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = peaks()
    b_alg, b_min, b_pass = algorithmic_bytes(h, w, ny, nx, True, False)
    kern = {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)} for k, v in prof.items() if v[1]}
    dom = max(kern, key=lambda k: kern[k]['ms_total'])
    pairs_per_launch = batch_rank * args.steps / kern[dom]['launches']
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
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu, 'run_info': run_info}
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


def montage_overlaps(rows, cols):
    """Overlaps of a rows x cols montage in tile order: (kind, r, c) with kind 'H' (left-right neighbours), 'V'
    (top-bottom), 'D' / 'E' (the two diagonals).  20 x 20 -> 380 + 380 + 722 = 1482 (feabas/stitcher.py:418-437)."""
    out = []
    for r in range(rows):
        for c in range(cols):
            if c + 1 < cols:
                out.append(('H', r, c))
            if r + 1 < rows:
                out.append(('V', r, c))
            if r + 1 < rows and c + 1 < cols:
                out.append(('D', r, c))
                out.append(('E', r, c))
    return out


def make_overlap_job(wl, k, seed, jitter=10):
    """Overlap k of the montage as its two uint8 strips, cut the way feabas/stitcher.py:556-568 cuts them: the overlap
    rectangle grown by ``margin`` and clipped to each tile, so the two strips extend to opposite sides.  Every overlap has
    its own canvas (generator seeded by (seed, k)): any rank can make any job."""
    import torch
    th, tw = wl['tile']
    m = wl['margin']
    ow, oh = int(round(tw * wl['overlap'])), int(round(th * wl['overlap']))
    kind = montage_overlaps(wl['rows'], wl['cols'])[k][0]
    # strip extents (x0, x1, y0, y1) in overlap coordinates for tile i and tile j
    if kind == 'H':
        wd, ht = ow, th
        e0, e1 = (-m, wd, 0, ht), (0, wd + m, 0, ht)
    elif kind == 'V':
        wd, ht = tw, oh
        e0, e1 = (0, wd, -m, ht), (0, wd, 0, ht + m)
    elif kind == 'D':
        wd, ht = ow, oh
        e0, e1 = (-m, wd, -m, ht), (0, wd + m, 0, ht + m)
    else:
        wd, ht = ow, oh
        e0, e1 = (0, wd + m, -m, ht), (-m, wd, 0, ht + m)
    gdev = gen_device()
    g = torch.Generator(device=gdev)
    g.manual_seed(seed * 1000003 + k)
    jx, jy = (int(v) for v in torch.randint(-jitter, jitter + 1, (2,), generator=g, device=gdev).cpu())
    pad = m + jitter + 8
    c = torch.randn((1, 1, ht + 2 * pad, wd + 2 * pad), generator=g, device=gdev)
    c = _blur(c, 2.5)
    c = c / c.std()

    def cut(ext, dx, dy):
        x0, x1, y0, y1 = ext
        v = c[0, 0, pad + y0 + dy:pad + y1 + dy, pad + x0 + dx:pad + x1 + dx]
        v = v + 0.2 * torch.randn(v.shape, generator=g, device=gdev)
        return (v * 40 + 128).clamp_(0, 255).to(torch.uint8).contiguous()
    return cut(e0, 0, 0), cut(e1, jx, jy)


def make_section_stack(wl, seed, lo, hi):
    """Band-passed float32 thumbnails of sections lo .. hi-1 of a stack: one canvas seen through a slow random walk of
    the field of view + independent noise per section (neighbours correlate, far sections drift apart)."""
    import torch
    size = wl['size']
    gdev = gen_device()
    g = torch.Generator(device=gdev)
    g.manual_seed(seed * 1000003)
    pad = 6 * wl['sections'] // 2 + 16
    base = _blur(torch.randn((1, 1, size + 2 * pad, size + 2 * pad), generator=g, device=gdev), 3.0)
    base = base / base.std()
    walk = torch.cumsum(torch.randint(-3, 4, (wl['sections'], 2), generator=g, device=gdev), dim=0).cpu()
    out = []
    for s_ in range(lo, hi):
        gs = torch.Generator(device=gdev)
        gs.manual_seed(seed * 1000003 + 1 + s_)
        dx, dy = int(walk[s_, 0]), int(walk[s_, 1])
        v = base[:, :, pad + dy:pad + dy + size, pad + dx:pad + dx + size]
        v = v + 0.2 * torch.randn(v.shape, generator=gs, device=gdev)
        v = _blur(v, 3.5) - _blur(_blur(v, 3.5), 3.5)
        out.append((40 * v[0, 0]).contiguous())
    return out


def section_pairs(n_sections, distance=2):
    """(i, j) of every pair of sections at most ``distance`` apart, in order (thumbnail_main.py compare_distance)."""
    return [(i, i + d) for i in range(n_sections) for d in range(1, distance + 1) if i + d < n_sections]


def make_jobs(wl, seed, lo=0, hi=None):
    """Jobs lo .. hi-1 of the workload's fixed job list as (array0, array1) pairs (torch tensors on the generator's device
    for the per-overlap / stack generators, numpy arrays for the small tile-cut workloads)."""
    if wl.get('synth') == 'per_overlap':
        n = len(montage_overlaps(wl['rows'], wl['cols']))
        hi = n if hi is None else hi
        return [make_overlap_job(wl, k, seed) for k in range(lo, hi)]
    if wl.get('synth') == 'stack':
        pairs = section_pairs(wl['sections'])
        hi = len(pairs) if hi is None else hi
        need = sorted({i for p_ in pairs[lo:hi] for i in p_})
        if not need:
            return []
        secs = dict(zip(range(need[0], need[-1] + 1), make_section_stack(wl, seed, need[0], need[-1] + 1)))
        return [(secs[i], secs[j]) for i, j in pairs[lo:hi]]
    jobs = make_overlap_strips(wl, seed) if wl['kind'] == 'stitch' else make_section_thumbs(wl, seed)
    return jobs[lo:len(jobs) if hi is None else hi]


def job_count(wl):
    if wl.get('synth') == 'per_overlap':
        return len(montage_overlaps(wl['rows'], wl['cols']))
    if wl.get('synth') == 'stack':
        return len(section_pairs(wl['sections']))
    if wl['kind'] == 'stitch':
        return len(make_overlap_strips(dict(wl, tile=(300, 400), margin=10), 1))     # count only
    return wl['pairs']


def _np(x):
    return x if not hasattr(x, 'cpu') else x.cpu().numpy()


def _cpu_stitch_init(seed, wl, n_jobs):
    os.environ['OMP_NUM_THREADS'] = '1'
    _W['wl'], _W['seed'], _W['n_jobs'], _W['cache'] = wl, seed, n_jobs, {}
    from oracle import matcher_oracle as mo
    _W['f'] = mo.stitching_oracle
    _W['mo'] = mo


def _cpu_job_inputs(k):
    """Job k of the fixed list (the same arrays the GPU arm works on), fetched from the shared sample file."""
    import numpy as np
    if k not in _W['cache']:
        with np.load(os.path.join(_W['wl']['_sample_dir'], f'job_{k}.npz')) as z:
            _W['cache'][k] = (z['a'], z['b'])
    return _W['cache'][k]


def _cpu_stitch_task(k):
    wl = _W['wl']
    a, b = _cpu_job_inputs(wl['_sample_ids'][k % len(wl['_sample_ids'])])
    trace = []
    if wl['kind'] == 'stitch':
        kw = {k_: v for k_, v in STITCH_KW.items() if k_ not in ('compute_photometric',)}
        out = _W['f'](a, b, trace=trace, **kw)
    else:
        mo = _W['mo']
        sec0 = mo._Section((-0.5, -0.5, a.shape[1] - 0.5, a.shape[0] - 0.5), locked=True)
        sec1 = mo._Section((-0.5, -0.5, b.shape[1] - 0.5, b.shape[0] - 0.5))
        out = mo.surrogate_loop_oracle(sec0, sec1, a, b, SECTION_KW['spacings'], conf_thresh=SECTION_KW['conf_thresh'],
                                       residue_mode='huber', residue_len=SECTION_KW['residue_len'], pad=True, batch_size=SECTION_KW['batch_size'],
                                       trace=trace)
    blocks = sum(t.get('nblocks', 0) + (1 if 'coarse' in t else 0) for t in trace)
    return 0 if out[0] is None else len(out[0]), blocks


def cpu_jobs_arm(wl, seed, n_sample):
    """Spawned single-thread workers running the oracle port of the matcher loop on a SAMPLE of the fixed job list: jobs
    spread evenly over the list (every kind of overlap is represented in proportion), written once to shared memory."""
    import numpy as np
    import psutil
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    cores = psutil.cpu_count(logical=False) or os.cpu_count() or 1
    total = job_count(wl)
    m_ = min(n_sample, total)
    ids = sorted({int(round(i * (total - 1) / max(m_ - 1, 1))) for i in range(m_)})
    shm = '/dev/shm' if os.path.isdir('/dev/shm') else tempfile.gettempdir()
    sdir = tempfile.mkdtemp(prefix='fb_bench_jobs_', dir=shm)
    if wl.get('synth'):
        for k in ids:
            a, b = make_jobs(wl, seed, k, k + 1)[0]
            np.savez(os.path.join(sdir, f'job_{k}.npz'), a=_np(a), b=_np(b))
    else:
        jobs = make_jobs(wl, seed)
        for k in ids:
            np.savez(os.path.join(sdir, f'job_{k}.npz'), a=_np(jobs[k][0]), b=_np(jobs[k][1]))
    wl = dict(wl, _sample_dir=sdir, _sample_ids=ids)
    pool = ProcessPoolExecutor(cores, mp_context=mp.get_context('spawn'), initializer=_cpu_stitch_init, initargs=(seed, wl, total))
    list(pool.map(_cpu_stitch_task, range(cores)))            # spawn + warm
    return pool, cores, ids, sdir


def _gpu_job_worker(workload, seed, steps, local, barrier, queue, part, parts):
    """One of P worker processes sharing a GPU (FEABAS's own parallel model: spawned workers, feabas/concurrent.py:59-96):
    every worker has its own CUDA context, takes the contiguous range `part` of `parts` of the fixed job list and
    advances it in lockstep (``stitching_matcher_many`` / ``section_matcher_many``), `steps` times."""
    try:
        import torch
        torch.cuda.set_device(local)
        import feabas_b200.cuda as fc
        from feabas_b200.cuda import shard
        wl = dict(WORKLOADS[workload])
        stitch = wl['kind'] == 'stitch'
        lo, hi = shard.shard_range(job_count(wl), parts, part)
        jobs = [(_np(a), _np(b)) for a, b in make_jobs(wl, seed, lo, hi)]
        lib = fc._lib.lib()

        def run():
            if stitch:
                fc.stitching_matcher_many(jobs, device=local, **STITCH_KW)
            else:
                fc.section_matcher_many([(fc.AffineMesh.from_bbox((0, 0, a.shape[1], a.shape[0]), uid=0),
                                          fc.AffineMesh.from_bbox((0, 0, a.shape[1], a.shape[0]), uid=1),
                                          fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local)) for a, b in jobs], **SECTION_KW)
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
    procs = [ctx.Process(target=_gpu_job_worker, args=(workload, 1, steps, local, barrier, queue, k, workers)) for k in range(workers)]
    for pr in procs:
        pr.start()
    try:
        barrier.wait(timeout=600)
    except Exception:
        for pr in procs:
            pr.terminate()
        return {'workers': workers, 'error': 'workers did not reach the start barrier'}
    t0 = time.perf_counter()
    res = [queue.get(timeout=900) for _ in procs]
    dt = time.perf_counter() - t0
    for pr in procs:
        pr.join(timeout=60)
    if any(r[0] == 'error' for r in res):
        return {'workers': workers, 'error': str([r for r in res if r[0] == 'error'][0][1])}
    return {'workers': workers, 'value': sum(r[0] for r in res) / dt, 'unit': 'matches/s', 'jobs_per_s': sum(r[1] for r in res) / dt,
            'steps_per_worker': steps, 'seconds': dt,
            'note': 'wall clock over P spawned worker processes sharing this GPU (one CUDA context each, host arrays in), every worker '
                    'advancing its contiguous share of the job list in lockstep: the coarse-to-fine loops are host (Python) bound in one '
                    'process, so FEABAS\'s own worker fan-out (feabas/concurrent.py) still pays on one GPU; supplementary'}


def _pack_results(results):
    """List of matcher results -> (counts, xy0, xy1, weight, strain) arrays for the host-side gather."""
    import numpy as np
    counts = np.array([0 if r[0] is None else len(r[0]) for r in results], dtype=np.int64)
    keep = [r for r in results if r[0] is not None]
    if keep:
        xy0 = np.concatenate([np.asarray(r[0], dtype=np.float64).reshape(-1, 2) for r in keep], axis=0)
        xy1 = np.concatenate([np.asarray(r[1], dtype=np.float64).reshape(-1, 2) for r in keep], axis=0)
        wt = np.concatenate([np.asarray(r[2], dtype=np.float64).reshape(-1) for r in keep], axis=0)
    else:
        xy0, xy1, wt = np.empty((0, 2)), np.empty((0, 2)), np.empty(0)
    strain = np.array([np.nan if (r[0] is None or r[3] is None) else float(r[3]) for r in results], dtype=np.float64)
    return counts, xy0, xy1, wt, strain


def bench_jobs(args, wl, rank, world, local, warmup):
    """Job workloads (configs[0]-[2]): ONE fixed list of overlaps / section pairs, rank r takes the contiguous index range
    ``shard_range(J, world, r)`` (strong scaling, no data-path collective), every rank advances its jobs in lockstep through
    ``stitching_matcher_many`` / ``section_matcher_many``, and the match lists are gathered on the host."""
    import numpy as np
    stitch = wl['kind'] == 'stitch'
    total = job_count(wl)
    if stitch:
        desc = (f"stitching_matcher on every overlap of a {wl['rows']}x{wl['cols']} montage of {wl['tile'][0]}x{wl['tile'][1]} uint8 tiles "
                f"({total} overlaps), {int(100 * wl['overlap'])} % overlap, margin {wl['margin']}, shipped YAML kwargs (sigma 2.5, coarse 0.5, pad, "
                'conf_thresh 0.33)')
    else:
        desc = (f"section_matcher on {total} pairs of {wl['size']}x{wl['size']} float32 band-passed thumbnails"
                + (f" ({wl['sections']} consecutive sections, compare distance 2)" if wl.get('sections') else '')
                + ', spacings [150, 50], pad, conf_thresh 0.35, huber residue 3 (affine stand-in mesh)')
    config = {'workload': f'{args.workload}: {desc}; unit = block match (one xcorr of one block pair)', 'jobs_per_step': total,
              'l2': 'per-level working sets range from below to far above L2; no flush (inputs of a step exceed L2 for the large lists)'}
    seed = 1
    if args.impl == 'reference':
        if rank != 0:
            return
        n_sample = max(16, min(64, total))
        pool, cores, ids, sdir = cpu_jobs_arm(wl, seed, n_sample)
        steps = max(1, min(args.steps, 3))
        n_tasks = max(len(ids), cores)                         # every worker busy: the sample is cycled
        t0 = time.perf_counter()
        res = []
        for _ in range(steps):
            res += list(pool.map(_cpu_stitch_task, range(n_tasks)))
        secs = time.perf_counter() - t0
        pool.shutdown()
        import shutil
        shutil.rmtree(sdir, ignore_errors=True)
        blocks = sum(r[1] for r in res) or sum(r[0] for r in res)
        val = blocks / secs
        sample = (f'{steps} passes over {n_tasks} jobs ({len(ids)} jobs spread evenly over the list of {total}, cycled; the same arrays the GPU '
                  f"arm works on), one job per task on {cores} single-thread workers (oracle port of "
                  f"{'stitching_matcher' if stitch else 'the section_matcher loop'})")
        print(json.dumps({
            'impl': 'reference', 'metric': 'xcorr_block_matches_per_sec', 'value': val, 'unit': 'matches/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': val, 'unit': 'matches/s', 'cores': cores, 'kind': 'port', 'sample': sample, 'jobs_per_s': n_tasks * steps / secs},
            'e2e': {'value': val, 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    run_info = {'numa_bound_cpus': bind_to_gpu_numa(local) if world > 1 else 0, 'inputs': f'{gen_device()} generator, seed {seed}'}
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    from feabas_b200.cuda import shard
    L = fc._lib
    lo, hi = shard.shard_range(total, world, rank)
    jobs = make_jobs(wl, seed, lo, hi)
    djobs = [(torch.as_tensor(a).to(dev), torch.as_tensor(b).to(dev)) for a, b in jobs]
    hjobs = [(_np(a), _np(b)) for a, b in jobs]
    run_info['jobs_this_rank'] = len(jobs)
    stream = torch.cuda.current_stream().cuda_stream
    lib = L.lib()

    def meshes(a):
        hh, ww = a.shape
        return fc.AffineMesh.from_bbox((0, 0, ww, hh), uid=0), fc.AffineMesh.from_bbox((0, 0, ww, hh), uid=1)

    def run(pairs):
        x0 = lib.fb_pair_count()
        if stitch:
            out = fc.stitching_matcher_many(pairs, device=local, **STITCH_KW)
            out = [r[:4] for r in out]
        else:
            out = fc.section_matcher_many([meshes(a) + (fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local)) for a, b in pairs],
                                          **SECTION_KW)
        return out, lib.fb_pair_count() - x0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def allsum(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return v

    for _ in range(warmup):
        results, pairs = run(djobs)
    units = allsum(pairs)                                      # xcorr block pairs per step, all ranks
    L.profile_read(local, stream, reset=True) if L.launch_count() else None
    L.set_option('profile', 1)
    mon, mon_path = clocks_monitor_start(local) if rank == 0 else (None, None)
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        run(djobs)
    e1.record()
    barrier()
    ms = allmax(e0.elapsed_time(e1))
    launches = L.launch_count() - launches0
    clocks = clocks_monitor_stop(mon, mon_path) if rank == 0 else None
    L.set_option('profile', 0)
    prof = L.profile_read(local, stream, reset=True)
    value = units * args.steps / (ms * 1e-3)
    # ---- end to end: host arrays in (uploaded inside the call), match lists out and gathered on rank 0's host
    e_steps = max(2, min(args.steps, 5))
    run(hjobs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        results, _ = run(hjobs)
        gathered = shard.gather_concat(_pack_results(results))
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    h2d = allsum(float(sum(a.nbytes + b.nbytes for a, b in hjobs)))
    e2e = {'value': units * e_steps / dt, 'unit': 'matches/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(5 * 8 * units),
           'steps': e_steps, 'jobs_per_s': total * e_steps / dt,
           'api': ('feabas_b200.cuda.stitching_matcher_many(uint8 host strips of this rank\'s overlap range, shipped YAML kwargs)' if stitch else
                   'feabas_b200.cuda.section_matcher_many(AffineMesh x 2, ArrayLoader(host float32 thumbnail) x 2 per pair)') +
                  ' + shard.gather_concat of the match lists (host, all_gather_object)'}
    # ---- the gathered list equals what ONE GPU computes: rank 0 re-runs a sample of jobs from every shard, one call per job
    counts, gx0, gx1, gw, gstrain = gathered
    assert counts.shape[0] == total, f'gathered {counts.shape[0]} jobs, expected {total}'
    if rank == 0:
        offs = np.concatenate(([0], np.cumsum(counts)))
        rng = np.random.default_rng(0)
        check = sorted(set(rng.integers(0, total, size=min(12, total)).tolist()) | {0, total - 1})
        for k in check:
            a, b = make_jobs(wl, seed, k, k + 1)[0]
            a, b = _np(a), _np(b)
            if stitch:
                want = fc.stitching_matcher(a, b, device=local, **STITCH_KW)[:4]
            else:
                want = fc.section_matcher(*meshes(a), fc.ArrayLoader(a, device=local), fc.ArrayLoader(b, device=local), **SECTION_KW)
            if want[0] is None:
                assert counts[k] == 0
                continue
            assert counts[k] == len(want[0]), f'job {k}: {counts[k]} matches gathered, {len(want[0])} from the single-job call'
            assert np.array_equal(gx0[offs[k]:offs[k + 1]], want[0]) and np.array_equal(gx1[offs[k]:offs[k + 1]], want[1]) and \
                np.array_equal(gw[offs[k]:offs[k + 1]], np.asarray(want[2], dtype=np.float64)), f'job {k}: gathered matches differ from the single-job call'
        run_info['gather_verified_jobs'] = len(check)
        run_info['match_points_per_step'] = int(counts.sum())
        run_info['jobs_without_matches'] = int(np.sum(counts == 0))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    kern = {k: {'ms_total': v[0], 'launches': v[1], 'ms_per_launch': (v[0] / v[1] if v[1] else None)} for k, v in prof.items() if v[1]}
    dom = max(kern, key=lambda k: kern[k]['ms_total'])
    xcorr_ms = sum(v['ms_total'] for v in kern.values())
    b_min = 2 * wl['h'] * wl['w'] * 4 + 20
    ach = pairs * args.steps * b_min / (kern[dom]['ms_total'] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': None, 'peak_kind': peak_kind,
                'note': 'rank 0: B_min of the finest-level block (on-chip fused kernel) x the block pairs of this rank / the time in that kernel; '
                        'xcorr kernels are %.0f %% of the step, the rest is image kernels + host control flow of the coarse-to-fine loops' % (100 * xcorr_ms / ms),
                'kernel_share_of_step': kern[dom]['ms_total'] / ms, 'xcorr_share_of_step': xcorr_ms / ms, 'kernels': kern}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        pool, cores, ids, sdir = cpu_jobs_arm(wl, seed, max(16, min(32, total)))
        n_tasks = max(len(ids), cores)
        t0 = time.perf_counter()
        res = list(pool.map(_cpu_stitch_task, range(n_tasks)))
        s_ = time.perf_counter() - t0
        pool.shutdown()
        import shutil
        shutil.rmtree(sdir, ignore_errors=True)
        blocks = sum(r[1] for r in res) or sum(r[0] for r in res)
        cpu = {'value': blocks / s_, 'unit': 'matches/s', 'cores': cores, 'kind': 'port', 'jobs_per_s': n_tasks / s_,
               'sample': f'{n_tasks} jobs ({len(ids)} spread evenly over the list of {total}, cycled; the arrays of this run), one job per task on {cores} '
                         f"single-thread workers (oracle port of {'stitching_matcher' if stitch else 'the section_matcher loop'}), {s_:.1f} s wall"}
    multi = None
    if args.workers > 0 and world == 1:
        del djobs
        torch.cuda.empty_cache()
        multi = multi_process_throughput(args.workload, args.workers, max(3, args.steps), local)
    line = {'metric': 'xcorr_block_matches_per_sec', 'value': value, 'unit': 'matches/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic', 'config': config, 'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline,
            'cpu_baseline': cpu, 'multi_process': multi,
            'run_info': dict(run_info, jobs_per_s=total * args.steps / (ms * 1e-3), unit_count='xcorr pairs (fb_pair_count), summed over ranks',
                             launches_counted='rank 0')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
class _ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled in-process through NVML every ~2 ms (>= 200 Hz) while the
    timed region runs: a 30 ms burst still gets a dozen samples (nvidia-smi -lms 100 saw none)."""

    def __init__(self, gpu_index):
        import threading
        self.sm, self.reasons, self.mx, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv, h = self._nv, self._h
        names = (('hw_slowdown', 'nvmlClocksEventReasonHwSlowdown', 'nvmlClocksThrottleReasonHwSlowdown'),
                 ('hw_thermal_slowdown', 'nvmlClocksEventReasonHwThermalSlowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown'),
                 ('sw_thermal_slowdown', 'nvmlClocksEventReasonSwThermalSlowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown'),
                 ('sw_power_cap', 'nvmlClocksEventReasonSwPowerCap', 'nvmlClocksThrottleReasonSwPowerCap'))
        bits = [(n_, getattr(nv, a_, None) or getattr(nv, b_, 0)) for n_, a_, b_ in names]
        query = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(query(h))
                for n_, bit in bits:
                    if bit and mask & bit:
                        self.reasons.add(n_)
                if len(self.sm) % 16 == 0:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.002)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.sm:
            return {'sm_mhz': None, 'sm_max_mhz': self.mx, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(self.sm), 'sm_max_mhz': self.mx, 'reasons': sorted(self.reasons), 'samples': len(self.sm),
                'sm_mhz_min': min(self.sm), 'power_w_max': max(self.power) if self.power else None, 'how': 'NVML in-process, ~2 ms period'}


def clocks_monitor_start(gpu_index):
    return _ClockSampler(gpu_index), None


def clocks_monitor_stop(sampler, _path=None):
    if sampler is None:
        return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
    return sampler.stop()


def h2d_ceiling(dev, world, dist, mib=256, reps=8):
    """What the host -> device path of this box delivers with ALL ranks copying at once: one pinned buffer and one
    cudaMemcpyAsync stream per rank, started together; GB/s summed over ranks (max-over-ranks time).  The end-to-end
    numbers are bound by this, not by a kernel."""
    import torch
    buf = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(mib << 20, dtype=torch.uint8, device=dev)
    dst.copy_(buf, non_blocking=True)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):                                   # best of three trials (the first ones still fault pages in on some boxes)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            dst.copy_(buf, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        best = max(best, world * reps * (mib << 20) / dt / 1e9)
    return best


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
    ap.add_argument('--force', default=None, choices=['generic', 'staged', 'fused', 'fused_smem'], help='force an execution shape (comparison runs)')
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
        return bench_jobs(args, wl, rank, world, local, warmup)
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
        arm = CpuArm(wl, per_worker, seed=100)
        for _ in range(warmup):
            arm.step()
        pairs = secs = 0
        for _ in range(args.steps):
            p_, s_ = arm.step()
            pairs += p_; secs += s_
        arm.close()
        val = pairs / secs
        sample = (f'{per_worker} pairs per worker x {arm.cores} single-thread workers per step: the first {arm.sample_pairs} pairs of the GPU '
                  f"arm's batch (same generator and seed, {gen_device()} generator), oracle port of xcorr_fft (scipy pocketfft)")
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
    run_info = {'numa_bound_cpus': bind_to_gpu_numa(local) if world > 1 else 0, 'inputs': f'{gen_device()} generator, seed 100 + rank'}
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import feabas_b200.cuda as fc
    L = fc._lib
    L.set_option('ws_bytes', int(args.ws_gib * (1 << 30)))
    for env, opt, scale in (('FB_PIPELINE_WAVES', 'pipeline_waves', 1), ('FB_FUSED_THREADS', 'fused_threads', 1),
                            ('FB_HOST_CHUNK_MIB', 'host_chunk_bytes', 1 << 20), ('FB_PIPELINE', 'pipeline', 1), ('FB_MAX_RADIX', 'max_radix', 1)):
        if os.environ.get(env):
            L.set_option(opt, int(os.environ[env]) * scale)
            run_info[opt] = int(os.environ[env])
    if args.fast_flags:
        L.set_option('fast_flags', args.fast_flags)
        run_info['fast_flags'] = args.fast_flags
    flags = 0x2 | (2 << 2) | (1 if pad else 0)
    info = L.plan_info(h, w, h, w, L.FB_F32, ny, nx, flags)
    fused = info['path'].startswith('fused')
    run_info['path'] = info['path'] if not args.force else 'forced ' + args.force

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

    def timed(steps):
        """``steps`` passes, device time between two events on the launching stream, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        ms_ = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    # correctness inside the bench: the synthetic ground truth must be recovered
    res = out.cpu().numpy()
    sh = shifts.cpu().numpy()
    n_ok = int(np.sum((np.round(res[0]) == sh[:, 0]) & (np.round(res[1]) == sh[:, 1])))
    # (tiny blocks cut from noisy canvases can legitimately lock onto a different peak: allow 0.5 %)
    assert n_ok >= 0.995 * batch or args.fast_flags >= 64, f'only {n_ok}/{batch} ground-truth displacements recovered'
    run_info['ground_truth_recovered'] = n_ok / batch

    L.profile_read(local, stream, reset=True) if L.launch_count() else None
    L.set_option('profile', 1)
    mon, mon_path = clocks_monitor_start(local) if rank == 0 else (None, None)
    launches0 = L.launch_count()
    ms = timed(args.steps)
    launches = L.launch_count() - launches0
    clocks = clocks_monitor_stop(mon, mon_path) if rank == 0 else None
    L.set_option('profile', 0)
    prof_timed = L.profile_read(local, stream, reset=True)
    value = world * batch * args.steps / (ms * 1e-3)
    # the other regime beside the timed one: a short run is a burst at full clocks, a 300-step run settles under the
    # board's power cap (sw_power_cap, ~1.8 GHz): both are reported
    other_steps = 300 if args.steps < 100 else 20
    ms_other = timed(other_steps)
    regimes = {('sustained' if other_steps == 300 else 'burst'): {'steps': other_steps, 'value': world * batch * other_steps / (ms_other * 1e-3),
                                                                  'ms_per_step': ms_other / other_steps},
               ('burst' if other_steps == 300 else 'sustained'): {'steps': args.steps, 'value': value, 'ms_per_step': ms / args.steps}}
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

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        e_steps = max(3, min(args.steps, 10))

        def e2e_run(an, bn):
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
            return dt
        ah, bh = a.cpu().pin_memory(), b.cpu().pin_memory()
        dt = e2e_run(ah.numpy(), bh.numpy())
        ceiling = h2d_ceiling(dev, world, dist)
        nbytes = int(ah.numel() * 4 + bh.numel() * 4)
        e2e = {'value': world * batch * e_steps / dt, 'unit': 'matches/s',
               'h2d_bytes_per_step': nbytes, 'd2h_bytes_per_step': int(5 * 8 * batch),
               'steps': e_steps, 'api': 'feabas_b200.cuda.xcorr_fft(numpy pinned host arrays) -> fb_xcorr_batch_host',
               'h2d_gbs': world * nbytes * e_steps / dt / 1e9, 'h2d_ceiling_gbs': ceiling,
               'frac_of_ceiling': world * nbytes * e_steps / dt / 1e9 / ceiling,
               'limiter': 'host -> device copies (PCIe / host memory): h2d_ceiling_gbs is what all ranks together get from pinned '
                          'cudaMemcpyAsync on this box, measured in this run'}
        del ah, bh
        # what a FEABAS caller passes: ordinary (pageable) numpy arrays, staged through the library's pinned slots
        ap_, bp_ = a.cpu().numpy().copy(), b.cpu().numpy().copy()
        dtp = e2e_run(ap_, bp_)
        e2e['pageable'] = {'value': world * batch * e_steps / dtp, 'unit': 'matches/s', 'h2d_gbs': world * nbytes * e_steps / dtp / 1e9,
                           'api': 'feabas_b200.cuda.xcorr_fft(pageable numpy arrays): memcpy into pinned staging slots, then H2D'}
        del ap_, bp_

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
        del a, b
        arm = CpuArm(wl, per_worker, seed=100)
        p_, s_ = arm.step()
        arm.close()
        cpu = {'value': p_ / s_, 'unit': 'matches/s', 'cores': arm.cores, 'kind': 'port',
               'sample': f'{p_} pairs ({per_worker} per single-thread worker, {arm.cores} workers) drawn from the first {arm.sample_pairs} pairs of '
                         f"this run's own batch (same generator and seed), oracle port of xcorr_fft (scipy pocketfft), {s_:.1f} s wall"}

    line = {'metric': 'xcorr_block_matches_per_sec', 'value': value, 'unit': 'matches/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu,
            'run_info': run_info, 'regimes': regimes}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
