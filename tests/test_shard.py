"""Row (e): sharding by index range + host-side gather, checked on CPU with the gloo backend
(world_size 2) and with threads.  The compute function is the oracle here -- the point is the
partition / gather logic, which must reproduce the unsharded result exactly and in index order."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from feabas_b200.cuda import shard


def test_shard_ranges_partition():
    for n in (0, 1, 2, 7, 8, 9, 100, 1023):
        for world in (1, 2, 3, 4, 8):
            rs = shard.shard_ranges(n, world)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _oracle_compute(a, b, conf_mode=2, **kw):
    from oracle import xcorr_oracle as xo
    kw.pop('device', None)
    return xo.xcorr_oracle(a, b, conf_mode=conf_mode, **kw)


def _pairs(n, seed=3):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, 24, 20)).astype(np.float32)
    b = np.stack([np.roll(a[i], (i % 5 - 2, 3 - i % 7), axis=(0, 1)) for i in range(n)]) \
        + 0.05 * rng.standard_normal((n, 24, 20)).astype(np.float32)
    return a, b


def test_multi_gpu_threads_match_unsharded():
    a, b = _pairs(11)
    ref = _oracle_compute(a, b, subpixel=True)
    for devices in ([0], [0, 1], [0, 1, 2, 3], list(range(16))):       # more "devices" than pairs: empty shards
        got = shard.xcorr_fft_multi_gpu(a, b, devices=devices, compute=_oracle_compute, subpixel=True)
        for g, r in zip(got, ref):
            assert np.array_equal(g, r)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        a, b = _pairs(n)
        dx, dy, conf = shard.xcorr_fft_sharded(a, b, compute=_oracle_compute, subpixel=True)
        # one pyramid level: a stand-in block matcher that returns one point pair per block
        bb0 = np.stack([np.arange(n) * 10, np.zeros(n), np.arange(n) * 10 + 8, np.full(n, 8)], axis=1)

        def matcher(m0, m1, l0, l1, b0, b1, **kw):
            return b0[:, :2].astype(float), b1[:, 2:].astype(float), b0[:, 0].astype(np.float32)

        xy0, xy1, w = shard.bboxes_matcher_sharded(matcher, None, None, None, None, bb0, bb0)
        np.savez(os.path.join(out_dir, f'r{rank}.npz'), dx=dx, dy=dy, conf=conf, xy0=xy0, xy1=xy1, w=w)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [9, 1])
def test_gloo_world2_matches_unsharded(tmp_path, n):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    a, b = _pairs(n)
    ref = _oracle_compute(a, b, subpixel=True)
    for rank in range(world):
        z = np.load(tmp_path / f'r{rank}.npz')
        assert np.array_equal(z['dx'], ref[0]) and np.array_equal(z['dy'], ref[1]) and np.array_equal(z['conf'], ref[2])
        assert np.array_equal(z['xy0'][:, 0], np.arange(n) * 10.0)          # index order preserved
        assert np.array_equal(z['w'], (np.arange(n) * 10).astype(np.float32))
        assert z['xy1'].shape == (n, 2)
