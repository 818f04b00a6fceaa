"""Sharding of the matcher's independent units over the GPUs of one box.

Block pairs of one ``xcorr_fft`` batch, equal-size block batches of one pyramid level
(feabas/matcher.py:804-822), overlaps of a section (feabas/stitcher.py:385-394) and section pairs are
independent: every rank takes ONE contiguous range of the (z-ordered) index, computes it on its own
GPU, and the host concatenates the match lists in index order.  There is no data-path collective --
the only communication is the host-side gather of ``(xy0, xy1, conf)`` (what matcher.py:657-666 does
for its worker processes).

Two launch models:

* one process per GPU under ``torch.distributed`` (``torchrun``): ``shard_range`` + ``gather_concat``;
* one process, one host thread per GPU: ``xcorr_fft_multi_gpu`` (the C ABI only serialises callers of the SAME
  (device, stream) context -- fb_xcorr.cu ``StreamCtx::mu`` -- so each thread drives its own device concurrently;
  ctypes releases the GIL for the duration of the call).
"""
import threading

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:                       # pragma: no cover
    torch = None
    dist = None


def shard_range(n, world, rank):
    """Contiguous index range ``[lo, hi)`` of rank ``rank`` of ``world``: ``range(r*n//w, (r+1)*n//w)``.
    Contiguous ranges of the z-ordered block list keep the source-image locality of each shard."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank {rank} of {world}')
    return (rank * n) // world, ((rank + 1) * n) // world


def shard_ranges(n, world):
    return [shard_range(n, world, r) for r in range(world)]


def _world(group=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_concat(parts, group=None):
    """Host-side gather: every rank passes its tuple of numpy arrays (its shard of each output, first axis
    = unit index) and receives the tuple of arrays concatenated over ranks in rank (= index) order.
    ``None`` entries (a rank with nothing to report) are skipped.  Works with any backend
    (``all_gather_object`` pickles through the host; the match lists are tiny next to the images)."""
    world, _ = _world(group)
    if world == 1:
        return tuple(parts)
    gathered = [None] * world
    dist.all_gather_object(gathered, tuple(parts), group=group)
    out = []
    for i in range(len(parts)):
        chunks = [g[i] for g in gathered if g is not None and g[i] is not None]
        out.append(np.concatenate(chunks, axis=0) if chunks else None)
    return tuple(out)


def _stack_ptp(stack):
    """np.ptp of a whole stack (numpy array or tensor) as a Python float."""
    if torch is not None and isinstance(stack, torch.Tensor):
        return float(stack.max() - stack.min()) if stack.dtype.is_floating_point else float(int(stack.max()) - int(stack.min()))
    stack = np.asarray(stack)
    return float(np.ptp(stack if stack.dtype.kind == 'f' else stack.astype(np.float64)))


def _global_ptp(img0, img1, kwargs):
    """The mask term of the band-pass uses np.ptp of the WHOLE stack (feabas/common.py:369): take it before the
    stack is cut into shards, so that the result does not depend on the number of GPUs."""
    if kwargs.get('sigma', 0) > 0 and 'ptp' not in kwargs and (kwargs.get('mask0') is not None or kwargs.get('mask1') is not None):
        kwargs = dict(kwargs, ptp=(_stack_ptp(img0), _stack_ptp(img1)))
    return kwargs


def xcorr_fft_sharded(img0, img1, conf_mode=2, group=None, compute=None, **kwargs):
    """``xcorr_fft`` over a batch that every rank holds (or can index) in full: rank r computes pairs
    ``shard_range(N, world, r)`` on its own GPU and all ranks return the full ``(dx, dy, conf)``.
    ``compute`` defaults to ``feabas_b200.cuda.xcorr_fft`` (the CPU tests inject the oracle to check
    the partition / gather logic without a GPU)."""
    if compute is None:
        from .xcorr import xcorr_fft as compute
    world, rank = _world(group)
    n = len(img0)
    lo, hi = shard_range(n, world, rank)
    kwargs = _global_ptp(img0, img1, kwargs)
    if hi > lo:
        dx, dy, conf = compute(img0[lo:hi], img1[lo:hi], conf_mode=conf_mode, **kwargs)
    else:
        dx, dy, conf = np.empty(0), np.empty(0), np.empty(0, dtype=np.float32)
    return gather_concat((np.asarray(dx), np.asarray(dy), np.asarray(conf)), group=group)


def bboxes_matcher_sharded(matcher, mesh0, mesh1, loader0, loader1, bboxes0, bboxes1, group=None, **kwargs):
    """One pyramid level (feabas/matcher.py:781-861) with the z-ordered block list split into one contiguous
    range per rank.  ``matcher`` is ``bboxes_mesh_renderer_matcher`` (ours or the reference's); every rank
    returns the concatenated ``(xy0, xy1, conf)`` of all ranks, in block order."""
    world, rank = _world(group)
    bboxes0, bboxes1 = np.asarray(bboxes0), np.asarray(bboxes1)
    lo, hi = shard_range(len(bboxes0), world, rank)
    if hi > lo:
        xy0, xy1, conf = matcher(mesh0, mesh1, loader0, loader1, bboxes0[lo:hi], bboxes1[lo:hi], **kwargs)
    else:
        xy0, xy1, conf = np.empty((0, 2)), np.empty((0, 2)), np.empty(0)
    return gather_concat((np.asarray(xy0), np.asarray(xy1), np.asarray(conf)), group=group)


def xcorr_fft_multi_gpu(img0, img1, conf_mode=2, devices=None, compute=None, **kwargs):
    """Single-process variant: one host thread per GPU, each running ``xcorr_fft(..., device=d)`` on its
    contiguous share of the host batch; results are written into one output in index order."""
    if compute is None:
        from .xcorr import xcorr_fft as compute
    if devices is None:
        from . import _lib
        devices = list(range(max(1, _lib.lib().fb_device_count())))
    n = len(img0)
    ranges = shard_ranges(n, len(devices))
    results, errors = [None] * len(devices), []
    kwargs = _global_ptp(img0, img1, kwargs)

    def work(i):
        lo, hi = ranges[i]
        if hi <= lo:
            return
        try:
            results[i] = compute(img0[lo:hi], img1[lo:hi], conf_mode=conf_mode, device=devices[i], **kwargs)
        except Exception as exc:            # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    parts = [r for r in results if r is not None]
    if not parts:
        return np.empty(0), np.empty(0), np.empty(0, dtype=np.float32)
    return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(3))
