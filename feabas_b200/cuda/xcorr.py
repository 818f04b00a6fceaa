"""``xcorr_fft`` with the reference's signature and return conventions
(feabas/matcher.py:22-135), computed by libfeabas_cuda.so on a B200.

Host side only does what the reference does around its FFT calls: channel
axis handling, the ``fftshp`` rule with the 5-smooth ``next_fast_len``
(matcher.py:59-62) and dtype conventions of the returned arrays.
"""
import ctypes

import numpy as np

from . import _lib
from .constant import FFT_CONF_MIRROR, FFT_CONF_NONE, FFT_CONF_STD

try:                                    # torch is plumbing: device memory + streams
    import torch
except Exception:                       # pragma: no cover
    torch = None


def next_fast_len(target):
    """scipy.fftpack.next_fast_len (5-smooth), as used at matcher.py:60,62."""
    return _lib.lib().fb_next_fast_len(int(target))


def fft_shape(shape0, shape1, pad=True):
    """matcher.py:59-62."""
    if pad:
        return tuple(next_fast_len(a + b - 1) for a, b in zip(shape0, shape1))
    return tuple(next_fast_len(max(a, b)) for a, b in zip(shape0, shape1))


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


_NP_CODES = {np.dtype(np.float32): _lib.FB_F32, np.dtype(np.uint8): _lib.FB_U8, np.dtype(np.float64): _lib.FB_F64}


def _canon_numpy(img):
    img = np.asarray(img)
    if img.dtype not in _NP_CODES:
        # scipy.fft promotes: float16 -> float32, every other int/bool -> float64
        img = img.astype(np.float32 if img.dtype == np.float16 else np.float64)
    return np.ascontiguousarray(img)


def _canon_torch(img):
    if img.dtype not in (torch.float32, torch.uint8, torch.float64):
        img = img.to(torch.float32 if img.dtype in (torch.float16, torch.bfloat16) else torch.float64)
    return img.contiguous()


def _flags(conf_mode, subpixel, pad, force=None, u8_as_f32=False):
    f = (int(conf_mode) & 3) << _lib.FB_CONF_SHIFT
    if subpixel:
        f |= _lib.FB_FLAG_SUBPIXEL
    if pad:
        f |= _lib.FB_FLAG_PAD
    if force == 'staged':
        f |= _lib.FB_FLAG_FORCE_STAGED
    elif force == 'fused':
        f |= _lib.FB_FLAG_FORCE_FUSED
    elif force == 'fused_smem':
        f |= _lib.FB_FLAG_FORCE_FUSED | _lib.FB_FLAG_FORCE_FUSED_SMEM
    elif force == 'generic':
        f |= _lib.FB_FLAG_FORCE_STAGED | _lib.FB_FLAG_FORCE_GENERIC
    if u8_as_f32:
        f |= _lib.FB_FLAG_U8_AS_F32
    return f


def xcorr_fft_device(img0, img1, conf_mode=FFT_CONF_MIRROR, subpixel=False, pad=True, fftshp=None,
                     force=None, u8_as_f32=False, out=None, want_debug=False):
    """Device-resident variant: ``img0``/``img1`` are CUDA tensors ``N x H x W``; returns a
    float64 CUDA tensor ``out`` of shape ``(5, N)`` holding dx, dy, conf, peak, mirror-max.
    Work is enqueued on torch's current stream; nothing is synchronised."""
    if not (_is_torch(img0) and _is_torch(img1) and img0.is_cuda and img1.is_cuda):
        raise TypeError('xcorr_fft_device needs CUDA tensors')
    if img0.dim() != 3 or img1.dim() != 3 or img0.shape[0] != img1.shape[0]:
        raise ValueError('expected two N x H x W stacks with the same N')
    if img0.dtype != img1.dtype:
        common = torch.promote_types(img0.dtype, img1.dtype)
        img0, img1 = img0.to(common), img1.to(common)
    img0, img1 = _canon_torch(img0), _canon_torch(img1)
    code = {torch.float32: _lib.FB_F32, torch.uint8: _lib.FB_U8, torch.float64: _lib.FB_F64}[img0.dtype]
    n, h0, w0 = img0.shape
    _, h1, w1 = img1.shape
    ny, nx = fftshp if fftshp is not None else fft_shape((h0, w0), (h1, w1), pad)
    dev = img0.device.index if img0.device.index is not None else torch.cuda.current_device()
    if out is None:
        out = torch.empty((5, n), dtype=torch.float64, device=img0.device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        base, step = out.data_ptr(), n * 8
        _lib.check(_lib.lib().fb_xcorr_batch_device(
            img0.data_ptr(), img1.data_ptr(), n, h0, w0, h1, w1, code, ny, nx,
            _flags(conf_mode, subpixel, pad, force, u8_as_f32),
            base, base + step, base + 2 * step, base + 3 * step, base + 4 * step, dev, stream))
    return out


def _conf_dtype(conf_mode, in_dtype):
    # matcher.py:133: `** np.prod(fftshp)` (an int64 scalar) promotes the STD confidence to
    # float64 whatever the image dtype; NONE / MIRROR are float32 (matcher.py:112,126)
    return np.float64 if conf_mode == FFT_CONF_STD else np.float32


def _run_ext(t0, t1, nchan, conf_mode, subpixel, pad, fftshp, norm=None, norm_mirror=None, want_surfaces=False):
    """``fb_xcorr_batch_device_ex`` on CUDA tensors ``N x (C x) H x W`` (channel axis already in front)."""
    code = {torch.float32: _lib.FB_F32, torch.uint8: _lib.FB_U8, torch.float64: _lib.FB_F64}[t0.dtype]
    n, (h0, w0), (h1, w1) = t0.shape[0], t0.shape[-2:], t1.shape[-2:]
    ny, nx = fftshp
    ctype = torch.float32 if t0.dtype == torch.float32 else torch.float64
    out = torch.empty((5, n), dtype=torch.float64, device=t0.device)
    surf = surf_m = None
    ext = _lib.XcorrExt(nchan=int(nchan))
    if want_surfaces:
        surf = torch.empty((n, ny, nx), dtype=ctype, device=t0.device)
        ext.surface = surf.data_ptr()
        if conf_mode == FFT_CONF_MIRROR:
            surf_m = torch.empty((n, ny, nx), dtype=ctype, device=t0.device)
            ext.surface_mirror = surf_m.data_ptr()
    if norm is not None:
        ext.norm = norm.data_ptr()
    if norm_mirror is not None:
        ext.norm_mirror = norm_mirror.data_ptr()
    dev = t0.device.index
    with torch.cuda.device(dev):
        base, step = out.data_ptr(), n * 8
        _lib.check(_lib.lib().fb_xcorr_batch_device_ex(
            t0.data_ptr(), t1.data_ptr(), n, h0, w0, h1, w1, code, ny, nx, _flags(conf_mode, subpixel, pad),
            base, base + step, base + 2 * step, base + 3 * step, base + 4 * step, dev,
            torch.cuda.current_stream().cuda_stream, ctypes.byref(ext)))
    return out, surf, surf_m


def _xcorr_fft_ext(img0, img1, conf_mode, subpixel, pad, normalize, mask0, mask1, device):
    """Multi-channel stacks (matcher.py:50-53,66-67) and ``normalize=True`` (matcher.py:71-81,119-124):
    the generic staged kernels through ``fb_xcorr_batch_device_ex``."""
    from .image import to_device
    t0, t1 = to_device(img0, device), None
    t1 = to_device(img1, t0.device.index)
    if t0.dtype != t1.dtype:
        common = torch.promote_types(t0.dtype, t1.dtype)
        t0, t1 = t0.to(common), t1.to(common)
    t0, t1 = _canon_torch(t0), _canon_torch(t1)
    nchan = 1
    if t0.dim() > 3 or t1.dim() > 3:
        if t0.dim() != 4 or t1.dim() != 4 or t0.shape[-1] != t1.shape[-1]:
            raise ValueError('multi-channel stacks must both be N x H x W x C with the same C')
        nchan = t0.shape[-1]
        t0 = t0.movedim(-1, 1).contiguous()
        t1 = t1.movedim(-1, 1).contiguous()
    if t0.shape[0] != t1.shape[0]:
        raise ValueError('expected two stacks with the same N')
    fftshp = fft_shape(tuple(t0.shape[-2:]), tuple(t1.shape[-2:]), pad)
    norm = norm_m = None
    if normalize:
        ctype = torch.float32 if t0.dtype == torch.float32 else torch.float64
        # the masks default to all-ones of each image's H x W (matcher.py:72-75); their correlation
        # surfaces come from the same kernels (one pair, no confidence needed beyond the mirror surface)
        m0 = torch.ones(tuple(t0.shape[-2:]), dtype=ctype, device=t0.device) if mask0 is None else to_device(mask0, t0.device.index).to(ctype)
        m1 = torch.ones(tuple(t1.shape[-2:]), dtype=ctype, device=t0.device) if mask1 is None else to_device(mask1, t0.device.index).to(ctype)
        mode = FFT_CONF_MIRROR if conf_mode == FFT_CONF_MIRROR else FFT_CONF_NONE
        _, s, sm = _run_ext(m0[None].contiguous(), m1[None].contiguous(), 1, mode, False, pad, fftshp, want_surfaces=True)
        norm = (s[0] / s[0].max().clamp(min=1)).clamp(min=0.1).contiguous()
        if sm is not None:
            norm_m = (sm[0] / sm[0].max().clamp(min=1)).clamp(min=0.1).contiguous()
    out, _, _ = _run_ext(t0, t1, nchan, conf_mode, subpixel, pad, fftshp, norm=norm, norm_mirror=norm_m)
    in_dtype = np.float32 if t0.dtype == torch.float32 else np.float64
    return out.cpu().numpy(), in_dtype


def xcorr_fft(img0, img1, conf_mode=FFT_CONF_MIRROR, **kwargs):
    """Drop-in for ``feabas.matcher.xcorr_fft`` (matcher.py:22-135).

    Args and kwargs as the reference: ``sigma`` (0), ``mask0``/``mask1`` (None),
    ``normalize`` (False), ``subpixel`` (False), ``pad`` (True).  Extra,
    non-reference kwargs: ``device`` (GPU index for host input), ``return_debug``
    (also return a dict with the surface maxima), ``force`` ('fused'|'staged'),
    ``ptp`` (``sigma`` > 0 with masks: the ``np.ptp`` pair ``(ptp0, ptp1)`` of the
    WHOLE stacks when this call only sees a shard of them, common.py:369).
    Returns ``(dx, dy, conf)``: float64, float64, float32 numpy arrays of length N.
    """
    sigma = kwargs.get('sigma', 0)
    normalize = kwargs.get('normalize', False)
    subpixel = kwargs.get('subpixel', False)
    pad = kwargs.get('pad', True)
    device = kwargs.get('device', None)
    return_debug = kwargs.get('return_debug', False)
    force = kwargs.get('force', None)
    ndim = max(img0.dim() if _is_torch(img0) else np.ndim(img0), img1.dim() if _is_torch(img1) else np.ndim(img1))
    if sigma > 0:
        # matcher.py:54-56: masked DoG band-pass of both stacks first (device kernel, float32 out; the
        # channel axis moves in front of H x W before filtering, matcher.py:50-53)
        from .image import masked_dog_filter, to_device
        a, b = to_device(img0, device), None
        b = to_device(img1, a.device.index)
        if ndim > 3:
            a, b = a.movedim(-1, 1).contiguous(), b.movedim(-1, 1).contiguous()
        ptp = kwargs.get('ptp', None)
        ptp0, ptp1 = (None, None) if ptp is None else (tuple(ptp) if np.ndim(ptp) else (ptp, ptp))
        a = masked_dog_filter(a, sigma, mask=kwargs.get('mask0', None), ptp=ptp0)
        b = masked_dog_filter(b, sigma, mask=kwargs.get('mask1', None), ptp=ptp1)
        if ndim > 3:
            a, b = a.movedim(1, -1), b.movedim(1, -1)
        img0, img1 = a, b
    on_gpu = _is_torch(img0) and img0.is_cuda
    if ndim > 3 or normalize:
        out, in_dtype = _xcorr_fft_ext(img0, img1, conf_mode, subpixel, pad, normalize,
                                       kwargs.get('mask0', None), kwargs.get('mask1', None), device)
    elif on_gpu:
        in_dtype = {torch.float32: np.float32}.get(img0.dtype, np.float64)
        out = xcorr_fft_device(img0, img1, conf_mode=conf_mode, subpixel=subpixel, pad=pad, force=force).cpu().numpy()
    else:
        if _is_torch(img0):
            img0 = img0.numpy()
        if _is_torch(img1):
            img1 = img1.numpy()
        a, b = np.asarray(img0), np.asarray(img1)
        if a.dtype != b.dtype:
            common = np.promote_types(a.dtype, b.dtype)
            a, b = a.astype(common), b.astype(common)
        a, b = _canon_numpy(a), _canon_numpy(b)
        if a.ndim != 3 or b.ndim != 3 or a.shape[0] != b.shape[0]:
            raise ValueError('expected two N x H x W stacks with the same N')
        in_dtype = np.float32 if a.dtype == np.float32 else np.float64
        n, h0, w0 = a.shape
        _, h1, w1 = b.shape
        ny, nx = fft_shape((h0, w0), (h1, w1), pad)
        out = np.empty((5, n), dtype=np.float64)
        if device is None:
            device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
        ptr = out.ctypes.data
        _lib.check(_lib.lib().fb_xcorr_batch_host(
            a.ctypes.data, b.ctypes.data, n, h0, w0, h1, w1, _NP_CODES[a.dtype], ny, nx,
            _flags(conf_mode, subpixel, pad, force), ptr, ptr + 8 * n, ptr + 16 * n, ptr + 24 * n, ptr + 32 * n,
            int(device), None))
    dx, dy = out[0].copy(), out[1].copy()
    conf = out[2].astype(_conf_dtype(conf_mode, in_dtype))
    if return_debug:
        return dx, dy, conf, dict(peak=out[3].copy(), mirror_max=out[4].copy())
    return dx, dy, conf
