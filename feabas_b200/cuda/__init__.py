"""feabas_b200.cuda -- the ``feabas/cuda`` package: B200-native drop-ins for the FFT
cross-correlation matcher of ``feabas/matcher.py`` over libfeabas_cuda.so."""
from .constant import FFT_CONF_MIRROR, FFT_CONF_NONE, FFT_CONF_STD
from .xcorr import fft_shape, next_fast_len, xcorr_fft, xcorr_fft_device
from . import _lib

__all__ = ['xcorr_fft', 'xcorr_fft_device', 'fft_shape', 'next_fast_len',
           'FFT_CONF_NONE', 'FFT_CONF_STD', 'FFT_CONF_MIRROR']
