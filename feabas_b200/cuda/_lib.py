"""ctypes binding of libfeabas_cuda.so (C ABI: include/feabas_cuda.h).

The library is built in-tree by ``python -m feabas_b200.csrc.build`` (or
``__graft_entry__.build()``).  There is no CPU fallback: if the shared library
is missing, or no CUDA device is present when a compute entry point is called,
the call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FEABAS_CUDA_LIB: another build of the same library (kernel-variant comparison runs)
LIB_PATH = os.environ.get('FEABAS_CUDA_LIB') or os.path.join(os.path.dirname(_HERE), 'csrc', 'libfeabas_cuda.so')

FB_F32, FB_U8, FB_F64 = 0, 1, 2
FB_FLAG_PAD = 0x1
FB_FLAG_SUBPIXEL = 0x2
FB_CONF_SHIFT = 2
FB_FLAG_FORCE_STAGED = 0x10
FB_FLAG_FORCE_FUSED = 0x20
FB_FLAG_U8_AS_F32 = 0x40
FB_FLAG_FORCE_GENERIC = 0x80
FB_FLAG_FORCE_FUSED_SMEM = 0x100
FB_DOG_UNSIGNED = 0x1
FB_DOG_EXACT = 0x2

_vp, _i, _ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
_XCORR_ARGS = [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]

class XcorrExt(ctypes.Structure):
    """fb_xcorr_ext of include/feabas_cuda.h."""
    _fields_ = [('nchan', _i), ('norm', _vp), ('norm_mirror', _vp), ('surface', _vp), ('surface_mirror', _vp)]


# every symbol include/feabas_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'fb_xcorr_batch_device': (_i, _XCORR_ARGS),
    'fb_xcorr_batch_host': (_i, _XCORR_ARGS),
    'fb_xcorr_batch': (_i, _XCORR_ARGS),
    'fb_xcorr_batch_device_ex': (_i, _XCORR_ARGS + [ctypes.POINTER(XcorrExt)]),
    'fb_masked_dog_workspace': (_ll, [_i, _i, _i]),
    'fb_masked_dog': (_i, [_vp, _vp, _i, _i, _i, _i, _i, ctypes.c_double, ctypes.c_double, _i, _vp, _vp, _ll, _i, _vp]),
    'fb_masked_dog_sparse': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_double, ctypes.c_double, _i, _vp, _vp, _ll, _i, _vp]),
    'fb_masked_dog_f64_workspace': (_ll, [_i, _i, _i]),
    'fb_masked_dog_f64': (_i, [_vp, _vp, _i, _i, _i, _i, ctypes.c_double, ctypes.c_double, _i, _vp, _vp, _ll, _i, _vp]),
    'fb_stack_minmax': (_i, [_vp, _i, _ll, _i, _vp, _i, _vp]),
    'fb_resize_area': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    'fb_resize_area_frac': (_i, [_vp, _i, _i, _i, _i, ctypes.c_double, ctypes.c_double, _vp, _i, _i, _i, _vp]),
    'fb_resize_nearest': (_i, [_vp, _i, _i, _i, ctypes.c_double, ctypes.c_double, _vp, _i, _i, _i, _vp]),
    'fb_crop_blocks': (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, ctypes.c_double, ctypes.c_double, ctypes.c_double, _vp, _vp, _vp, _vp, _i, _vp]),
    'fb_crop_blocks_multi': (_i, [_vp, _i, _vp, _i, _i, _i, ctypes.c_double, _vp, _i, _vp]),
    'fb_next_fast_len': (_i, [_i]),
    'fb_xcorr_plan_info': (_i, [_i, _i, _i, _i, _i, _i, _i, _i, ctypes.POINTER(_ll)]),
    'fb_set_option': (_i, [ctypes.c_char_p, _ll]),
    'fb_profile_read': (_i, [_i, _vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_ll), _i]),
    'fb_launch_count': (_ll, []),
    'fb_pair_count': (_ll, []),
    'fb_release': (_i, [_i]),
    'fb_device_count': (_i, []),
    'fb_last_error': (ctypes.c_char_p, []),
    'fb_version': (ctypes.c_char_p, []),
}

_lib = None


class FeabasCudaError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the .so is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FeabasCudaError(
                f'{LIB_PATH} not found: build it with `python -m feabas_b200.csrc.build` '
                '(feabas_b200 has no CPU fallback)')
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        # FEABAS_CUDA_OPTIONS="name=value,...": fb_set_option switches for a whole run (experiments, comparison runs)
        for item in filter(None, os.environ.get('FEABAS_CUDA_OPTIONS', '').split(',')):
            name, _, value = item.partition('=')
            rc = handle.fb_set_option(name.strip().encode(), int(value))
            if rc != 0:
                raise FeabasCudaError(f'FEABAS_CUDA_OPTIONS: {item!r}: {handle.fb_last_error().decode()}')
    return _lib


def check(rc):
    if rc != 0:
        raise FeabasCudaError(f'libfeabas_cuda error {rc}: {lib().fb_last_error().decode()}')


def plan_info(h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags):
    buf = (_ll * 8)()
    check(lib().fb_xcorr_plan_info(h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags, buf))
    keys = ('path', 'ws_bytes_per_pair', 'smem_fused', 'smem_row', 'smem_col', 'row_tile', 'col_tile', 'launches_per_chunk')
    out = dict(zip(keys, list(buf)))
    out['path'] = {1: 'fused', 2: 'staged', 3: 'staged-fast', 4: 'fused-warp'}[out['path']]
    return out


def launch_count():
    return int(lib().fb_launch_count())


def set_option(name, value):
    check(lib().fb_set_option(name.encode(), int(value)))


KERNEL_SLOTS = ('rows_forward', 'columns', 'rows_inverse', 'finalize', 'fused')


def profile_read(device, stream, reset=True):
    """{kernel: (total_ms, launches)} measured with CUDA events while option 'profile' is on; ``stream='all'``: summed
    over every stream context of the device (FB_ALL_STREAMS)."""
    if stream == 'all':
        stream = ctypes.c_void_p(-1)
    ms = (ctypes.c_double * 5)()
    cnt = (_ll * 5)()
    check(lib().fb_profile_read(int(device), stream, ms, cnt, int(reset)))
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(KERNEL_SLOTS)}
