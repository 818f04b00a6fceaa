"""feabas_b200.cuda -- the ``feabas/cuda`` package: B200-native drop-ins for the FFT
cross-correlation matcher of ``feabas/matcher.py`` over libfeabas_cuda.so (no CPU fallback)."""
from .constant import FFT_CONF_MIRROR, FFT_CONF_NONE, FFT_CONF_STD, Match
from .xcorr import fft_shape, next_fast_len, xcorr_fft, xcorr_fft_device
from .blocks import (bbox_centers, bbox_sizes, distributor_cartesian_bbox, divide_bbox, intersect_bbox, z_order)
from .image import masked_dog_filter, resize_area, resize_mask, crop_blocks
from .surrogate import AffineMesh, AffineSLM, ArrayLoader
from .matcher import (bboxes_mesh_renderer_matcher, bboxes_mesh_renderer_matcher_many, global_translation_matcher, iterative_xcorr_matcher_w_mesh,
                      section_matcher, section_matcher_many, set_mesh_factory, stitching_matcher, stitching_matcher_many)
from . import _lib, matchio

__all__ = ['xcorr_fft', 'xcorr_fft_device', 'fft_shape', 'next_fast_len',
           'global_translation_matcher', 'stitching_matcher', 'stitching_matcher_many', 'section_matcher', 'section_matcher_many',
           'iterative_xcorr_matcher_w_mesh',
           'bboxes_mesh_renderer_matcher', 'bboxes_mesh_renderer_matcher_many', 'distributor_cartesian_bbox', 'divide_bbox', 'intersect_bbox', 'z_order',
           'bbox_centers', 'bbox_sizes', 'masked_dog_filter', 'resize_area', 'resize_mask', 'crop_blocks',
           'AffineMesh', 'AffineSLM', 'ArrayLoader', 'set_mesh_factory', 'install', 'Match',
           'FFT_CONF_NONE', 'FFT_CONF_STD', 'FFT_CONF_MIRROR']


def install(matcher_module=None, common_module=None):
    """Rebind the arithmetic the reference's own control flow calls through module globals
    (feabas/matcher.py:153,213,846; feabas/common.py:353) to the CUDA implementations, so that
    ``feabas.matcher.stitching_matcher`` / ``section_matcher`` run unchanged on top of them.
    Call it at import time in every process (FEABAS spawns workers, feabas/concurrent.py:70)."""
    if matcher_module is None:
        import feabas.matcher as matcher_module          # pragma: no cover - needs FEABAS
    if common_module is None:
        import feabas.common as common_module            # pragma: no cover
    matcher_module.xcorr_fft = xcorr_fft
    matcher_module.global_translation_matcher = global_translation_matcher
    common_module.masked_dog_filter = masked_dog_filter
    return matcher_module
