/* feabas_cuda.h -- C ABI of libfeabas_cuda.so (B200 / sm_100a).
 *
 * The reference (YuelongWu/feabas, pure Python) has no FFI for this path; its
 * boundary is the Python function feabas.matcher.xcorr_fft and its callers,
 * imported by name at feabas/stitcher.py:20, feabas/aligner.py:19 and
 * feabas/thumbnail.py:17.  The entry points below are what a `feabas/cuda/`
 * package binds with ctypes to replace the arithmetic of those functions
 * (binding stub: INTEGRATION.md; Python mirror: feabas_b200/cuda/).
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, no exceptions.
 * Every function returns 0 on success or a negative FB_E* code; the message
 * for the last failure on the calling thread is fb_last_error().  The library
 * owns only cached per-device FFT tables and workspaces (fb_release frees them).
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef FEABAS_CUDA_H
#define FEABAS_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define FB_OK 0
#define FB_EINVAL (-1)     /* bad argument (shape, dtype, null pointer)              */
#define FB_ESIZE (-2)      /* FFT size not 2^a 3^b 5^c or too large for the kernels  */
#define FB_ECUDA (-3)      /* CUDA runtime error, see fb_last_error()                */
#define FB_ENOMEM (-4)     /* workspace allocation failed                            */

/* in_dtype: element type of the image stacks.  Compute type follows the
 * reference's scipy.fft promotion (feabas/matcher.py:63-64): float32 stays
 * float32 (complex64 spectra), uint8 and float64 run in float64.           */
#define FB_F32 0
#define FB_U8 1
#define FB_F64 2

/* flags */
#define FB_FLAG_PAD 0x1          /* informational: fft_h/fft_w were computed with pad=True  */
#define FB_FLAG_SUBPIXEL 0x2     /* 3x3 quadratic refinement, feabas/matcher.py:84-106      */
#define FB_CONF_SHIFT 2          /* bits 2-3: conf_mode, values of feabas/constant.py:39-41 */
#define FB_CONF_NONE 0
#define FB_CONF_STD 1
#define FB_CONF_MIRROR 2
#define FB_FLAG_FORCE_STAGED 0x10 /* testing: always use the HBM-staged 4-kernel pipeline    */
#define FB_FLAG_FORCE_FUSED 0x20  /* testing: always use the single-CTA fused kernel         */
#define FB_FLAG_U8_AS_F32 0x40    /* opt-in: compute uint8 input in float32 (not reference-exact) */
#define FB_FLAG_FORCE_GENERIC 0x80 /* testing: staged pipeline with the generic mixed-radix kernels */
#define FB_FLAG_FORCE_FUSED_SMEM 0x100 /* testing: the single-CTA kernel with shared-memory radix passes, never the
                                          warp-per-line one */

/* xcorr_fft (feabas/matcher.py:22-135) on a stack of n image pairs, sigma == 0,
 * single channel, no mask normalisation (those: fb_xcorr_batch_device_ex below).
 *
 *   img0 : n x h0 x w0, img1 : n x h1 x w1, row-major, dtype in_dtype
 *   fft_h, fft_w : the reference's fftshp (matcher.py:59-62), 5-smooth
 *   dx, dy, conf : n doubles each.  conf holds the float32-rounded value for
 *                  float32 compute, as the reference returns float32.
 *   peak, mirror : optional (may be NULL) n doubles: max of the correlation
 *                  surface and max |mirror surface| (0 unless MIRROR).
 *
 * fb_xcorr_batch_device: every pointer is device memory on `device`; work is
 *   enqueued on `stream` (NULL = legacy default stream) and the call returns
 *   without synchronising.
 * fb_xcorr_batch_host: every pointer is host memory (pinned or pageable);
 *   copies are pipelined against compute and the call returns when the
 *   outputs are valid.
 * fb_xcorr_batch: dispatches on the memory kind of img0.
 */
int fb_xcorr_batch_device(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                          int in_dtype, int fft_h, int fft_w, int flags,
                          double* dx, double* dy, double* conf, double* peak, double* mirror,
                          int device, void* stream);
int fb_xcorr_batch_host(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                        int in_dtype, int fft_h, int fft_w, int flags,
                        double* dx, double* dy, double* conf, double* peak, double* mirror,
                        int device, void* stream);
int fb_xcorr_batch(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                   int in_dtype, int fft_h, int fft_w, int flags,
                   double* dx, double* dy, double* conf, double* peak, double* mirror,
                   int device, void* stream);

/* The rarely used arguments of xcorr_fft, device pointers on `device` (NULL / 0 = off).  Any of them
 * selects the generic HBM-staged kernels.
 *   nchan          : channels per image; the stacks are n x nchan x h x w (the reference moves the channel
 *                    axis in front of H x W, matcher.py:50-53) and the cross-power is averaged over
 *                    channels before the inverse transform (matcher.py:66-67,115-116)
 *   norm           : fft_h x fft_w divisors of the correlation surface, applied before the peak search --
 *                    normalize=True (matcher.py:71-81); the caller builds them from the masks with a
 *                    first call that requests `surface`
 *   norm_mirror    : the same for the mirror surface (matcher.py:119-124)
 *   surface        : out, n x fft_h x fft_w: the correlation surface irfft2(conj(F0) F1)
 *   surface_mirror : out, n x fft_h x fft_w: |irfft2(F0 F1)| (FB_CONF_MIRROR only)
 * Element type of the four arrays: float for FB_F32 input, double otherwise (the compute type).        */
typedef struct fb_xcorr_ext {
    int nchan;
    const void* norm;
    const void* norm_mirror;
    void* surface;
    void* surface_mirror;
} fb_xcorr_ext;

int fb_xcorr_batch_device_ex(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                             int in_dtype, int fft_h, int fft_w, int flags,
                             double* dx, double* dy, double* conf, double* peak, double* mirror,
                             int device, void* stream, const fb_xcorr_ext* ext);

/* ---- image operators either side of the matcher (device pointers on `device`, work enqueued on
 * `stream`, no synchronisation) ------------------------------------------------------------- */

#define FB_DOG_UNSIGNED 0x1   /* signed=False: absolute value of the band-pass                      */
#define FB_DOG_EXACT 0x2      /* float64 accumulation in scipy.ndimage.correlate1d's order: the
                                 reference's rounding; default is float32 accumulation           */

/* common.masked_dog_filter (feabas/common.py:353-377) on a stack of n images h x w
 * (in_dtype FB_F32 or FB_U8), two cascaded separable Gaussians (scipy.ndimage.gaussian_filter1d:
 * radius int(4 sigma + 0.5), mode='nearest'), out = G(img) - G(G(img)), float32.
 * mask: NULL, or mask_n (1 or n) images of h x w bytes, nonzero = keep; when given, the term
 * 2 G_{sqrt2 sigma}(ptp * (mask == 0)) is subtracted from |out| (sign kept).  The caller decides
 * "mask is all true -> pass NULL" as the reference does (common.py:368).
 * ptp: np.ptp of the WHOLE array the reference would have been handed (common.py:369); NaN = take
 * it from this stack on the device.
 * work: scratch of fb_masked_dog_workspace(n, h, w) bytes.                                       */
long long fb_masked_dog_workspace(int n, int h, int w);
int fb_masked_dog(const void* img, const unsigned char* mask, int n, int h, int w, int in_dtype, int mask_n,
                  double sigma, double ptp, int flags, float* out, void* work, long long work_bytes,
                  int device, void* stream);
/* The same with masks for SOME images only: mask holds mask_n images, mask_images[i] (device, mask_n ints) says
 * which image of the stack mask i belongs to; images not listed have no masked pixel (their mask term is zero, the
 * result is what the reference computes for them).  The block passes of the matcher use it: only blocks that hang
 * over the border of the mesh carry a mask (feabas/renderer.py:436-449), the band-pass still sees the whole batch
 * (ptp of the whole stack, feabas/common.py:369).                                                              */
int fb_masked_dog_sparse(const void* img, const unsigned char* mask, const int* mask_images, int n, int h, int w, int in_dtype,
                         int mask_n, double sigma, double ptp, int flags, float* out, void* work, long long work_bytes,
                         int device, void* stream);

/* The same for float64 images (feabas/common.py:363 converts only non-floating dtypes, a float64 stack stays
 * float64 through scipy.ndimage and comes out float64): every operation in double in scipy's order, out:
 * n x h x w doubles.  mask: NULL or mask_n (1 or n) images.  work: fb_masked_dog_f64_workspace(n, h, w) bytes. */
long long fb_masked_dog_f64_workspace(int n, int h, int w);
int fb_masked_dog_f64(const double* img, const unsigned char* mask, int n, int h, int w, int mask_n,
                      double sigma, double ptp, int flags, double* out, void* work, long long work_bytes,
                      int device, void* stream);

/* {min, max} (float32) of each of n images of `elems` elements: np.ptp of blocks
 * (feabas/matcher.py:196,205) and of stacks (feabas/common.py:369).  minmax: n x 2 floats.      */
int fb_stack_minmax(const void* stack, int n, long long elems, int in_dtype, float* minmax, int device, void* stream);

/* cv2.resize(src, None, fx=1/k, fy=1/k, interpolation=cv2.INTER_AREA) as called at
 * feabas/matcher.py:254-256,321-322: k x k cell means, OpenCV's rounding (uint8 bit-exact).
 * oh, ow: the output size OpenCV picks, round-half-even(h / k), round-half-even(w / k).           */
int fb_resize_area(const void* src, int n, int h, int w, int in_dtype, int k, void* dst, int oh, int ow,
                   int device, void* stream);

/* The same call for ANY shrinking factor (coarse_downsample / fine_downsample are free parameters of
 * stitching_matcher, feabas/matcher.py:233-234): OpenCV's general INTER_AREA path -- per-axis tables of source
 * pixels and float32 coverage weights, float32 accumulation along x then y, round half to even for uint8
 * (bit-exact for uint8 and float32).  inv_fx, inv_fy = 1 / fx, 1 / fy as doubles (>= 1); oh, ow as above.
 * OpenCV takes the k x k path above only when |1/f - round(1/f)| < DBL_EPSILON on both axes.           */
int fb_resize_area_frac(const void* src, int n, int h, int w, int in_dtype, double inv_fx, double inv_fy,
                        void* dst, int oh, int ow, int device, void* stream);

/* cv2.resize(..., interpolation=cv2.INTER_NEAREST) of uint8 masks (feabas/matcher.py:257-264):
 * dst(y, x) = src(min(floor(y * inv_fy), h - 1), min(floor(x * inv_fx), w - 1)).                */
int fb_resize_nearest(const unsigned char* src, int n, int h, int w, double inv_fy, double inv_fx,
                      unsigned char* dst, int oh, int ow, int device, void* stream);

/* Block extraction for affine block maps: MeshRenderer.crop_multiple (feabas/renderer.py:601-648)
 * -> crop_field_affine (:419-450) -> common.render_by_subregions (feabas/common.py:256-350) ->
 * cv2.remap(INTER_LINEAR, BORDER_CONSTANT).  Gathers n blocks of bh x bw pixels from ONE source
 * image (ih x iw, FB_F32 or FB_U8; output has the same dtype).
 * blocks: n x 10 doubles (device): x0, y0, step_x, step_y, A00, A10, t0, A01, A11, t1; output
 * pixel (row, col) samples the source at
 *     xx = x0 + col * step_x,  yy = y0 + row * step_y            (np.linspace, endpoint=False)
 *     xs = xx * A00 + yy * A10 + t0,   ys = xx * A01 + yy * A11 + t1        (float64)
 * then, as OpenCV does, (xs - origin_x, ys - origin_y) is rounded to float32 and to 1/32 px,
 * and the four neighbours are blended with OpenCV's float (float32 images) or 15-bit integer
 * (uint8 images) weights; pixels outside the image read `fillval`.
 * origin_x, origin_y: integer-valued; the reference uses floor(min field) - 4 of the batch.
 * cover (HOST pointer, may be NULL): xmin, ymin, xmax, ymax of the region of the source covered by
 * the mesh (MeshRenderer's covered_region, feabas/renderer.py:98-101: the mesh shrunk by half a
 * pixel); pixels whose source position is not STRICTLY inside it (shapely.contains_xy,
 * renderer.py:447) are not rendered (they keep `fillval`) and, when mask_out != NULL
 * (n x bh x bw bytes, device), are flagged 0 there: the validity mask of
 * crop_field_affine(precise_mask=True) that masked_dog_filter consumes.
 * block_slot (device, n ints, may be NULL = block b writes mask image b; only read with cover):
 * < 0: the block counts as covered as a whole -- the reference skips the per-pixel test when less
 * than one square pixel of the block's footprint is uncovered (renderer.py:443-444), the caller
 * evaluates that rule -- and writes no mask; >= 0: the image of mask_out that receives the block's
 * mask (mask_out then only holds the partially covered blocks, see fb_masked_dog_sparse).        */
int fb_crop_blocks(const void* img, int ih, int iw, int in_dtype, const double* blocks, int n, int bh, int bw,
                   double origin_x, double origin_y, double fillval, void* out,
                   const double* cover, const int* block_slot, unsigned char* mask_out,
                   int device, void* stream);

/* The same gather for blocks that come from MANY source images in one launch: the block lists of all the
 * overlaps / section pairs that are at the same pyramid level (feabas/stitcher.py:385-394 fans them out to
 * worker processes; here they share one batch).  sources: device array of n records, one per block: the
 * block's source image (device pointer, dtype in_dtype), its size and the origin of the reference's source
 * crop for the batch the block belongs to.  No covered-region test (the sigma == 0 paths).              */
typedef struct fb_crop_src {
    const void* img;
    int ih, iw;
    double origin_x, origin_y;
} fb_crop_src;

int fb_crop_blocks_multi(const fb_crop_src* sources, int in_dtype, const double* blocks, int n, int bh, int bw,
                         double fillval, void* out, int device, void* stream);

/* Smallest 2^a 3^b 5^c >= target: scipy.fftpack.next_fast_len as used at
 * feabas/matcher.py:60,62.  Pure host arithmetic.                          */
int fb_next_fast_len(int target);

/* How a problem class will be executed.  info[0] = 1 fused (shared-memory passes) / 2 staged (generic
 * mixed-radix kernels) / 3 staged (register-resident kernels) / 4 fused (warp-per-line register transforms),
 * info[1] = bytes of HBM workspace per pair, info[2..4] = shared memory per
 * CTA of the fused / row / column kernels, info[5] = row tile lines,
 * info[6] = column tile width, info[7] = kernel launches per chunk.
 * Pure host arithmetic.                                                    */
int fb_xcorr_plan_info(int h0, int w0, int h1, int w1, int in_dtype, int fft_h, int fft_w, int flags,
                       long long* info8);

/* Tuning knobs: "ws_bytes" (ceiling of the HBM workspace per stream context, default 8 GiB: 650 pairs of 512^2
 * blocks at FFT 1024^2; allocated as needed, halved when the device cannot give it), "host_chunk_bytes" (input bytes per host-path chunk, default 64 MiB),
 * "copy_threads" (host threads that stage pageable input into the pinned slots, default 6),
 * "warp_fused" (0/1: small grids on the warp-per-line fused kernel, default 1),
 * "k2_solo" (0/1: experiment, column stage with one warp per column pair -- measured slower, default 0),
 * "profile" (0/1: time every kernel with CUDA events, see fb_profile_read).      */
int fb_set_option(const char* name, long long value);

/* Per-kernel device time, measured with CUDA events on the launching stream while
 * option "profile" is 1.  Slots: 0 rows-forward, 1 columns, 2 rows-inverse,
 * 3 finalize, 4 fused.  ms5 / launches5: arrays of 5 (accumulated since the
 * last reset).  Synchronises the pending events of that (device, stream);
 * stream == FB_ALL_STREAMS: the sum over every stream context of the device.  */
#define FB_ALL_STREAMS ((void*)(-1))
int fb_profile_read(int device, void* stream, double* ms5, long long* launches5, int reset);

/* Number of kernels this library has launched in this process.              */
long long fb_launch_count(void);
/* Block pairs handed to the xcorr kernels since the library was loaded (every fb_xcorr_* entry point): the unit
 * bench.py counts for workloads that call the matcher layers (stitching_matcher, ...). */
long long fb_pair_count(void);

/* Free cached tables / workspaces of `device` (-1: all).                     */
int fb_release(int device);

int fb_device_count(void);
const char* fb_last_error(void);
const char* fb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FEABAS_CUDA_H */
