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

/* xcorr_fft (feabas/matcher.py:22-135) on a stack of n image pairs, sigma == 0,
 * single channel, no mask normalisation.
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

/* Smallest 2^a 3^b 5^c >= target: scipy.fftpack.next_fast_len as used at
 * feabas/matcher.py:60,62.  Pure host arithmetic.                          */
int fb_next_fast_len(int target);

/* How a problem class will be executed.  info[0] = 1 fused / 2 staged (generic
 * mixed-radix kernels) / 3 staged (register-resident power-of-two kernels),
 * info[1] = bytes of HBM workspace per pair, info[2..4] = shared memory per
 * CTA of the fused / row / column kernels, info[5] = row tile lines,
 * info[6] = column tile width, info[7] = kernel launches per chunk.
 * Pure host arithmetic.                                                    */
int fb_xcorr_plan_info(int h0, int w0, int h1, int w1, int in_dtype, int fft_h, int fft_w, int flags,
                       long long* info8);

/* Tuning knobs: "ws_bytes" (HBM workspace budget per stream context, default
 * 2 GiB), "host_chunk_bytes" (input bytes per host-path chunk, default 64 MiB),
 * "profile" (0/1: time every kernel with CUDA events, see fb_profile_read).      */
int fb_set_option(const char* name, long long value);

/* Per-kernel device time, measured with CUDA events on the launching stream while
 * option "profile" is 1.  Slots: 0 rows-forward, 1 columns, 2 rows-inverse,
 * 3 finalize, 4 fused.  ms5 / launches5: arrays of 5 (accumulated since the
 * last reset).  Synchronises the pending events of that (device, stream).     */
int fb_profile_read(int device, void* stream, double* ms5, long long* launches5, int reset);

/* Number of kernels this library has launched in this process.              */
long long fb_launch_count(void);

/* Free cached tables / workspaces of `device` (-1: all).                     */
int fb_release(int device);

int fb_device_count(void);
const char* fb_last_error(void);
const char* fb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FEABAS_CUDA_H */
