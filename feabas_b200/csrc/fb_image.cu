// fb_image.cu -- the image operators either side of the FFT matcher, on device pointers:
//
//   fb_masked_dog        common.masked_dog_filter            feabas/common.py:353-377
//   fb_resize_area       cv2.resize(INTER_AREA), factor 1/k  feabas/matcher.py:254-256
//   fb_resize_nearest    cv2.resize(INTER_NEAREST) of masks  feabas/matcher.py:257-264
//   fb_crop_blocks       MeshRenderer.crop_multiple for affine block maps
//                        feabas/renderer.py:419-450,601-648 -> feabas/common.py:256-350 (cv2.remap)
//   fb_stack_minmax      np.ptp per image / per stack         feabas/matcher.py:196,205; common.py:369
//
// All of them are HBM-bound pixel kernels; arithmetic follows the reference's rounding where the
// header says so (float64 accumulation of scipy.ndimage.correlate1d, OpenCV's fixed-point tables).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "../../include/feabas_cuda.h"
#include "fb_common.h"

namespace {

constexpr int kMaxRadius = 40;     // sigma <= 10 at truncate = 4
constexpr int kTH = 32, kTW = 64, kNT = 256;

struct Taps {
    int radius;
    double w[kMaxRadius + 1];      // w[j] = weight at offset +-(radius - j); w[radius] = centre
};

// scipy.ndimage._filters._gaussian_kernel1d (order 0), radius = int(truncate * sigma + 0.5), truncate = 4
bool make_taps(double sigma, Taps& t)
{
    int r = (int)(4.0 * sigma + 0.5);
    if (r < 0 || r > kMaxRadius || !(sigma > 0)) return false;
    std::vector<double> phi(2 * r + 1);
    const double s2 = sigma * sigma;
    double sum = 0;
    for (int x = -r; x <= r; ++x) { phi[x + r] = exp(-0.5 / s2 * (double)(x * x)); }
    for (double v : phi) sum += v;
    t.radius = r;
    for (int j = 0; j <= r; ++j) t.w[j] = phi[j] / sum;
    return true;
}

enum { MODE_BLUR = 0, MODE_DOG = 1, MODE_MASK = 2 };

template <typename TS> struct Vec4;
template <> struct Vec4<unsigned char> { typedef uchar4 type; };
template <> struct Vec4<float> { typedef float4 type; };


struct GaussParams {
    const void* src;       // BLUR: image (TS); DOG: first blur (float); MASK: mask bytes (nonzero = keep)
    float* dst;            // BLUR: blur; DOG: blur - blur(blur); MASK: in-place mask suppression of dst
    int n, h, w;
    long long src_stride;  // elements between consecutive images of src (0: one mask for all)
    const int* dst_index;  // MASK: image z of src belongs to image dst_index[z] of dst (null: z); n counts the images of src
    const float* span;     // MASK: device pointer to {min, max} of the image stack, or null
    float span_value;      //       used when span == null
    float sc2, s02;        // MASK: sigma_c^2, sigma^2 as float32 (common.py:371)
    int take_abs;          // unsigned output
    Taps taps;
};

template <typename ACC> struct Sym;
template <> struct Sym<double> {
    // scipy ni_filters.c, symmetric branch: tmp = centre * w; tmp += (x[-i] + x[i]) * w[i] from the outside in,
    // without contraction (the reference build has no FMA)
    static __device__ __forceinline__ double init(float c, double w) { return __dmul_rn((double)c, w); }
    static __device__ __forceinline__ double step(double acc, float a, float b, double w)
    {
        return __dadd_rn(acc, __dmul_rn(__dadd_rn((double)a, (double)b), w));
    }
};
template <> struct Sym<float> {
    static __device__ __forceinline__ float init(float c, double w) { return c * (float)w; }
    static __device__ __forceinline__ float step(float acc, float a, float b, double w) { return fmaf(a + b, (float)w, acc); }
};

template <typename TS, int MODE> __device__ __forceinline__ float load_src(const GaussParams& p, const TS* base, int y, int x, float span)
{
    y = min(max(y, 0), p.h - 1);                   // mode='nearest'
    x = min(max(x, 0), p.w - 1);
    const TS v = __ldg(base + (size_t)y * p.w + x);
    if (MODE == MODE_MASK) return v ? 0.f : span;  // np.ptp(img) * (mask == 0)
    return (float)v;
}

// One separable Gaussian (row pass, float32 rounding, column pass) of a kTH x kTW output tile.
template <typename TS, int MODE, typename ACC>
__global__ void __launch_bounds__(kNT) fbk_gauss2d(const __grid_constant__ GaussParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = p.taps.radius;
    const int in_w = kTW + 2 * r, in_h = kTH + 2 * r;
    const int in_pitch = in_w | 1;
    float* s_in = reinterpret_cast<float*>(smem_raw);             // [in_h][in_pitch]
    float* s_row = s_in + (size_t)in_h * in_pitch;                // [in_h][kTW]
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH, tid = threadIdx.x;
    const int img = (MODE == MODE_MASK && p.dst_index) ? p.dst_index[blockIdx.z] : (int)blockIdx.z;     // image of dst
    const TS* base = reinterpret_cast<const TS*>(p.src) + (size_t)blockIdx.z * p.src_stride;
    float span = 0.f;
    if (MODE == MODE_MASK) span = p.span ? (p.span[1] - p.span[0]) : p.span_value;
    for (int i = tid; i < in_h * in_w; i += kNT) {
        const int yy = i / in_w, xx = i - yy * in_w;
        s_in[yy * in_pitch + xx] = load_src<TS, MODE>(p, base, y0 + yy - r, x0 + xx - r, span);
    }
    __syncthreads();
    for (int i = tid; i < in_h * kTW; i += kNT) {
        const int yy = i / kTW, xx = i - yy * kTW;
        const float* c = s_in + yy * in_pitch + xx + r;
        ACC acc = Sym<ACC>::init(c[0], p.taps.w[r]);
        for (int j = 0; j < r; ++j) acc = Sym<ACC>::step(acc, c[j - r], c[r - j], p.taps.w[j]);
        s_row[yy * kTW + xx] = (float)acc;
    }
    __syncthreads();
    for (int i = tid; i < kTH * kTW; i += kNT) {
        const int yy = i / kTW, xx = i - yy * kTW;
        const int y = y0 + yy, x = x0 + xx;
        if (y >= p.h || x >= p.w) continue;
        const float* c = s_row + (yy + r) * kTW + xx;
        ACC acc = Sym<ACC>::init(c[0], p.taps.w[r]);
        for (int j = 0; j < r; ++j) acc = Sym<ACC>::step(acc, c[(j - r) * kTW], c[(r - j) * kTW], p.taps.w[j]);
        const float g = (float)acc;
        float* o = p.dst + ((size_t)img * p.h + y) * p.w + x;
        if (MODE == MODE_BLUR) {
            *o = g;
        } else if (MODE == MODE_DOG) {
            float f = __fsub_rn(s_in[(yy + r) * in_pitch + xx + r], g);     // img0f - img1f
            *o = p.take_abs ? fabsf(f) : f;
        } else {
            const float mf = __fdiv_rn(__fmul_rn(g, p.sc2), p.s02);
            const float f = *o;
            float a = fmaxf(__fsub_rn(fabsf(f), mf), 0.f);
            if (!p.take_abs) a = f > 0.f ? a : (f < 0.f ? -a : __fmul_rn(a, 0.f));
            *o = a;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// float32-accumulating variant (the default): same arithmetic, in the same order, as
// fbk_gauss2d<.., float>, but register blocked -- a thread produces 8 consecutive outputs along the
// filter direction from a sliding window (15 + 15 shared-memory loads per 8 taps x 8 outputs
// instead of 2 per tap and output), on 64 x 64 tiles.  The tap list is padded to a multiple of 8
// with zero weights: fmaf(x, 0, acc) == acc for finite x, and every padded read lands on
// initialised (pixel or zeroed slack) shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kFT = 64, kFP = 8, kSlack = 8, kRowPitch = kFT + 1;

struct GaussTapsF {
    float wc;                          // centre weight
    float wt[kMaxRadius + 8];          // wt[j] = weight at offset +-(radius - j), zero for j >= radius
    int rpad;                          // radius rounded up to a multiple of the chunk size
};

__host__ __device__ inline int gaussf_in_pitch(int r) { return (kFT + 2 * r + kFP) | 1; }

// CH: taps per chunk (5..8), chosen on the host so that the tap list needs the least zero padding
template <typename TS, int MODE, int CH>
__global__ void __launch_bounds__(kNT) fbk_gauss2d_f32(const __grid_constant__ GaussParams p, const __grid_constant__ GaussTapsF tf)
{
    constexpr int WN = CH + kFP - 1;                 // window: CH taps x 8 outputs touch CH + 7 consecutive values
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int r = p.taps.radius, rp = tf.rpad;
    const int in_w = kFT + 2 * r, in_h = kFT + 2 * r, in_pitch = gaussf_in_pitch(r);
    float* s_in_base = reinterpret_cast<float*>(smem_raw);                    // slack | [in_h][in_pitch] | slack
    float* s_in = s_in_base + kSlack;
    float* s_row_base = s_in + (size_t)in_h * in_pitch + kSlack;              // slack rows | [in_h][kRowPitch] | slack rows
    float* s_row = s_row_base + kSlack * kRowPitch;
    const int x0 = blockIdx.x * kFT, y0 = blockIdx.y * kFT, tid = threadIdx.x;
    const int img = (MODE == MODE_MASK && p.dst_index) ? p.dst_index[blockIdx.z] : (int)blockIdx.z;     // image of dst
    const TS* base = reinterpret_cast<const TS*>(p.src) + (size_t)blockIdx.z * p.src_stride;
    float span = 0.f;
    if (MODE == MODE_MASK) span = p.span ? (p.span[1] - p.span[0]) : p.span_value;
    // input tile (+ zeroed pitch padding and slack)
    {
        const float inv_pitch = 1.0f / (float)in_pitch;
        for (int i = tid; i < in_h * in_pitch; i += kNT) {
            const int yy = (int)(((float)i + 0.5f) * inv_pitch), xx = i - yy * in_pitch;
            s_in[i] = xx < in_w ? load_src<TS, MODE>(p, base, y0 + yy - r, x0 + xx - r, span) : 0.f;
        }
        for (int i = tid; i < kSlack; i += kNT) { s_in_base[i] = 0.f; s_in[(size_t)in_h * in_pitch + i] = 0.f; }
        for (int i = tid; i < kSlack * kRowPitch; i += kNT) { s_row_base[i] = 0.f; s_row[(size_t)in_h * kRowPitch + i] = 0.f; }
    }
    __syncthreads();
    // row pass: item = (row yy, strip of 8 outputs); consecutive threads take consecutive rows (odd pitch: no conflicts)
    {
        const float inv_h = 1.0f / (float)in_h;
        for (int i = tid; i < in_h * (kFT / kFP); i += kNT) {
            const int strip = (int)(((float)i + 0.5f) * inv_h), yy = i - strip * in_h;
            const float* row = s_in + yy * in_pitch + strip * kFP;
            float acc[kFP];
#pragma unroll
            for (int o = 0; o < kFP; ++o) acc[o] = row[r + o] * tf.wc;
            for (int jb = 0; jb < rp; jb += CH) {
                float a[WN], b[WN];
                const float* pa = row + jb;
                const float* pb = row + 2 * r - jb - (CH - 1);
#pragma unroll
                for (int k = 0; k < WN; ++k) { a[k] = pa[k]; b[k] = pb[k]; }
#pragma unroll
                for (int jj = 0; jj < CH; ++jj) {
                    const float wgt = tf.wt[jb + jj];
#pragma unroll
                    for (int o = 0; o < kFP; ++o) acc[o] = fmaf(a[jj + o] + b[CH - 1 - jj + o], wgt, acc[o]);
                }
            }
            float* dst = s_row + yy * kRowPitch + strip * kFP;
#pragma unroll
            for (int o = 0; o < kFP; ++o) dst[o] = acc[o];
        }
    }
    __syncthreads();
    // column pass: item = (column xx, strip of 8 output rows); consecutive threads take consecutive columns
    for (int i = tid; i < kFT * (kFT / kFP); i += kNT) {
        const int ys = i / kFT, xx = i - ys * kFT;
        const float* col = s_row + (ys * kFP) * kRowPitch + xx;
        float acc[kFP];
#pragma unroll
        for (int o = 0; o < kFP; ++o) acc[o] = col[(r + o) * kRowPitch] * tf.wc;
        for (int jb = 0; jb < rp; jb += CH) {
            float a[WN], b[WN];
            const float* pa = col + jb * kRowPitch;
            const float* pb = col + (2 * r - jb - (CH - 1)) * kRowPitch;
#pragma unroll
            for (int k = 0; k < WN; ++k) { a[k] = pa[k * kRowPitch]; b[k] = pb[k * kRowPitch]; }
#pragma unroll
            for (int jj = 0; jj < CH; ++jj) {
                const float wgt = tf.wt[jb + jj];
#pragma unroll
                for (int o = 0; o < kFP; ++o) acc[o] = fmaf(a[jj + o] + b[CH - 1 - jj + o], wgt, acc[o]);
            }
        }
        const int x = x0 + xx;
        if (x >= p.w) continue;
#pragma unroll
        for (int o = 0; o < kFP; ++o) {
            const int yy = ys * kFP + o, y = y0 + yy;
            if (y >= p.h) break;
            const float g = acc[o];
            float* out = p.dst + ((size_t)img * p.h + y) * p.w + x;
            if (MODE == MODE_BLUR) {
                *out = g;
            } else if (MODE == MODE_DOG) {
                float f = __fsub_rn(s_in[(yy + r) * in_pitch + xx + r], g);     // img0f - img1f
                *out = p.take_abs ? fabsf(f) : f;
            } else {
                const float mf = __fdiv_rn(__fmul_rn(g, p.sc2), p.s02);
                const float f = *out;
                float v = fmaxf(__fsub_rn(fabsf(f), mf), 0.f);
                if (!p.take_abs) v = f > 0.f ? v : (f < 0.f ? -v : __fmul_rn(v, 0.f));
                *out = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Register-window variant (the default for radius <= 24): same arithmetic, in the same order, as
// fbk_gauss2d_f32 -- acc = centre * wc, then fmaf(x[-j] + x[+j], w, acc) from the outermost tap inwards -- but
// a thread pulls the 8 + 2 R values its 8 outputs need into registers with 128-bit shared-memory loads and runs
// every tap from registers (compile-time radius bucket R, tap list padded with zero weights on the outside:
// fmaf(x, 0, acc) == acc).  The row pass leaves its result TRANSPOSED, so the column pass is the same code reading
// contiguous memory.  Instructions per output and pass: 2 R + 1 floating point, (8 + 2 R) / 32 loads, against ~95 in
// the kernel above (ncu: that one is issue bound with LSU and ALU as busy as the FMA pipe).
// ---------------------------------------------------------------------------------------------
template <int R> struct RegGauss {
    static constexpr int IN = kFT + 2 * R;                         // input tile side
    static constexpr int PITCH = ((IN + 31 - 4) / 32) * 32 + 4 >= IN ? ((IN + 31 - 4) / 32) * 32 + 4 : ((IN + 31 - 4) / 32) * 32 + 36;   // >= IN, = 4 (mod 32)
    static constexpr int WIN = kFP + 2 * R;                        // values a thread needs (a multiple of 4)
    static constexpr size_t SMEM = ((size_t)IN * PITCH + (size_t)kFT * PITCH) * sizeof(float);
};

struct RegTaps {
    float wc;
    float w[24];                        // w[j]: weight at offset +-(R - j), zero for the padded outer taps
};

// 8 outputs along a contiguous line: out[o] = centre + sum_j w[j] (x[o + j] + x[o + 2 R - j]), x = window of 8 + 2 R values
template <int R>
__device__ __forceinline__ void reg_line(const float* __restrict__ src, const RegTaps& t, float (&acc)[kFP])
{
    constexpr int WIN = RegGauss<R>::WIN;
    float x[WIN];
#pragma unroll
    for (int k = 0; k < WIN; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + k);
        x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
    }
#pragma unroll
    for (int o = 0; o < kFP; ++o) acc[o] = x[R + o] * t.wc;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const float wgt = t.w[j];
#pragma unroll
        for (int o = 0; o < kFP; ++o) acc[o] = fmaf(x[o + j] + x[o + 2 * R - j], wgt, acc[o]);
    }
}

template <typename TS, int MODE, int R>
__global__ void __launch_bounds__(kNT) fbk_gauss2d_reg(const __grid_constant__ GaussParams p, const __grid_constant__ RegTaps tf)
{
    using G = RegGauss<R>;
    constexpr int IN = G::IN, PITCH = G::PITCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_in = reinterpret_cast<float*>(smem_raw);              // [IN][PITCH]: input tile, halo R
    float* s_t = s_in + (size_t)IN * PITCH;                        // [kFT][PITCH]: row pass, transposed: s_t[x][y]
    const int x0 = blockIdx.x * kFT, y0 = blockIdx.y * kFT, tid = threadIdx.x;
    const int img = (MODE == MODE_MASK && p.dst_index) ? p.dst_index[blockIdx.z] : (int)blockIdx.z;     // image of dst
    const TS* base = reinterpret_cast<const TS*>(p.src) + (size_t)blockIdx.z * p.src_stride;
    float span = 0.f;
    if (MODE == MODE_MASK) span = p.span ? (p.span[1] - p.span[0]) : p.span_value;
    // tile load.  Rows whose length is a multiple of 4 (and a 16-byte aligned base) are read in aligned groups of 4 pixels
    // (one 4- or 16-byte load each; IN, R and the tile origin are multiples of 4, so a group lies entirely inside the row or
    // entirely beyond one end of it, where mode='nearest' repeats the end pixel): a quarter of the load instructions and of
    // the exposed latency of the scalar loop below (ncu, round 2: long_scoreboard was half of all stall samples).
    const bool vec = (p.w & 3) == 0 && (reinterpret_cast<size_t>(base) & 15) == 0 && (p.src_stride & 3) == 0;
    bool masked_any = false;                               // MODE_MASK: this thread loaded a masked (zero) mask pixel
    if (vec) {
        typedef typename Vec4<TS>::type V4;
        constexpr int GW = IN / 4, TOT = IN * GW, UN = 3;
        for (int i0 = tid; i0 < TOT; i0 += kNT * UN) {
            V4 v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * kNT;
                const int yy = i / GW, g = i - yy * GW;
                const int ysrc = min(max(y0 + yy - R, 0), p.h - 1), xs = x0 + 4 * g - R;
                if (i < TOT) {
                    const TS* row = base + (size_t)ysrc * p.w;
                    if (xs >= 0 && xs < p.w) {
                        v[u] = __ldg(reinterpret_cast<const V4*>(row + xs));
                    } else {
                        const TS e = __ldg(row + (xs < 0 ? 0 : p.w - 1));
                        v[u].x = e; v[u].y = e; v[u].z = e; v[u].w = e;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * kNT;
                const int yy = i / GW, g = i - yy * GW;
                if (i < TOT) {
                    float4 f;
                    if (MODE == MODE_MASK) {
                        f.x = v[u].x ? 0.f : span; f.y = v[u].y ? 0.f : span; f.z = v[u].z ? 0.f : span; f.w = v[u].w ? 0.f : span;
                        masked_any = masked_any || !(v[u].x && v[u].y && v[u].z && v[u].w);
                    }
                    else { f.x = (float)v[u].x; f.y = (float)v[u].y; f.z = (float)v[u].z; f.w = (float)v[u].w; }
                    *reinterpret_cast<float4*>(s_in + yy * PITCH + 4 * g) = f;
                }
            }
        }
    } else {
        // batches of independent loads (the tile load is pure latency otherwise): UN values per thread in flight
        constexpr int TOT = IN * IN, UN = 6;
        for (int i0 = tid; i0 < TOT; i0 += kNT * UN) {
            TS v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * kNT;
                const int yy = i / IN, xx = i - yy * IN;           // IN is a compile-time constant
                const int ysrc = min(max(y0 + yy - R, 0), p.h - 1), xsrc = min(max(x0 + xx - R, 0), p.w - 1);
                v[u] = i < TOT ? __ldg(base + (size_t)ysrc * p.w + xsrc) : TS(0);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * kNT;
                const int yy = i / IN, xx = i - yy * IN;
                if (i < TOT) {
                    s_in[yy * PITCH + xx] = MODE == MODE_MASK ? (v[u] ? 0.f : span) : (float)v[u];
                    if (MODE == MODE_MASK) masked_any = masked_any || !v[u];
                }
            }
        }
    }
    if (MODE == MODE_MASK) {
        // a tile (with its halo) without a masked pixel contributes a zero term: |x| - 0 keeps x bit for bit.  Blocks that
        // hang over the border of the mesh are masked along one edge only, most of their tiles stop here.
        if (!__syncthreads_or(masked_any)) return;
    } else {
        __syncthreads();
    }
    // row pass: item = (row yy, strip of 8 outputs); consecutive threads take consecutive rows (pitch = 4 mod 32: the
    // 128-bit loads of 8 consecutive rows fall into different banks)
    for (int i = tid; i < IN * (kFT / kFP); i += kNT) {
        const int strip = i / IN, yy = i - strip * IN;
        float acc[kFP];
        reg_line<R>(s_in + yy * PITCH + strip * kFP, tf, acc);
#pragma unroll
        for (int o = 0; o < kFP; ++o) s_t[(strip * kFP + o) * PITCH + yy] = acc[o];
    }
    __syncthreads();
    // column pass: item = (column xx, strip of 8 output rows); consecutive threads take consecutive columns
    for (int i = tid; i < kFT * (kFT / kFP); i += kNT) {
        const int ys = i / kFT, xx = i - ys * kFT;
        float acc[kFP];
        reg_line<R>(s_t + xx * PITCH + ys * kFP, tf, acc);
        const int x = x0 + xx;
        if (x >= p.w) continue;
#pragma unroll
        for (int o = 0; o < kFP; ++o) {
            const int yy = ys * kFP + o, y = y0 + yy;
            if (y >= p.h) break;
            const float g = acc[o];
            float* out = p.dst + ((size_t)img * p.h + y) * p.w + x;
            if (MODE == MODE_BLUR) {
                *out = g;
            } else if (MODE == MODE_DOG) {
                float f = __fsub_rn(s_in[(yy + R) * PITCH + xx + R], g);     // img0f - img1f
                *out = p.take_abs ? fabsf(f) : f;
            } else {
                const float mf = __fdiv_rn(__fmul_rn(g, p.sc2), p.s02);
                const float f = *out;
                float v = fmaxf(__fsub_rn(fabsf(f), mf), 0.f);
                if (!p.take_abs) v = f > 0.f ? v : (f < 0.f ? -v : __fmul_rn(v, 0.f));
                *out = v;
            }
        }
    }
}

// per-image {min, max}; one CTA per image
template <typename TS>
__global__ void __launch_bounds__(512) fbk_minmax(const TS* src, long long elems, float* out)
{
    __shared__ float s_lo[16], s_hi[16];
    const TS* base = src + (size_t)blockIdx.x * elems;
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = threadIdx.x; i < elems; i += 512) {
        const float v = (float)__ldg(base + i);
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    for (int off = 16; off; off >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 16; ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
        out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi;
    }
}

// {min, max} over rows of a [n][2] table -> out[0..1]
__global__ void fbk_minmax_fold(const float* table, int n, float* out)
{
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n; i += 32) { lo = fminf(lo, table[2 * i]); hi = fmaxf(hi, table[2 * i + 1]); }
    for (int off = 16; off; off >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (threadIdx.x == 0) { out[0] = lo; out[1] = hi; }
}

// cv2.resize(INTER_AREA) with fx = fy = 1/k: mean of the k x k cell; cells cut by the border are
// averaged over the pixels they contain (OpenCV resizeAreaFast_).  uint8: (sum + 2) >> 2 for k = 2,
// else round-half-even of float(sum) * float(1/k^2) resp. float(sum) / count.
template <typename TS>
__global__ void __launch_bounds__(256) fbk_resize_area(const TS* src, int n, int h, int w, int k, int oh, int ow, TS* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= ow) return;
    const TS* s = src + (size_t)img * h * w;
    const int ys = y * k, xs = x * k;
    const int ny = min(k, h - ys), nx = min(k, w - xs);
    TS out;
    if (sizeof(TS) == 1) {
        int sum = 0;
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) sum += (int)s[(size_t)(ys + j) * w + xs + i];
        float v;
        if (ny == k && nx == k) {
            if (k == 2) { dst[((size_t)img * oh + y) * ow + x] = (TS)((sum + 2) >> 2); return; }
            v = __fmul_rn((float)sum, 1.f / (float)(k * k));
        } else if (ny <= 0 || nx <= 0) {
            v = 0.f;
        } else {
            v = __fdiv_rn((float)sum, (float)(ny * nx));
        }
        int q = __float2int_rn(v);
        out = (TS)(q < 0 ? 0 : (q > 255 ? 255 : q));
    } else {
        float sum = 0.f;
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) sum = __fadd_rn(sum, (float)s[(size_t)(ys + j) * w + xs + i]);
        if (ny == k && nx == k) out = (TS)__fmul_rn(sum, 1.f / (float)(k * k));
        else out = (TS)((ny <= 0 || nx <= 0) ? 0.f : __fdiv_rn(sum, (float)(ny * nx)));
    }
    dst[((size_t)img * oh + y) * ow + x] = out;
}

// cv2.resize(INTER_NEAREST): dst(y, x) = src(min(floor(y * ify), h - 1), min(floor(x * ifx), w - 1))
__global__ void __launch_bounds__(256) fbk_resize_nearest(const unsigned char* src, int n, int h, int w, double ify, double ifx,
                                                          int oh, int ow, unsigned char* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= ow) return;
    const int sy = min((int)floor(y * ify), h - 1), sx = min((int)floor(x * ifx), w - 1);
    dst[((size_t)img * oh + y) * ow + x] = src[((size_t)img * h + sy) * w + sx];
}

struct CropParams {
    const void* img;
    int ih, iw;
    const double* blocks;  // [n][10]: x0, y0, step_x, step_y, A00, A10, t0, A01, A11, t1
    int n, bh, bw;
    double ox, oy;         // integer-valued origin of the reference's source crop (common.py:300-304)
    float fill;
    void* out;
    int has_cover;         // only pixels whose source position lies STRICTLY inside (cx0, cx1) x (cy0, cy1) are rendered
    double cx0, cy0, cx1, cy1;
    const int* block_slot;             // optional [n]: < 0 = the whole block counts as covered (renderer.py:443-444), else the
                                       // image of mask_out that receives the block's mask (compact: only partial blocks have one)
    unsigned char* mask_out;
    const fb_crop_src* srcs;           // optional [n]: per-block source image and origin (blocks of many images in one launch)
};

struct CropSrc {                       // the source of one block, resolved from CropParams / fb_crop_src
    const void* img; int ih, iw; double ox, oy;
};
__device__ __forceinline__ CropSrc crop_source(const CropParams& p, int b)
{
    if (p.srcs) { const fb_crop_src s = p.srcs[b]; return CropSrc{s.img, s.ih, s.iw, s.origin_x, s.origin_y}; }
    return CropSrc{p.img, p.ih, p.iw, p.ox, p.oy};
}

template <typename TS> __device__ __forceinline__ float fetch(const CropSrc& c, float fill, const TS* img, int y, int x)
{
    if ((unsigned)y < (unsigned)c.ih && (unsigned)x < (unsigned)c.iw) return (float)__ldg(img + (size_t)y * c.iw + x);
    return fill;
}

// renderer.py:419-450 (float64 field), common.py:318-321 (minus origin, float32), cv2.remap INTER_LINEAR:
// fixed point with 5 fractional bits, float weights for float images, 15-bit integer weights for uint8.
template <typename TS>
__global__ void __launch_bounds__(256) fbk_crop_blocks(const __grid_constant__ CropParams p)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.bh * p.bw) return;
    const int row = i / p.bw, col = i - row * p.bw;
    const double* q = p.blocks + (size_t)b * 10;
    const double xx = __dadd_rn(q[0], __dmul_rn((double)col, q[2]));
    const double yy = __dadd_rn(q[1], __dmul_rn((double)row, q[3]));
    const double xs = __dadd_rn(__dadd_rn(__dmul_rn(xx, q[4]), __dmul_rn(yy, q[5])), q[6]);
    const double ys = __dadd_rn(__dadd_rn(__dmul_rn(xx, q[7]), __dmul_rn(yy, q[8])), q[9]);
    if (p.has_cover) {
        const int slot = p.block_slot ? p.block_slot[b] : b;
        const bool inside = slot < 0 || (xs > p.cx0 && xs < p.cx1 && ys > p.cy0 && ys < p.cy1);
        if (p.mask_out && slot >= 0) p.mask_out[((size_t)slot * p.bh + row) * p.bw + col] = inside ? 1 : 0;
        if (!inside) {
            reinterpret_cast<TS*>(p.out)[((size_t)b * p.bh + row) * p.bw + col] = (TS)p.fill;
            return;
        }
    }
    const CropSrc c = crop_source(p, b);
    const float xf = (float)__dsub_rn(xs, c.ox), yf = (float)__dsub_rn(ys, c.oy);
    const int fx = __float2int_rn(__fmul_rn(xf, 32.f)), fy = __float2int_rn(__fmul_rn(yf, 32.f));
    const int ix = (fx >> 5) + (int)c.ox, iy = (fy >> 5) + (int)c.oy;
    const int ax = fx & 31, ay = fy & 31;
    const TS* img = reinterpret_cast<const TS*>(c.img);
    const float v00 = fetch<TS>(c, p.fill, img, iy, ix), v01 = fetch<TS>(c, p.fill, img, iy, ix + 1);
    const float v10 = fetch<TS>(c, p.fill, img, iy + 1, ix), v11 = fetch<TS>(c, p.fill, img, iy + 1, ix + 1);
    TS* out = reinterpret_cast<TS*>(p.out) + ((size_t)b * p.bh + row) * p.bw + col;
    if (sizeof(TS) == 1) {
        // weights (32 - a) * (32 - b) * 32 sum to 2^15 exactly
        const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
        const int acc = (int)v00 * w00 + (int)v01 * w01 + (int)v10 * w10 + (int)v11 * w11;
        const int r = (acc + (1 << 14)) >> 15;
        *out = (TS)(r < 0 ? 0 : (r > 255 ? 255 : r));
    } else {
        const float cx1 = (float)ax * (1.f / 32.f), cy1 = (float)ay * (1.f / 32.f);
        const float cx0 = 1.f - cx1, cy0 = 1.f - cy1;
        const float w00 = __fmul_rn(cy0, cx0), w01 = __fmul_rn(cy0, cx1), w10 = __fmul_rn(cy1, cx0), w11 = __fmul_rn(cy1, cx1);
        float acc = __fmul_rn(v00, w00);
        acc = __fadd_rn(acc, __fmul_rn(v01, w01));
        acc = __fadd_rn(acc, __fmul_rn(v10, w10));
        acc = __fadd_rn(acc, __fmul_rn(v11, w11));
        *out = (TS)acc;
    }
}

// Four consecutive pixels of a row per thread (bw % 4 == 0): the same arithmetic per pixel as fbk_crop_blocks, with the
// row terms of the coordinate field shared and one 4-wide store of the pixels (and of the coverage mask).

template <typename TS>
__global__ void __launch_bounds__(256) fbk_crop_blocks4(const __grid_constant__ CropParams p)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int qw = p.bw >> 2;
    if (i >= p.bh * qw) return;
    const int row = (int)(((float)i + 0.5f) * (1.0f / (float)qw)), col0 = (i - row * qw) << 2;
    const double* q = p.blocks + (size_t)b * 10;
    const double q0 = q[0], q2 = q[2], a00 = q[4], a01 = q[7], t0 = q[6], t1 = q[9];
    const double yy = __dadd_rn(q[1], __dmul_rn((double)row, q[3]));
    const double ya = __dmul_rn(yy, q[5]), yb = __dmul_rn(yy, q[8]);
    const CropSrc c = crop_source(p, b);
    const TS* img = reinterpret_cast<const TS*>(c.img);
    TS px[4];
    unsigned char inside4[4];
    const int slot = (p.has_cover && p.block_slot) ? p.block_slot[b] : b;
    const bool whole = p.has_cover && slot < 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double xx = __dadd_rn(q0, __dmul_rn((double)(col0 + k), q2));
        const double xs = __dadd_rn(__dadd_rn(__dmul_rn(xx, a00), ya), t0);
        const double ys = __dadd_rn(__dadd_rn(__dmul_rn(xx, a01), yb), t1);
        bool inside = true;
        if (p.has_cover) inside = whole || (xs > p.cx0 && xs < p.cx1 && ys > p.cy0 && ys < p.cy1);
        inside4[k] = inside ? 1 : 0;
        if (!inside) { px[k] = (TS)p.fill; continue; }
        const float xf = (float)__dsub_rn(xs, c.ox), yf = (float)__dsub_rn(ys, c.oy);
        const int fx = __float2int_rn(__fmul_rn(xf, 32.f)), fy = __float2int_rn(__fmul_rn(yf, 32.f));
        const int ix = (fx >> 5) + (int)c.ox, iy = (fy >> 5) + (int)c.oy;
        const int ax = fx & 31, ay = fy & 31;
        const float v00 = fetch<TS>(c, p.fill, img, iy, ix), v01 = fetch<TS>(c, p.fill, img, iy, ix + 1);
        const float v10 = fetch<TS>(c, p.fill, img, iy + 1, ix), v11 = fetch<TS>(c, p.fill, img, iy + 1, ix + 1);
        if (sizeof(TS) == 1) {
            const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
            const int acc = (int)v00 * w00 + (int)v01 * w01 + (int)v10 * w10 + (int)v11 * w11;
            const int r = (acc + (1 << 14)) >> 15;
            px[k] = (TS)(r < 0 ? 0 : (r > 255 ? 255 : r));
        } else {
            const float cx1 = (float)ax * (1.f / 32.f), cy1 = (float)ay * (1.f / 32.f);
            const float cx0 = 1.f - cx1, cy0 = 1.f - cy1;
            const float w00 = __fmul_rn(cy0, cx0), w01 = __fmul_rn(cy0, cx1), w10 = __fmul_rn(cy1, cx0), w11 = __fmul_rn(cy1, cx1);
            float acc = __fmul_rn(v00, w00);
            acc = __fadd_rn(acc, __fmul_rn(v01, w01));
            acc = __fadd_rn(acc, __fmul_rn(v10, w10));
            acc = __fadd_rn(acc, __fmul_rn(v11, w11));
            px[k] = (TS)acc;
        }
    }
    const size_t o = ((size_t)b * p.bh + row) * p.bw + col0;          // a multiple of 4: bw % 4 == 0
    typename Vec4<TS>::type v;
    v.x = px[0]; v.y = px[1]; v.z = px[2]; v.w = px[3];
    *reinterpret_cast<typename Vec4<TS>::type*>(reinterpret_cast<TS*>(p.out) + o) = v;
    if (p.has_cover && p.mask_out && slot >= 0)
        *reinterpret_cast<uchar4*>(p.mask_out + ((size_t)slot * p.bh + row) * p.bw + col0) = make_uchar4(inside4[0], inside4[1], inside4[2], inside4[3]);
}

int check_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fb_failf(FB_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fb_failf(FB_EINVAL, "device %d out of range (%d devices)", device, ndev);
    FB_CU(cudaSetDevice(device));
    return FB_OK;
}

bool g_gauss_legacy = false;      // FB_GAUSS_LEGACY=1: the unblocked float kernel (comparison runs)

size_t gauss_smem(int r)
{
    const int in_w = kTW + 2 * r, in_h = kTH + 2 * r;
    return ((size_t)in_h * (in_w | 1) + (size_t)in_h * kTW) * sizeof(float);
}

template <typename TS, int MODE, typename ACC>
int launch_gauss(const GaussParams& p, cudaStream_t st)
{
    const size_t smem = gauss_smem(p.taps.radius);
    FB_CU(cudaFuncSetAttribute(fbk_gauss2d<TS, MODE, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.w + kTW - 1) / kTW, (p.h + kTH - 1) / kTH, p.n);
    fbk_gauss2d<TS, MODE, ACC><<<grid, kNT, smem, st>>>(p);
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

size_t gaussf_smem(int r)
{
    const int in_h = kFT + 2 * r;
    return ((size_t)in_h * gaussf_in_pitch(r) + 2 * kSlack + (size_t)(in_h + 2 * kSlack) * kRowPitch) * sizeof(float);
}

template <typename TS, int MODE, int CH>
int launch_gauss_f32_ch(const GaussParams& p, GaussTapsF& tf, cudaStream_t st)
{
    const int r = p.taps.radius;
    tf.rpad = (r + CH - 1) / CH * CH;
    const size_t smem = gaussf_smem(r);
    FB_CU(cudaFuncSetAttribute(fbk_gauss2d_f32<TS, MODE, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.w + kFT - 1) / kFT, (p.h + kFT - 1) / kFT, p.n);
    fbk_gauss2d_f32<TS, MODE, CH><<<grid, kNT, smem, st>>>(p, tf);
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

template <typename TS, int MODE>
int launch_gauss_f32(const GaussParams& p, cudaStream_t st)
{
    GaussTapsF tf{};
    const int r = p.taps.radius;
    tf.wc = (float)p.taps.w[r];
    for (int j = 0; j < r; ++j) tf.wt[j] = (float)p.taps.w[j];
    // chunk size with the least zero padding of the tap list (the larger one on ties); padded reads stay inside
    // the slack of at most 7 values either side for every choice
    int best = 8, pad = (8 - r % 8) % 8;
    for (int ch = 7; ch >= 5; --ch) { const int q = (ch - r % ch) % ch; if (q < pad) { pad = q; best = ch; } }
    switch (best) {
        case 5: return launch_gauss_f32_ch<TS, MODE, 5>(p, tf, st);
        case 6: return launch_gauss_f32_ch<TS, MODE, 6>(p, tf, st);
        case 7: return launch_gauss_f32_ch<TS, MODE, 7>(p, tf, st);
        default: return launch_gauss_f32_ch<TS, MODE, 8>(p, tf, st);
    }
}

template <typename TS, int MODE, int R>
int launch_gauss_reg_r(const GaussParams& p, cudaStream_t st)
{
    RegTaps tf{};
    const int r = p.taps.radius;
    tf.wc = (float)p.taps.w[r];
    for (int j = 0; j < R; ++j) tf.w[j] = j < R - r ? 0.f : (float)p.taps.w[j - (R - r)];     // zero taps on the outside
    const size_t smem = RegGauss<R>::SMEM;
    FB_CU(cudaFuncSetAttribute(fbk_gauss2d_reg<TS, MODE, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((p.w + kFT - 1) / kFT, (p.h + kFT - 1) / kFT, p.n);
    fbk_gauss2d_reg<TS, MODE, R><<<grid, kNT, smem, st>>>(p, tf);
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

template <typename TS, int MODE>
int launch_gauss_reg(const GaussParams& p, cudaStream_t st)
{
    const int r = p.taps.radius;
    if (r <= 8) return launch_gauss_reg_r<TS, MODE, 8>(p, st);
    if (r <= 12) return launch_gauss_reg_r<TS, MODE, 12>(p, st);
    if (r <= 16) return launch_gauss_reg_r<TS, MODE, 16>(p, st);
    if (r <= 20) return launch_gauss_reg_r<TS, MODE, 20>(p, st);
    return launch_gauss_reg_r<TS, MODE, 24>(p, st);
}

bool g_gauss_blocked = false;     // FB_GAUSS_BLOCKED=1: the shared-memory window kernel of round 1 (comparison runs)

template <typename TS, int MODE>
int launch_gauss_acc(const GaussParams& p, bool exact, cudaStream_t st)
{
    if (exact) return launch_gauss<TS, MODE, double>(p, st);
    if (g_gauss_legacy) return launch_gauss<TS, MODE, float>(p, st);
    if (p.taps.radius <= 24 && !g_gauss_blocked) return launch_gauss_reg<TS, MODE>(p, st);
    return launch_gauss_f32<TS, MODE>(p, st);
}

}  // namespace

extern "C" long long fb_masked_dog_workspace(int n, int h, int w)
{
    if (n < 0 || h < 1 || w < 1) return 0;
    return (long long)n * h * w * (long long)sizeof(float) + (long long)(n + 1) * 2 * sizeof(float) + 256;
}

extern "C" int fb_masked_dog(const void* img, const unsigned char* mask, int n, int h, int w, int in_dtype, int mask_n,
                             double sigma, double ptp, int flags, float* out, void* work, long long work_bytes,
                             int device, void* stream)
{
    return fb_masked_dog_sparse(img, mask, nullptr, n, h, w, in_dtype, mask_n, sigma, ptp, flags, out, work, work_bytes, device, stream);
}

extern "C" int fb_masked_dog_sparse(const void* img, const unsigned char* mask, const int* mask_images, int n, int h, int w, int in_dtype,
                                    int mask_n, double sigma, double ptp, int flags, float* out, void* work, long long work_bytes,
                                    int device, void* stream)
{
    if (n < 0 || h < 1 || w < 1) return fb_failf(FB_EINVAL, "bad shape n=%d %dx%d", n, h, w);
    g_gauss_legacy = getenv("FB_GAUSS_LEGACY") != nullptr;
    g_gauss_blocked = getenv("FB_GAUSS_BLOCKED") != nullptr;
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "masked_dog: dtype %d not supported (float32 / uint8)", in_dtype);
    if (mask && !mask_images && mask_n != 1 && mask_n != n) return fb_failf(FB_EINVAL, "mask_n must be 1 or n");
    if (mask && mask_images && (mask_n < 1 || mask_n > n)) return fb_failf(FB_EINVAL, "mask_n must be 1 .. n with an image list");
    if (n == 0) return FB_OK;
    if (!img || !out || !work) return fb_failf(FB_EINVAL, "null pointer");
    if (work_bytes < fb_masked_dog_workspace(n, h, w)) return fb_failf(FB_EINVAL, "workspace too small (%lld < %lld)", work_bytes, fb_masked_dog_workspace(n, h, w));
    GaussParams p{};
    if (!make_taps(sigma, p.taps)) return fb_failf(FB_ESIZE, "sigma %g outside (0, %g]", sigma, (kMaxRadius + 0.49) / 4.0);
    int rc = check_device(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const bool exact = flags & FB_DOG_EXACT;
    float* blur = reinterpret_cast<float*>(work);
    float* table = blur + (size_t)n * h * w;                      // [n][2] + [2]
    p.n = n; p.h = h; p.w = w; p.src_stride = (long long)h * w;
    // first blur
    p.src = img; p.dst = blur;
    rc = in_dtype == FB_F32 ? launch_gauss_acc<float, MODE_BLUR>(p, exact, st) : launch_gauss_acc<unsigned char, MODE_BLUR>(p, exact, st);
    if (rc != FB_OK) return rc;
    // second blur and difference
    p.src = blur; p.dst = out; p.take_abs = (!mask && (flags & FB_DOG_UNSIGNED)) ? 1 : 0;
    rc = launch_gauss_acc<float, MODE_DOG>(p, exact, st);
    if (rc != FB_OK || !mask) return rc;
    // suppression of the signal bleeding in from outside the mask
    if (ptp != ptp) {
        if (in_dtype == FB_F32) fbk_minmax<float><<<n, 512, 0, st>>>(reinterpret_cast<const float*>(img), (long long)h * w, table);
        else fbk_minmax<unsigned char><<<n, 512, 0, st>>>(reinterpret_cast<const unsigned char*>(img), (long long)h * w, table);
        fbk_minmax_fold<<<1, 32, 0, st>>>(table, n, table + 2 * (size_t)n);
        fb_count_launches(2);
        p.span = table + 2 * (size_t)n;
    } else {
        p.span = nullptr; p.span_value = (float)ptp;
    }
    const double sigma_c = sqrt(sigma * sigma + sigma * sigma);
    if (!make_taps(sigma_c, p.taps)) return fb_failf(FB_ESIZE, "sigma %g too large for the mask term", sigma);
    p.sc2 = (float)(sigma_c * sigma_c); p.s02 = (float)(sigma * sigma);
    p.src = mask; p.dst = out; p.take_abs = (flags & FB_DOG_UNSIGNED) ? 1 : 0;
    if (mask_images) {
        // only the listed images carry masked pixels: for the others the term is zero and |x| - 0 keeps x as it is
        p.n = mask_n; p.dst_index = mask_images; p.src_stride = (long long)h * w;
        if ((flags & FB_DOG_UNSIGNED) && mask_n < n) return fb_failf(FB_EINVAL, "unsigned output needs the mask pass on every image");
    } else {
        p.src_stride = mask_n == 1 ? 0 : (long long)h * w;
    }
    return launch_gauss_acc<unsigned char, MODE_MASK>(p, exact, st);
}

extern "C" int fb_stack_minmax(const void* stack, int n, long long elems, int in_dtype, float* minmax, int device, void* stream)
{
    if (n < 0 || elems < 1) return fb_failf(FB_EINVAL, "bad shape");
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "stack_minmax: dtype %d not supported", in_dtype);
    if (n == 0) return FB_OK;
    if (!stack || !minmax) return fb_failf(FB_EINVAL, "null pointer");
    int rc = check_device(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == FB_F32) fbk_minmax<float><<<n, 512, 0, st>>>(reinterpret_cast<const float*>(stack), elems, minmax);
    else fbk_minmax<unsigned char><<<n, 512, 0, st>>>(reinterpret_cast<const unsigned char*>(stack), elems, minmax);
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_resize_area(const void* src, int n, int h, int w, int in_dtype, int k, void* dst, int oh, int ow,
                              int device, void* stream)
{
    if (n < 0 || h < 1 || w < 1 || k < 1 || oh < 1 || ow < 1) return fb_failf(FB_EINVAL, "bad shape");
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "resize_area: dtype %d not supported", in_dtype);
    if ((long long)(oh - 1) * k >= h || (long long)(ow - 1) * k >= w) return fb_failf(FB_EINVAL, "output %dx%d too large for %dx%d / %d", oh, ow, h, w, k);
    if (n == 0) return FB_OK;
    if (!src || !dst) return fb_failf(FB_EINVAL, "null pointer");
    int rc = check_device(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((ow + 255) / 256, oh, n);
    if (in_dtype == FB_F32) fbk_resize_area<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), n, h, w, k, oh, ow, reinterpret_cast<float*>(dst));
    else fbk_resize_area<unsigned char><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(src), n, h, w, k, oh, ow, reinterpret_cast<unsigned char*>(dst));
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_resize_nearest(const unsigned char* src, int n, int h, int w, double inv_fy, double inv_fx,
                                 unsigned char* dst, int oh, int ow, int device, void* stream)
{
    if (n < 0 || h < 1 || w < 1 || oh < 1 || ow < 1 || !(inv_fx > 0) || !(inv_fy > 0)) return fb_failf(FB_EINVAL, "bad shape");
    if (n == 0) return FB_OK;
    if (!src || !dst) return fb_failf(FB_EINVAL, "null pointer");
    int rc = check_device(device);
    if (rc != FB_OK) return rc;
    dim3 grid((ow + 255) / 256, oh, n);
    fbk_resize_nearest<<<grid, 256, 0, (cudaStream_t)stream>>>(src, n, h, w, inv_fy, inv_fx, oh, ow, dst);
    fb_count_launches(1);
    FB_CU(cudaGetLastError());
    return FB_OK;
}

static int crop_launch(CropParams& p, int in_dtype, double fillval, int device, void* stream)
{
    const int n = p.n, bh = p.bh, bw = p.bw;
    if (n > 65535 * 64) return fb_failf(FB_EINVAL, "too many blocks in one call");
    int rc = check_device(device);
    if (rc != FB_OK) return rc;
    p.fill = in_dtype == FB_U8 ? (float)(fillval < 0 ? 0 : (fillval > 255 ? 255 : rint(fillval))) : (float)fillval;
    const bool vec4 = bw % 4 == 0 && ((size_t)p.out % 16 == 0) && (!p.mask_out || (size_t)p.mask_out % 4 == 0) && !getenv("FB_CROP_SCALAR");
    const size_t esz = in_dtype == FB_U8 ? 1 : 4;
    // grid.y carries the block index (at most 65535 per launch): long lists go out in slices
    for (int lo = 0; lo < n; lo += 65535) {
        CropParams q = p;
        q.n = n - lo < 65535 ? n - lo : 65535;
        q.blocks = p.blocks + (size_t)lo * 10;
        q.out = (char*)p.out + (size_t)lo * bh * bw * esz;
        if (p.mask_out && !p.block_slot) q.mask_out = p.mask_out + (size_t)lo * bh * bw;
        if (p.block_slot) q.block_slot = p.block_slot + lo;
        if (p.srcs) q.srcs = p.srcs + lo;
        if (vec4) {
            dim3 grid((bh * (bw / 4) + 255) / 256, q.n);
            if (in_dtype == FB_F32) fbk_crop_blocks4<float><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
            else fbk_crop_blocks4<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
        } else {
            dim3 grid((bh * bw + 255) / 256, q.n);
            if (in_dtype == FB_F32) fbk_crop_blocks<float><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
            else fbk_crop_blocks<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>(q);
        }
        fb_count_launches(1);
    }
    FB_CU(cudaGetLastError());
    return FB_OK;
}

extern "C" int fb_crop_blocks(const void* img, int ih, int iw, int in_dtype, const double* blocks, int n, int bh, int bw,
                              double origin_x, double origin_y, double fillval, void* out,
                              const double* cover, const int* block_slot, unsigned char* mask_out, int device, void* stream)
{
    if (n < 0 || ih < 1 || iw < 1 || bh < 1 || bw < 1) return fb_failf(FB_EINVAL, "bad shape");
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "crop_blocks: dtype %d not supported", in_dtype);
    if (n == 0) return FB_OK;
    if (!img || !blocks || !out) return fb_failf(FB_EINVAL, "null pointer");
    CropParams p{};
    p.img = img; p.ih = ih; p.iw = iw; p.blocks = blocks; p.n = n; p.bh = bh; p.bw = bw;
    p.ox = origin_x; p.oy = origin_y; p.out = out;
    p.has_cover = cover ? 1 : 0; p.mask_out = cover ? mask_out : nullptr;
    if (cover) { p.cx0 = cover[0]; p.cy0 = cover[1]; p.cx1 = cover[2]; p.cy1 = cover[3]; p.block_slot = block_slot; }
    return crop_launch(p, in_dtype, fillval, device, stream);
}

extern "C" int fb_crop_blocks_multi(const fb_crop_src* sources, int in_dtype, const double* blocks, int n, int bh, int bw,
                                    double fillval, void* out, int device, void* stream)
{
    if (n < 0 || bh < 1 || bw < 1) return fb_failf(FB_EINVAL, "bad shape");
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "crop_blocks_multi: dtype %d not supported", in_dtype);
    if (n == 0) return FB_OK;
    if (!sources || !blocks || !out) return fb_failf(FB_EINVAL, "null pointer");
    CropParams p{};
    p.srcs = sources; p.blocks = blocks; p.n = n; p.bh = bh; p.bw = bw; p.out = out;
    return crop_launch(p, in_dtype, fillval, device, stream);
}
