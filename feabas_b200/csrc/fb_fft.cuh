// fb_fft.cuh -- in-place mixed-radix (2,3,4,5,6,8,9,10,12,15,16) FFT passes over shared-memory tiles.
//
// The tile holds `nl` independent lines interleaved line-minor:
//     element e of line l lives at  s[e * pitch + l]
// so that the threads of a warp, which are mapped line-minor as well, touch
// consecutive shared-memory words in every pass (pitch is chosen odd by the
// callers that transpose rows into the tile, which also makes that transpose
// conflict free).
//
// Forward transforms are decimation-in-frequency (natural order in, digit
// reversed order out); inverse transforms are the exact mirror image
// (decimation-in-time, digit reversed in, natural out).  Point-wise products
// in between therefore need no reordering, and the two places that pair
// frequency k with N-k (real/complex packing of image rows) go through the
// plan's `pos` table.
//
// Everything here compiles both as CUDA device code and as plain C++ (the
// host emulator under tests/host_emu runs the same bodies with a thread team),
// hence the FB_* macros instead of raw CUDA built-ins.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FB_HD __host__ __device__ __forceinline__
#define FB_DEV __device__ __forceinline__
#else
#define FB_HD inline
#define FB_DEV inline
#endif

#if defined(__CUDA_ARCH__)
#define FB_LDG(p) __ldg(p)
#else
#define FB_LDG(p) (*(p))
#endif

namespace fb {

constexpr int kMaxPass = 14;

template <typename T> struct alignas(2 * sizeof(T)) cx { T x, y; };

template <typename T> FB_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename T> FB_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> FB_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> FB_HD cx<T> cmul(cx<T> a, cx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
template <typename T> FB_HD cx<T> cmulc(cx<T> a, cx<T> b) { return mk<T>(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
template <typename T> FB_HD cx<T> cscale(cx<T> a, T s) { return mk<T>(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <typename T, bool INV> FB_HD cx<T> rot(cx<T> z) { return INV ? mk<T>(-z.y, z.x) : mk<T>(z.y, -z.x); }

#if defined(__CUDA_ARCH__)
FB_DEV cx<float> ldg(const cx<float>* p) { float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<float>(v.x, v.y); }
FB_DEV cx<double> ldg(const cx<double>* p) { double2 v = __ldg(reinterpret_cast<const double2*>(p)); return mk<double>(v.x, v.y); }
#else
template <typename T> inline cx<T> ldg(const cx<T>* p) { return *p; }
#endif

// 1-D plan as the kernels see it (passed by value inside the launch params).
struct Plan1D {
    int n;                  // transform length
    int npass;              // number of radix passes
    int radix[kMaxPass];    // DIF order
    const void* tw;         // cx<T>[n], tw[k] = exp(-2 pi i k / n)
    const int* pos;         // pos[k] = position of natural frequency k after the forward passes
};

// ----------------------------------------------------------------------------
// radix butterflies: v[t] <- sum_q v[q] * exp(-/+ 2 pi i q t / R)
// ----------------------------------------------------------------------------
template <typename T, int R, bool INV> struct Bfly;

template <typename T, bool INV> struct Bfly<T, 2, INV> {
    static FB_HD void run(cx<T>* v) { cx<T> a = v[0]; v[0] = a + v[1]; v[1] = a - v[1]; }
};

template <typename T, bool INV> struct Bfly<T, 3, INV> {
    static FB_HD void run(cx<T>* v) {
        const T c = T(-0.5);
        const T s = INV ? T(0.86602540378443864676) : T(-0.86602540378443864676);
        cx<T> t = v[1] + v[2], u = v[1] - v[2];
        cx<T> m = mk<T>(v[0].x + c * t.x, v[0].y + c * t.y);
        cx<T> r = mk<T>(-s * u.y, s * u.x);      // i * s * u
        v[0] = v[0] + t; v[1] = m + r; v[2] = m - r;
    }
};

template <typename T, bool INV> struct Bfly<T, 4, INV> {
    static FB_HD void run(cx<T>* v) {
        cx<T> a = v[0] + v[2], b = v[0] - v[2], c = v[1] + v[3], d = rot<T, INV>(v[1] - v[3]);
        v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d;
    }
};

template <typename T, bool INV> struct Bfly<T, 5, INV> {
    static FB_HD void run(cx<T>* v) {
        const T c1 = T(0.30901699437494742410), c2 = T(-0.80901699437494742410);
        const T s1 = T(0.95105651629515357212), s2 = T(0.58778525229247312917);
        cx<T> t1 = v[1] + v[4], t2 = v[2] + v[3], u1 = v[1] - v[4], u2 = v[2] - v[3];
        cx<T> a1 = mk<T>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        cx<T> a2 = mk<T>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        cx<T> b1 = mk<T>(s1 * u1.x + s2 * u2.x, s1 * u1.y + s2 * u2.y);
        cx<T> b2 = mk<T>(s2 * u1.x - s1 * u2.x, s2 * u1.y - s1 * u2.y);
        cx<T> r1 = rot<T, INV>(b1), r2 = rot<T, INV>(b2);   // -/+ i b
        v[0] = v[0] + t1 + t2;
        v[1] = a1 + r1; v[4] = a1 - r1; v[2] = a2 + r2; v[3] = a2 - r2;
    }
};

template <typename T, bool INV> struct Bfly<T, 8, INV> {
    static FB_HD void run(cx<T>* v) {
        const T h = T(0.70710678118654752440);
        cx<T> e[4] = {v[0], v[2], v[4], v[6]};
        cx<T> o[4] = {v[1], v[3], v[5], v[7]};
        Bfly<T, 4, INV>::run(e);
        Bfly<T, 4, INV>::run(o);
        // o[k] *= w8^k,  w8 = exp(-/+ i pi / 4)
        cx<T> o1 = INV ? mk<T>(h * (o[1].x - o[1].y), h * (o[1].x + o[1].y))
                       : mk<T>(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
        cx<T> o2 = rot<T, INV>(o[2]);
        cx<T> o3 = INV ? mk<T>(-h * (o[3].x + o[3].y), h * (o[3].x - o[3].y))
                       : mk<T>(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
        v[0] = e[0] + o[0]; v[4] = e[0] - o[0];
        v[1] = e[1] + o1;   v[5] = e[1] - o1;
        v[2] = e[2] + o2;   v[6] = e[2] - o2;
        v[3] = e[3] + o3;   v[7] = e[3] - o3;
    }
};

}  // namespace fb
#include "fb_gfft.cuh"     // composite radices 6, 9, 10, 12, 15, 16 (mixed-radix register FFT)
namespace fb {

// ----------------------------------------------------------------------------
// one pass over a tile.  L = current sub-transform length (N for the first
// forward pass), m = L / R.  Thread i handles butterfly (line = i % nl,
// bj = i / nl) for i = tid, tid + nthr, ...
// ----------------------------------------------------------------------------
template <typename T, int R, bool INV>
FB_DEV void fft_pass(cx<T>* s, int pitch, int nl, int N, int L, const cx<T>* tw, int tid, int nthr)
{
    const int m = L / R;
    const int tws = N / L;
    const int total = (N / R) * nl;
    const float inv_nl = 1.0f / (float)nl, inv_m = 1.0f / (float)m;
    for (int i = tid; i < total; i += nthr) {
        // exact floor divisions for the index ranges used here (i < 2^22)
        int bj = (int)(((float)i + 0.5f) * inv_nl);
        int line = i - bj * nl;
        int b = (int)(((float)bj + 0.5f) * inv_m);
        int j = bj - b * m;
        cx<T>* p = s + (size_t)(b * L + j) * pitch + line;
        const int st = m * pitch;
        cx<T> v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p[q * st];
        if (!INV) {
            Bfly<T, R, false>::run(v);
            if (j != 0) {
#pragma unroll
                for (int t = 1; t < R; ++t) v[t] = cmul(v[t], ldg(tw + j * t * tws));
            }
        } else {
            if (j != 0) {
#pragma unroll
                for (int t = 1; t < R; ++t) v[t] = cmulc(v[t], ldg(tw + j * t * tws));
            }
            Bfly<T, R, true>::run(v);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) p[q * st] = v[q];
    }
}

// radices 2, 3, 4, 5, 8 only, inlined: the staged kernels (two 512-thread CTAs per SM need <= 64 registers)
template <typename T, bool INV>
FB_DEV void fft_pass_basic(int R, cx<T>* s, int pitch, int nl, int N, int L, const cx<T>* tw, int tid, int nthr)
{
    switch (R) {
        case 2: fft_pass<T, 2, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 3: fft_pass<T, 3, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 4: fft_pass<T, 4, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 5: fft_pass<T, 5, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        default: fft_pass<T, 8, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
    }
}

// All radices (the fused kernel: one CTA per SM, registers are not the limit).
// One out-of-line copy per (type, direction): the kernels call it from several places and the composite
// radices unroll to a few hundred instructions each.
#if defined(__CUDACC__)
#define FB_NOINLINE __device__ __noinline__
#else
#define FB_NOINLINE inline
#endif
template <typename T, bool INV>
FB_NOINLINE void fft_pass_any(int R, cx<T>* s, int pitch, int nl, int N, int L, const cx<T>* tw, int tid, int nthr)
{
#if defined(__CUDA_ARCH__)
    __builtin_assume(__isShared(s));        // out of line: keep LDS / STS instead of generic loads and stores
#endif
    switch (R) {
        case 2: fft_pass<T, 2, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 3: fft_pass<T, 3, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 4: fft_pass<T, 4, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 5: fft_pass<T, 5, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 6: fft_pass<T, 6, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 8: fft_pass<T, 8, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 9: fft_pass<T, 9, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 10: fft_pass<T, 10, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 12: fft_pass<T, 12, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        case 15: fft_pass<T, 15, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
        default: fft_pass<T, 16, INV>(s, pitch, nl, N, L, tw, tid, nthr); break;
    }
}

}  // namespace fb
