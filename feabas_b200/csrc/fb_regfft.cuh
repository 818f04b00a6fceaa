// fb_regfft.cuh -- register-resident FFTs of 8 / 16 / 32 / 64 complex points per thread.
//
// Radix-2 decimation-in-frequency, fully unrolled, twiddles as compile-time
// constants with the trivial ones (1, -i, exp(-i pi/4)) special-cased.  The
// transform is in place on v[0..N) and leaves X[k] in v[bitrev(k)]; callers
// index the result through brev<N>(k), which folds away at compile time.
#pragma once
#include "fb_fft.cuh"

namespace fb {

template <int N> FB_HD constexpr int brev(int k)
{
    int r = 0;
    for (int b = 1; b < N; b <<= 1) { r = (r << 1) | (k & 1); k >>= 1; }
    return r;
}

// cos / sin of 2 pi j / 64 for j = 0..16 (first quadrant), double precision literals
FB_HD constexpr double cos64(int j)
{
    constexpr double c[17] = {1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
                              0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708,
                              0.77301045336273696081, 0.70710678118654752440, 0.63439328416364549822,
                              0.55557023301960222474, 0.47139673682599764856, 0.38268343236508977173,
                              0.29028467725446236764, 0.19509032201612826785, 0.09801714032956060199, 0.0};
    // reduce j (mod 64) to the first quadrant
    j &= 63;
    int q = j >> 4, r = j & 15;
    double cs = c[r], sn = c[16 - r];
    return q == 0 ? cs : (q == 1 ? -sn : (q == 2 ? -cs : sn));
}
FB_HD constexpr double sin64(int j) { return cos64(j - 16); }

// v *= exp(-/+ 2 pi i j / n), n in {2, 4, 8, 16, 32, 64}
template <typename T, int N, int J, bool INV> FB_HD cx<T> twiddle_const(cx<T> v)
{
    constexpr int j64 = (J * (64 / N)) & 63;
    if constexpr (j64 == 0) {
        return v;
    } else if constexpr (j64 == 16) {
        return rot<T, INV>(v);
    } else if constexpr (j64 == 32) {
        return mk<T>(-v.x, -v.y);
    } else if constexpr (j64 == 48) {
        return rot<T, !INV>(v);
    } else if constexpr (j64 == 8) {
        constexpr T h = T(0.70710678118654752440);
        return INV ? mk<T>(h * (v.x - v.y), h * (v.x + v.y)) : mk<T>(h * (v.x + v.y), h * (v.y - v.x));
    } else if constexpr (j64 == 24) {
        constexpr T h = T(0.70710678118654752440);
        return INV ? mk<T>(-h * (v.x + v.y), h * (v.x - v.y)) : mk<T>(h * (v.y - v.x), -h * (v.x + v.y));
    } else {
        constexpr T c = T(cos64(j64));
        constexpr T s = T(INV ? sin64(j64) : -sin64(j64));
        return mk<T>(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

template <typename T, int N, bool INV, int J = 0> struct DifLevel {
    static FB_HD void run(cx<T>* v)
    {
        cx<T> a = v[J], b = v[J + N / 2];
        v[J] = a + b;
        v[J + N / 2] = twiddle_const<T, N, J, INV>(a - b);
        if constexpr (J + 1 < N / 2) DifLevel<T, N, INV, J + 1>::run(v);
    }
};

// first level when v[N/2..N) is known to be zero (zero padded input)
template <typename T, int N, bool INV, int J = 0> struct DifLevelPruned {
    static FB_HD void run(cx<T>* v)
    {
        v[J + N / 2] = twiddle_const<T, N, J, INV>(v[J]);
        if constexpr (J + 1 < N / 2) DifLevelPruned<T, N, INV, J + 1>::run(v);
    }
};

template <typename T, int N, bool INV> struct RegFFT {
    static FB_HD void run(cx<T>* v)
    {
        DifLevel<T, N, INV>::run(v);
        RegFFT<T, N / 2, INV>::run(v);
        RegFFT<T, N / 2, INV>::run(v + N / 2);
    }
    // upper half of the input is zero
    static FB_HD void run_pruned(cx<T>* v)
    {
        DifLevelPruned<T, N, INV>::run(v);
        RegFFT<T, N / 2, INV>::run(v);
        RegFFT<T, N / 2, INV>::run(v + N / 2);
    }
};
template <typename T, bool INV> struct RegFFT<T, 1, INV> {
    static FB_HD void run(cx<T>*) {}
};

// (the mixed-radix register FFT GRegFFT / gpos lives in fb_gfft.cuh, shared with the shared-memory passes)
FB_HD constexpr bool is_pow2(int n) { return (n & (n - 1)) == 0; }

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// Packed variant for float on sm_100a: a complex number lives in an aligned register pair, so
// complex add / subtract / scale are ONE FADD2 / FMUL2 (f32x2) instead of two scalar
// instructions.  The FP32 lanes do the same work, but the issue slots halve -- the column and row
// kernels are issue bound (profiles/r1: not_selected + selected > 50 % of samples).  Multiplies
// by -i and the cross terms of general twiddles stay scalar (they mix .x and .y).  Forward only.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ cx<float> padd(cx<float> a, cx<float> b)
{
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return mk<float>(r.x, r.y);
}
__device__ __forceinline__ cx<float> psub(cx<float> a, cx<float> b)
{
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return mk<float>(r.x, r.y);
}
__device__ __forceinline__ cx<float> pscale(cx<float> a, float s)
{
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(s, s));
    return mk<float>(r.x, r.y);
}

template <int N, int J> __device__ __forceinline__ cx<float> ptwiddle_diff(cx<float> a, cx<float> b)
{
    constexpr int j64 = (J * (64 / N)) & 63;
    if constexpr (j64 == 0) {
        return psub(a, b);
    } else if constexpr (j64 == 16) {           // -i (a - b)
        return mk<float>(a.y - b.y, b.x - a.x);
    } else if constexpr (j64 == 32) {
        return psub(b, a);
    } else if constexpr (j64 == 48) {           // +i (a - b)
        return mk<float>(b.y - a.y, a.x - b.x);
    } else if constexpr (j64 == 8) {
        const cx<float> d = psub(a, b);
        return pscale(mk<float>(d.x + d.y, d.y - d.x), 0.70710678118654752440f);
    } else if constexpr (j64 == 24) {
        const cx<float> d = psub(a, b);
        return pscale(mk<float>(d.y - d.x, -d.x - d.y), 0.70710678118654752440f);
    } else {
        const cx<float> d = psub(a, b);
        constexpr float c = (float)cos64(j64);
        constexpr float s = (float)(-sin64(j64));
        return mk<float>(d.x * c - d.y * s, d.x * s + d.y * c);
    }
}

template <int N, int J = 0> struct PDifLevel {
    static __device__ __forceinline__ void run(cx<float>* v)
    {
        const cx<float> a = v[J], b = v[J + N / 2];
        v[J] = padd(a, b);
        v[J + N / 2] = ptwiddle_diff<N, J>(a, b);
        if constexpr (J + 1 < N / 2) PDifLevel<N, J + 1>::run(v);
    }
};

template <int N> struct PRegFFT {
    static __device__ __forceinline__ void run(cx<float>* v)
    {
        PDifLevel<N>::run(v);
        PRegFFT<N / 2>::run(v);
        PRegFFT<N / 2>::run(v + N / 2);
    }
};
template <> struct PRegFFT<1> {
    static __device__ __forceinline__ void run(cx<float>*) {}
};

// ---------------------------------------------------------------------------------------------
// Decimation-in-TIME variant of the packed radix-2 transform: same interface (natural order in,
// X[k] in v[brev<N>(k)]), but the twiddle multiplies the odd branch BEFORE the butterfly, so a
// general butterfly is x = e + w o (four FFMA with immediate constants), y = 2 e - x (two FFMA):
// 6 FP32 lane operations instead of the 8 of the DIF form (2 FADD2 + 2 FMUL + 2 FFMA); the
// sqrt(1/2) butterflies take 6 instead of 8.  388 instead of 456 lane operations per radix-32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ cx<float> pfma(float s, cx<float> a, cx<float> b)      // s * a + b, packed
{
    const float2 r = __ffma2_rn(make_float2(s, s), make_float2(a.x, a.y), make_float2(b.x, b.y));
    return mk<float>(r.x, r.y);
}

// e, o -> e + w o, e - w o with w = exp(-2 pi i J / N)
template <int N, int J> __device__ __forceinline__ void dit_bfly(cx<float>& e, cx<float>& o)
{
    constexpr int j64 = (J * (64 / N)) & 63;
    const cx<float> a = e, b = o;
    if constexpr (j64 == 0) {
        e = padd(a, b); o = psub(a, b);
    } else if constexpr (j64 == 16) {           // w = -i: w b = (b.y, -b.x)
        e = mk<float>(a.x + b.y, a.y - b.x); o = mk<float>(a.x - b.y, a.y + b.x);
    } else if constexpr (j64 == 8) {            // w = h (1 - i): w b = h (b.x + b.y, b.y - b.x)
        constexpr float h = 0.70710678118654752440f;
        const cx<float> sd = mk<float>(b.x + b.y, b.y - b.x);
        e = pfma(h, sd, a); o = pfma(-h, sd, a);
    } else if constexpr (j64 == 24) {           // w = -h (1 + i): w b = h (b.y - b.x, -(b.x + b.y))
        constexpr float h = 0.70710678118654752440f;
        const cx<float> ds = mk<float>(b.y - b.x, -b.x - b.y);
        e = pfma(h, ds, a); o = pfma(-h, ds, a);
    } else {
        constexpr float c = (float)cos64(j64);
        constexpr float s = (float)sin64(j64);
        // w b = (c b.x + s b.y, c b.y - s b.x)
        const float xr = fmaf(c, b.x, fmaf(s, b.y, a.x));
        const float xi = fmaf(c, b.y, fmaf(-s, b.x, a.y));
        e = mk<float>(xr, xi);
        o = mk<float>(fmaf(2.0f, a.x, -xr), fmaf(2.0f, a.y, -xi));
    }
}

template <int N, int S, int K = 0> struct PDitCombine {
    static __device__ __forceinline__ void run(cx<float>* v)
    {
        constexpr int pe = 2 * S * brev<N / 2>(K);
        dit_bfly<N, K>(v[pe], v[pe + S]);
        if constexpr (K + 1 < N / 2) PDitCombine<N, S, K + 1>::run(v);
    }
};

// transforms v[0], v[S], ..., v[(N - 1) S] in place; X[k] ends at v[S * brev<N>(k)].
// PR: the upper half of the (top-level) input is zero -- the size-2 leaves pair element i with i + N_top / 2,
// so they degenerate to copies
template <int N, int S = 1, bool PR = false> struct PDitFFT {
    static __device__ __forceinline__ void run(cx<float>* v)
    {
        if constexpr (N == 2 && PR) {
            v[S] = v[0];
        } else {
            PDitFFT<N / 2, 2 * S, PR>::run(v);
            PDitFFT<N / 2, 2 * S, PR>::run(v + S);
            PDitCombine<N, S>::run(v);
        }
    }
};
template <int S, bool PR> struct PDitFFT<1, S, PR> {
    static __device__ __forceinline__ void run(cx<float>*) {}
};

#ifndef FB_DIT
#define FB_DIT 1
#endif
constexpr bool kUseDit = FB_DIT != 0;

// front end of the warp FFTs: packed radix-2 code for powers of two, the mixed-radix code otherwise
template <int N> struct LaneFFT {
    static __device__ __forceinline__ void run(cx<float>* v)
    {
        if constexpr (is_pow2(N) && kUseDit) PDitFFT<N>::run(v);
        else if constexpr (is_pow2(N)) PRegFFT<N>::run(v);
        else GRegFFT<float, N, false>::run(v);
    }
    template <bool PRUNED> static __device__ __forceinline__ void run_first(cx<float>* v, bool pruned_now)
    {
        if constexpr (is_pow2(N) && kUseDit) {
            if (PRUNED && pruned_now) PDitFFT<N, 1, true>::run(v); else PDitFFT<N>::run(v);
        } else if constexpr (is_pow2(N)) {
            if (PRUNED && pruned_now) DifLevelPruned<float, N, false>::run(v); else PDifLevel<N>::run(v);
            PRegFFT<N / 2>::run(v);
            PRegFFT<N / 2>::run(v + N / 2);
        } else {
            if (PRUNED && pruned_now) GRegFFT<float, N, false>::run_pruned(v); else GRegFFT<float, N, false>::run(v);
        }
    }
};
#endif

}  // namespace fb
