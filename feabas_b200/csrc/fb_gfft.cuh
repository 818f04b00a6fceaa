// fb_gfft.cuh -- mixed-radix (2, 3, 5) register FFT of a small 5-smooth length N with compile-time
// twiddles.  Used (a) as the composite-radix "butterfly" of the shared-memory passes in fb_fft.cuh
// (radix 6, 9, 10, 12, 15, 16: half as many passes over a tile as with radix <= 8 alone) and (b) as
// the per-lane transform of the register-resident fast path for line lengths with factors 3 / 5.
//
// Decimation in frequency: one radix-R level (R = 2 while N is even, then 3, then 5) followed by
// R transforms of length N / R on the contiguous sub-blocks.  X[k] ends up in v[gpos<N>(k)];
// for powers of two gpos is the bit reversal.  Included by fb_fft.cuh after the Bfly definitions.
#pragma once

namespace fb {

FB_HD constexpr int first_radix(int n) { return n % 2 == 0 ? 2 : (n % 3 == 0 ? 3 : (n % 5 == 0 ? 5 : n)); }

template <int N> FB_HD constexpr int gpos(int k)
{
    if constexpr (N == 1) {
        return 0;
    } else {
        constexpr int R = first_radix(N);
        return (k % R) * (N / R) + gpos<N / R>(k / R);
    }
}

// sin / cos of 2 pi m / n at compile time: octant reduction (exact at multiples of pi / 4), Taylor series inside
FB_HD constexpr double series_sin(double x)
{
    double term = x, sum = x;
    for (int i = 1; i < 12; ++i) { term *= -x * x / ((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
FB_HD constexpr double series_cos(double x)
{
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) { term *= -x * x / ((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos(2 pi m / n)
FB_HD constexpr double cos_frac(int m, int n)
{
    constexpr double kPi = 3.14159265358979323846264338327950288;
    m %= n;
    if (m < 0) m += n;
    if (2 * m > n) m = n - m;                     // cos is even around pi
    if (4 * m > n) return -cos_frac(n - 2 * m, 2 * n);      // cos(pi - y) = -cos(y)
    if (m == 0) return 1.0;
    if (4 * m == n) return 0.0;
    if (8 * m == n) return 0.70710678118654752440;
    if (8 * m > n) return series_sin(kPi * (n - 4 * m) / (2.0 * n));       // cos(x) = sin(pi / 2 - x)
    return series_cos(2.0 * kPi * m / n);
}
FB_HD constexpr double sin_frac(int m, int n) { return cos_frac(4 * m - n, 4 * n); }   // sin(x) = cos(x - pi / 2)

// v * exp(-/+ 2 pi i M / N), trivial rotations special-cased
template <typename T, int N, int M, bool INV> FB_HD cx<T> gtwiddle(cx<T> v)
{
    constexpr int m = ((M % N) + N) % N;
    if constexpr (m == 0) {
        return v;
    } else if constexpr (4 * m == N) {
        return rot<T, INV>(v);
    } else if constexpr (2 * m == N) {
        return mk<T>(-v.x, -v.y);
    } else if constexpr (4 * m == 3 * N) {
        return rot<T, !INV>(v);
    } else {
        constexpr T c = (T)cos_frac(m, N);
        constexpr T s = (T)(INV ? -sin_frac(m, N) : sin_frac(m, N));
        return mk<T>(v.x * c + v.y * s, v.y * c - v.x * s);
    }
}

template <typename T, int N, int R, int J, bool INV, int S = 1> struct GTwLevel {      // y[S] *= w_N^(J S), S = 1..R-1
    static FB_HD void run(cx<T>* y)
    {
        y[S] = gtwiddle<T, N, J * S, INV>(y[S]);
        if constexpr (S + 1 < R) GTwLevel<T, N, R, J, INV, S + 1>::run(y);
    }
};

template <typename T, int N, int R, bool INV, bool PRUNED, int J = 0> struct GDifLevel {
    static FB_HD void run(cx<T>* v)
    {
        constexpr int Q = N / R;
        cx<T> y[R];
#pragma unroll
        for (int q = 0; q < R; ++q) y[q] = v[J + q * Q];
        if constexpr (PRUNED && R == 2) {
            y[1] = y[0];                             // upper half of the input is zero
        } else {
            Bfly<T, R, INV>::run(y);
        }
        GTwLevel<T, N, R, J, INV>::run(y);
#pragma unroll
        for (int q = 0; q < R; ++q) v[J + q * Q] = y[q];
        if constexpr (J + 1 < Q) GDifLevel<T, N, R, INV, PRUNED, J + 1>::run(v);
    }
};

template <typename T, int N, bool INV, int B = 0> struct GSub {             // the R sub-transforms of length N / R
    static FB_HD void run(cx<T>* v);
};

template <typename T, int N, bool INV> struct GRegFFT {
    static FB_HD void run(cx<T>* v)
    {
        if constexpr (N > 1) {
            constexpr int R = first_radix(N);
            GDifLevel<T, N, R, INV, false>::run(v);
            GSub<T, N, INV>::run(v);
        }
    }
    // upper half of the input is zero (N even)
    static FB_HD void run_pruned(cx<T>* v)
    {
        static_assert(N % 2 == 0, "pruned first level is radix 2");
        GDifLevel<T, N, 2, INV, true>::run(v);
        GSub<T, N, INV>::run(v);
    }
};
template <typename T, int N, bool INV, int B> FB_HD void GSub<T, N, INV, B>::run(cx<T>* v)
{
    constexpr int R = first_radix(N);
    GRegFFT<T, N / R, INV>::run(v + B * (N / R));
    if constexpr (B + 1 < R) GSub<T, N, INV, B + 1>::run(v);
}

// composite radices of the shared-memory passes: natural order in, natural order out
template <typename T, int R, bool INV> struct BflyC {
    static FB_HD void run(cx<T>* v)
    {
        cx<T> u[R];
#pragma unroll
        for (int q = 0; q < R; ++q) u[q] = v[q];
        GRegFFT<T, R, INV>::run(u);
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = u[gpos<R>(t)];
    }
};
template <typename T, bool INV> struct Bfly<T, 6, INV> : BflyC<T, 6, INV> {};
template <typename T, bool INV> struct Bfly<T, 9, INV> : BflyC<T, 9, INV> {};
template <typename T, bool INV> struct Bfly<T, 10, INV> : BflyC<T, 10, INV> {};
template <typename T, bool INV> struct Bfly<T, 12, INV> : BflyC<T, 12, INV> {};
template <typename T, bool INV> struct Bfly<T, 15, INV> : BflyC<T, 15, INV> {};
template <typename T, bool INV> struct Bfly<T, 16, INV> : BflyC<T, 16, INV> {};

}  // namespace fb
