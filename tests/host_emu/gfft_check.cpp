// Test infrastructure: host-side check of the compile-time mixed-radix register FFT (fb_gfft.cuh) against a
// direct DFT, of its pruned variant, of the composite-radix butterflies, and of the pass-radix planner.
#include <cmath>
#include <complex>
#include <cstdio>

#include "../../feabas_b200/csrc/fb_host_plan.h"

using namespace fb;
typedef std::complex<double> C;

template <typename T, int N, bool INV> double check()
{
    cx<T> v[N], w[N], w2[N];
    C x[N];
    for (int i = 0; i < N; ++i) {
        v[i] = mk<T>((T)std::sin(0.7 * i + 0.3), (T)std::cos(1.3 * i));
        x[i] = C(v[i].x, v[i].y);
    }
    GRegFFT<T, N, INV>::run(v);
    double err = 0, mx = 0;
    for (int k = 0; k < N; ++k) {
        C s = 0;
        for (int n = 0; n < N; ++n) s += x[n] * std::polar(1.0, (INV ? 2 : -2) * M_PI * n * k / N);
        err = std::max(err, std::abs(C(v[gpos<N>(k)].x, v[gpos<N>(k)].y) - s));
        mx = std::max(mx, std::abs(s));
    }
    if constexpr (N % 2 == 0) {                       // pruned first level == full transform of a half-empty input
        for (int i = 0; i < N; ++i) w[i] = w2[i] = i < N / 2 ? mk<T>((T)x[i].real(), (T)x[i].imag()) : mk<T>(T(0), T(0));
        GRegFFT<T, N, INV>::run(w);
        GRegFFT<T, N, INV>::run_pruned(w2);
        for (int i = 0; i < N; ++i) err = std::max(err, (double)std::hypot(w[i].x - w2[i].x, w[i].y - w2[i].y));
    }
    return err / mx;
}

template <typename T, int R, bool INV> double check_bfly()
{
    cx<T> v[R];
    C x[R];
    for (int i = 0; i < R; ++i) { v[i] = mk<T>((T)std::cos(0.9 * i), (T)std::sin(0.4 * i + 1)); x[i] = C(v[i].x, v[i].y); }
    Bfly<T, R, INV>::run(v);
    double err = 0, mx = 0;
    for (int k = 0; k < R; ++k) {
        C s = 0;
        for (int n = 0; n < R; ++n) s += x[n] * std::polar(1.0, (INV ? 2 : -2) * M_PI * n * k / R);
        err = std::max(err, std::abs(C(v[k].x, v[k].y) - s));
        mx = std::max(mx, std::abs(s));
    }
    return err / mx;
}

int main()
{
    double ef = 0, ed = 0;
#define CHK(N) ef = std::max(ef, check<float, N, false>()); ef = std::max(ef, check<float, N, true>()); \
               ed = std::max(ed, check<double, N, false>()); ed = std::max(ed, check<double, N, true>());
    CHK(2) CHK(3) CHK(4) CHK(5) CHK(6) CHK(8) CHK(9) CHK(10) CHK(12) CHK(15) CHK(16) CHK(20) CHK(24) CHK(30) CHK(40) CHK(48)
#define CHKB(R) ef = std::max(ef, check_bfly<float, R, false>()); ef = std::max(ef, check_bfly<float, R, true>()); \
                ed = std::max(ed, check_bfly<double, R, false>()); ed = std::max(ed, check_bfly<double, R, true>());
    CHKB(6) CHKB(9) CHKB(10) CHKB(12) CHKB(15) CHKB(16)
    int bad = 0;
    for (int n = 1; n <= 8192; ++n) {
        if (!is_5smooth(n)) continue;
        for (int wide = 0; wide < 2; ++wide) {
            auto r = wide ? radix_sequence(n) : radix_sequence_basic(n);
            long long p = 1;
            for (int x : r) { p *= x; if (x > (wide ? 16 : 8) || x < 2) ++bad; }
            if (p != n || (int)r.size() > kMaxPass) ++bad;
            auto pos = digit_positions(n, r);
            std::vector<char> seen(n, 0);
            for (int k = 0; k < n; ++k) { if (pos[k] < 0 || pos[k] >= n || seen[pos[k]]) ++bad; else seen[pos[k]] = 1; }
        }
    }
    auto r150 = radix_sequence(150), r135 = radix_sequence(135);
    printf("float %.3g double %.3g planner_bad %d passes150 %d passes135 %d\n", ef, ed, bad, (int)r150.size(), (int)r135.size());
    return (ef < 2e-6 && ed < 5e-14 && bad == 0 && r150.size() == 2 && r135.size() == 2) ? 0 : 1;
}
