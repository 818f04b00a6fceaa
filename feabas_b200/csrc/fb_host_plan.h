// fb_host_plan.h -- host-side planning shared by the CUDA library and the host emulator:
// radix factorisation, digit-reversal tables, twiddle tables, tile / shared-memory sizing.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "fb_fft.cuh"

namespace fb {

// 5-smooth check (the reference only ever produces 2^a 3^b 5^c sizes:
// scipy.fftpack.next_fast_len at feabas/matcher.py:60,62).
inline bool is_5smooth(int n)
{
    if (n < 1) return false;
    for (int p : {2, 3, 5}) while (n % p == 0) n /= p;
    return n == 1;
}

// Basic DIF radix sequence (staged kernels): 8s, then 4/2 for the rest of the power of two, then 5s and 3s.
inline std::vector<int> radix_sequence_basic(int n)
{
    std::vector<int> r;
    int a = 0;
    while (n % 2 == 0) { n /= 2; ++a; }
    while (a >= 3 && a != 4) { r.push_back(8); a -= 3; }
    if (a == 4) { r.push_back(4); r.push_back(4); a = 0; }
    if (a == 2) r.push_back(4);
    if (a == 1) r.push_back(2);
    while (n % 5 == 0) { r.push_back(5); n /= 5; }
    while (n % 3 == 0) { r.push_back(3); n /= 3; }
    return r;   // n == 1 expected; caller checks is_5smooth first
}

// Wide DIF radix sequence (fused kernel) over the pass radices {2,3,4,5,6,8,9,10,12,15,16}: the fewest passes, and among
// those the most balanced factorisation (smallest largest radix, then smallest sum), largest radix first.
inline void radix_search(int n, int max_r, std::vector<int>& cur, std::vector<int>& best)
{
    static const int kR[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};
    if (n == 1) {
        auto key = [](const std::vector<int>& v) {
            long long mx = 0, sum = 0;
            for (int r : v) { mx = r > mx ? r : mx; sum += r; }
            return ((long long)v.size() << 40) | (mx << 20) | sum;
        };
        if (best.empty() || key(cur) < key(best)) best = cur;
        return;
    }
    if (!best.empty() && cur.size() + 1 > best.size()) return;
    for (int r : kR) {
        if (r > max_r || n % r) continue;
        cur.push_back(r);
        radix_search(n / r, r, cur, best);
        cur.pop_back();
    }
}
inline std::vector<int> radix_sequence(int n, int max_radix = 16)
{
    std::vector<int> cur, best;
    if (n == 1) return best;
    radix_search(n, max_radix, cur, best);
    return best;   // empty if n is not 5-smooth; caller checks is_5smooth first
}

// pos[k]: where natural frequency k sits after the DIF passes.
// pos_N(k) = (k % R1) * (N / R1) + pos_{N/R1}(k / R1)
inline std::vector<int> digit_positions(int n, const std::vector<int>& radix)
{
    std::vector<int> pos(n);
    for (int k = 0; k < n; ++k) {
        int kk = k, len = n, p = 0;
        for (int R : radix) {
            len /= R;
            p += (kk % R) * len;
            kk /= R;
        }
        pos[k] = p;
    }
    return pos;
}

template <typename T>
inline std::vector<cx<T>> twiddle_table(int n)
{
    std::vector<cx<T>> tw(n);
    for (int k = 0; k < n; ++k) {
        // exact quadrant reduction keeps the table symmetric and accurate
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        tw[k].x = (T)std::cos(a);
        tw[k].y = (T)std::sin(a);
    }
    return tw;
}

// ---------------------------------------------------------------------------
// Geometry of one xcorr problem class and the tiling the kernels use for it.
// ---------------------------------------------------------------------------
struct Geometry {
    int h0, w0, h1, w1;   // image sizes
    int ny, nx;           // FFT grid (reference fftshp, feabas/matcher.py:59-62)
    int kp;               // nx/2 + 1 half-spectrum columns
    int esize;            // bytes per complex element (8 for f32 compute, 16 for f64)
    bool mirror;          // second inverse transform needed (FFT_CONF_MIRROR)

    // derived by choose_tiles()
    bool fused;           // whole pair resident in one CTA's shared memory
    int tl_row;           // lines per row-stage tile (a line = 2 real rows, or a P/Q row pair)
    int tc_col;           // columns per image in the column-stage tile
    int fpitch;           // complex elements per row in the global spectrum workspaces
    int nthreads_row, nthreads_col, nthreads_fused;
    size_t smem_row, smem_col, smem_fused;
    int spitch;           // fused: complex elements per row of the resident spectra (both images)
    int tl_fused;         // fused: scratch lines
};

constexpr size_t kMaxSmem = 227 * 1024;

inline size_t row_tile_bytes(int nx, int tl, int esize) { return (size_t)nx * (tl + 1) * esize; }

inline bool choose_tiles(Geometry& g, size_t smem_budget_multi = 112 * 1024)
{
    g.kp = g.nx / 2 + 1;
    g.fpitch = (g.kp + 7) & ~7;
    // ---- fused: resident spectra S[ny][2*kp] + scratch [nx][tl+1] + reduction scratch
    g.spitch = 2 * g.kp;
    size_t sbytes = (size_t)g.ny * g.spitch * g.esize;
    g.fused = false;
    g.tl_fused = 0;
    for (int tl = 32; tl >= 4; tl >>= 1) {
        size_t tot = sbytes + row_tile_bytes(g.nx, tl, g.esize) + 4096;
        if (tot <= kMaxSmem) { g.fused = true; g.tl_fused = tl; g.smem_fused = tot; break; }
    }
    g.nthreads_fused = 512;
    // ---- staged: pick the largest tiles that keep two CTAs per SM when possible
    g.tl_row = 0;
    for (int pass = 0; pass < 2 && !g.tl_row; ++pass) {
        size_t budget = pass == 0 ? smem_budget_multi : kMaxSmem - 2048;
        for (int tl = 16; tl >= 1; tl >>= 1)
            if (row_tile_bytes(g.nx, tl, g.esize) + 2048 <= budget) { g.tl_row = tl; break; }
        if (pass == 0 && g.tl_row && g.tl_row < 8) g.tl_row = 0;   // prefer wider tiles over occupancy
    }
    g.tc_col = 0;
    for (int pass = 0; pass < 2 && !g.tc_col; ++pass) {
        size_t budget = pass == 0 ? smem_budget_multi : kMaxSmem - 2048;
        for (int tc = 8; tc >= 1; tc >>= 1)
            if ((size_t)g.ny * (2 * tc + 1) * g.esize + 2048 <= budget) { g.tc_col = tc; break; }
        if (pass == 0 && g.tc_col && g.tc_col < 8) g.tc_col = 0;
    }
    if (!g.tl_row || !g.tc_col) return g.fused;
    g.smem_row = row_tile_bytes(g.nx, g.tl_row, g.esize) + 2048;
    g.smem_col = (size_t)g.ny * (2 * g.tc_col + 1) * g.esize + 2048;
    g.nthreads_row = (g.nx * g.tl_row >= 8192) ? 512 : 256;
    g.nthreads_col = (g.ny * 2 * g.tc_col >= 8192) ? 512 : 256;
    return true;
}

}  // namespace fb
