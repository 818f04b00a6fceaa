// fb_xcorr.cuh -- kernel bodies of the FFT cross-correlation matcher.
//
// Replaces the arithmetic of feabas/matcher.py:22-135 (xcorr_fft): zero padded
// rfft2 of both stacks, conj / plain cross-power, irfft2, flat arg-max, 3x3
// quadratic sub-pixel fit, centre / wrap correction and the NONE / STD / MIRROR
// confidence measures.
//
// Two execution shapes share the phase functions below:
//   * staged  (large FFT grids): K1 rows forward -> K2 columns (forward, cross
//     power(s), inverse) -> K3 rows inverse + arg-max partials -> K4 finalize,
//     with half spectra in HBM workspaces F0/F1/G;
//   * fused   (both half spectra of a pair fit in one SM's shared memory): the
//     same phases run back to back in one CTA, nothing but the images is read
//     and nothing but 5 numbers per pair is written.
//
// The bodies are written against (bid, tid, nthr) and FB_SYNC() so that the
// host emulator in tests/host_emu can execute them unchanged on the CPU.
#pragma once
#include "fb_fft.cuh"

#if defined(__CUDA_ARCH__)
#define FB_SYNC() __syncthreads()
#else
namespace fb { void emu_sync(); }
#define FB_SYNC() ::fb::emu_sync()
#endif

namespace fb {

enum { CONF_NONE = 0, CONF_STD = 1, CONF_MIRROR = 2 };   // feabas/constant.py:39-41

struct Partial {          // one per (pair, K3 tile)
    double val;           // best correlation value in the tile
    double mir;           // max |mirror surface| in the tile
    double sum, sumsq;    // for FFT_CONF_STD
    int idx;              // flat index (y * nx + x) of the best value, lowest on ties
    int pad;
};

struct XcParams {
    const void* img0;     // [n][h0][w0]
    const void* img1;     // [n][h1][w1]
    int n;
    int h0, w0, h1, w1;
    int ny, nx, kp;
    Plan1D px, py;
    void* F0;             // cx<T> [n][h0][fpitch]   row spectra of img0 (natural k order)
    void* F1;             // cx<T> [n][h1][fpitch]
    void* G;              // cx<T> [n][ny][2*fpitch] P | Q after the column stage
    int fpitch;
    Partial* part;        // [n][nrt]
    int nrt;
    double* dx;           // [n] outputs (device pointers); peak / mir may be null
    double* dy;
    double* conf;
    double* peak;
    double* mir;
    int conf_mode;
    int subpixel;
    double scale;         // 1 / (ny * nx)
    double out_scale;     // K4: factor applied to the reported peak / mirror maxima (fast path: the surfaces
                          //     in G are unscaled, out_scale = scale; otherwise 1)
    int tl;               // lines per row tile
    int tc;               // columns per image per column tile
    int spitch;           // fused: row pitch of the resident spectra
    // optional extensions (generic staged path only; all null / 1 by default)
    const void* norm;     // T[ny][nx]: the correlation surface is divided by it before the peak search
                          //           (mask normalisation, matcher.py:71-81)
    const void* norm_m;   // T[ny][nx]: same for the mirror surface (matcher.py:119-124)
    void* surf;           // T[n][ny][nx] out: the (unnormalised) correlation surface
    void* surf_m;         // T[n][ny][nx] out: |mirror surface|
    int gt_layout;        // K4: != 0: G holds the fast path's conjugated surfaces, tiled
                          //     [ny / gt_layout][P|Q][kx][gt_layout] (gt_layout = rows per K3 tile), and
                          //     the partial's idx only identifies the ROW of the maximum
    int fin_narrow;       // K4: != 0: the three rows around the peak are recomputed one at a time (scratch of ONE
                          //     line of nx points instead of a 4-line tile; long lines, e.g. FFT 8192 or float64 4096)
};

template <typename T> struct Acc {
    T val; T mir; double sum, sumsq; int idx; int any;
};

template <typename T> FB_HD void acc_init(Acc<T>& a) { a.val = T(0); a.mir = T(0); a.sum = 0; a.sumsq = 0; a.idx = 0; a.any = 0; }

template <typename T> FB_HD void acc_take(Acc<T>& a, T v, int idx)
{
    // np.argmax semantics: largest value, lowest flat index among equals (matcher.py:82)
    if (!a.any || v > a.val || (v == a.val && idx < a.idx)) { a.val = v; a.idx = idx; a.any = 1; }
}

template <typename T> FB_HD void acc_merge(Acc<T>& a, const Acc<T>& b)
{
    if (b.any) acc_take(a, b.val, b.idx);
    a.mir = b.mir > a.mir ? b.mir : a.mir;
    a.sum += b.sum; a.sumsq += b.sumsq;
}

// Block-wide reduction; the result is broadcast to every thread.  `red` is
// shared scratch of at least 33 Acc<T> entries.
template <typename T>
FB_DEV Acc<T> block_reduce(Acc<T> a, Acc<T>* red, int tid, int nthr)
{
#if defined(__CUDA_ARCH__)
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        Acc<T> o;
        o.val = __shfl_xor_sync(full, a.val, off);
        o.mir = __shfl_xor_sync(full, a.mir, off);
        o.sum = __shfl_xor_sync(full, a.sum, off);
        o.sumsq = __shfl_xor_sync(full, a.sumsq, off);
        o.idx = __shfl_xor_sync(full, a.idx, off);
        o.any = __shfl_xor_sync(full, a.any, off);
        // keep the summation order identical in both partners
        Acc<T> lo = (tid & off) ? o : a, hi = (tid & off) ? a : o;
        acc_merge(lo, hi);
        a = lo;
    }
    const int warp = tid >> 5, nwarp = (nthr + 31) >> 5;
    if ((tid & 31) == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
        Acc<T> r = red[0];
        for (int w = 1; w < nwarp; ++w) acc_merge(r, red[w]);
        red[32] = r;
    }
    __syncthreads();
    Acc<T> r = red[32];
    __syncthreads();
    return r;
#else
    // emulator: nthr <= 32
    red[tid] = a;
    FB_SYNC();
    if (tid == 0) {
        Acc<T> r = red[0];
        for (int w = 1; w < nthr; ++w) acc_merge(r, red[w]);
        red[32] = r;
    }
    FB_SYNC();
    Acc<T> r = red[32];
    FB_SYNC();
    return r;
#endif
}

// MAXR: largest pass radix the plan may contain (8: basic inlined passes; 16: all radices, out of line)
// floor(i / d) for 0 <= i < 2^22 through the rounded reciprocal inv = 1.0f / d (exact in that range: the
// true quotient of i + 0.5 is at least 0.5 / d away from an integer, the float error is below that)
FB_HD int fdiv(int i, float inv) { return (int)(((float)i + 0.5f) * inv); }

template <typename T, bool INV, int MAXR = 8>
FB_DEV void fft_lines(const Plan1D& pl, cx<T>* s, int pitch, int nl, int tid, int nthr)
{
    const cx<T>* tw = reinterpret_cast<const cx<T>*>(pl.tw);
    if (!INV) {
        int L = pl.n;
        for (int p = 0; p < pl.npass; ++p) {
            if constexpr (MAXR > 8) fft_pass_any<T, false>(pl.radix[p], s, pitch, nl, pl.n, L, tw, tid, nthr);
            else fft_pass_basic<T, false>(pl.radix[p], s, pitch, nl, pl.n, L, tw, tid, nthr);
            L /= pl.radix[p];
            FB_SYNC();
        }
    } else {
        int L = 1;
        for (int p = pl.npass - 1; p >= 0; --p) {
            L *= pl.radix[p];
            if constexpr (MAXR > 8) fft_pass_any<T, true>(pl.radix[p], s, pitch, nl, pl.n, L, tw, tid, nthr);
            else fft_pass_basic<T, true>(pl.radix[p], s, pitch, nl, pl.n, L, tw, tid, nthr);
            FB_SYNC();
        }
    }
}

// ----------------------------------------------------------------------------
// Phase A: forward row transforms.  Line l of the tile carries image rows
// 2*(line0+l) (real part) and 2*(line0+l)+1 (imaginary part), zero extended to
// nx (rfft2(..., s=fftshp) semantics, matcher.py:63-64); one complex FFT, then
// the two half spectra are separated and written in natural k order to
// dst[row * dpitch + k].
// ----------------------------------------------------------------------------
template <typename T, typename TI, int MAXR = 8>
FB_DEV void rows_forward_tile(const XcParams& p, const TI* img, int H, int W, int line0, int nl,
                              cx<T>* dst, int dpitch, cx<T>* s, int pitch, int tid, int nthr)
{
    const int nx = p.nx;
    const float inv_nx = 1.0f / (float)nx;
    for (int idx = tid; idx < nx * nl; idx += nthr) {
        int l = fdiv(idx, inv_nx), x = idx - l * nx;
        int r0 = 2 * (line0 + l), r1 = r0 + 1;
        T a = T(0), b = T(0);
        if (x < W) {
            a = (T)FB_LDG(img + (size_t)r0 * W + x);
            if (r1 < H) b = (T)FB_LDG(img + (size_t)r1 * W + x);
        }
        s[(size_t)x * pitch + l] = mk<T>(a, b);
    }
    FB_SYNC();
    fft_lines<T, false, MAXR>(p.px, s, pitch, nl, tid, nthr);
    const int kp = p.kp;
    const int* pos = p.px.pos;
    const float inv_kp = 1.0f / (float)kp;
    for (int idx = tid; idx < kp * nl; idx += nthr) {
        int l = fdiv(idx, inv_kp), k = idx - l * kp;
        int km = k ? nx - k : 0;
        cx<T> zk = s[(size_t)FB_LDG(pos + k) * pitch + l];
        cx<T> zm = s[(size_t)FB_LDG(pos + km) * pitch + l];
        int r0 = 2 * (line0 + l), r1 = r0 + 1;
        dst[(size_t)r0 * dpitch + k] = mk<T>(T(0.5) * (zk.x + zm.x), T(0.5) * (zk.y - zm.y));
        if (r1 < H) dst[(size_t)r1 * dpitch + k] = mk<T>(T(0.5) * (zk.y + zm.y), T(0.5) * (zm.x - zk.x));
    }
    FB_SYNC();
}

// ----------------------------------------------------------------------------
// Phase B: column stage on a tile S[ny][pitch] whose first `half` lines are
// columns of F0 and next `half` lines the same columns of F1 (rows >= H zero).
// Forward both, P = conj(F0) F1 (matcher.py:65), Q = F0 F1 (matcher.py:114),
// scaled by 1/(ny nx), inverse both (only P unless mirror).
// ----------------------------------------------------------------------------
template <typename T, int MAXR = 8>
FB_DEV void cols_stage(const XcParams& p, cx<T>* S, int pitch, int half, bool mirror, int tid, int nthr)
{
    fft_lines<T, false, MAXR>(p.py, S, pitch, 2 * half, tid, nthr);
    const T sc = (T)p.scale;
    const float inv_half = 1.0f / (float)half;
    for (int idx = tid; idx < p.ny * half; idx += nthr) {
        int e = fdiv(idx, inv_half), c = idx - e * half;
        cx<T>* q = S + (size_t)e * pitch + c;
        cx<T> a = q[0], b = q[half];
        q[0] = cscale(cmulc(b, a), sc);
        if (mirror) q[half] = cscale(cmul(a, b), sc);
    }
    FB_SYNC();
    fft_lines<T, true, MAXR>(p.py, S, pitch, mirror ? 2 * half : half, tid, nthr);
}

// ----------------------------------------------------------------------------
// Phase C: inverse row transforms of a group of surface rows.  With the mirror
// term, line l = (P row y, Q row y): Z = P + iQ gives C[y] in the real part and
// the mirror surface in the imaginary part.  Without it, line l = P rows
// (y, y+1).  Hermitian extension, inverse FFT, then every thread scans its
// share of the tile into `acc`.
// ----------------------------------------------------------------------------
template <typename T>
FB_DEV void rows_inverse_fill(const XcParams& p, const cx<T>* X, const cx<T>* Y, cx<T>* s, int pitch, int l,
                              int tid_in_line, int stride, size_t ks = 1)
{
    const int nx = p.nx, kp = p.kp;
    const int* pos = p.px.pos;
    for (int k = tid_in_line; k < kp; k += stride) {
        const size_t off = (size_t)k * ks;
        cx<T> a = X[off];
        cx<T> b = Y ? Y[off] : mk<T>(T(0), T(0));
        if (p.gt_layout) { a.y = -a.y; b.y = -b.y; }
        if (k == 0 || 2 * k == nx) {
            s[(size_t)FB_LDG(pos + k) * pitch + l] = mk<T>(a.x, b.x);
        } else {
            s[(size_t)FB_LDG(pos + k) * pitch + l] = mk<T>(a.x - b.y, a.y + b.x);
            s[(size_t)FB_LDG(pos + nx - k) * pitch + l] = mk<T>(a.x + b.y, b.x - a.y);
        }
    }
}

template <typename T, int MAXR = 8>
FB_DEV void rows_inverse_tile(const XcParams& p, const cx<T>* Pb, const cx<T>* Qb, int rpitch, int row0, int nl,
                              bool mirror, Acc<T>& acc, cx<T>* s, int pitch, int tid, int nthr, int pair = 0)
{
    const int nx = p.nx, ny = p.ny, kp = p.kp;
    const float inv_kp = 1.0f / (float)kp;
    for (int idx = tid; idx < kp * nl; idx += nthr) {
        int l = fdiv(idx, inv_kp), k = idx - l * kp;
        const cx<T>* X; const cx<T>* Y;
        if (mirror) {
            int y = row0 + l;
            X = Pb + (size_t)y * rpitch; Y = Qb + (size_t)y * rpitch;
        } else {
            int y = row0 + 2 * l;
            X = Pb + (size_t)y * rpitch; Y = (y + 1 < ny) ? Pb + (size_t)(y + 1) * rpitch : nullptr;
        }
        rows_inverse_fill<T>(p, X, Y, s, pitch, l, k, kp);   // one k per call
    }
    FB_SYNC();
    fft_lines<T, true, MAXR>(p.px, s, pitch, nl, tid, nthr);
    const bool want_std = p.conf_mode == CONF_STD;
    const float inv_nl = 1.0f / (float)nl;
    const T* norm = reinterpret_cast<const T*>(p.norm);
    const T* norm_m = reinterpret_cast<const T*>(p.norm_m);
    T* surf = p.surf ? reinterpret_cast<T*>(p.surf) + (size_t)pair * ny * nx : nullptr;
    T* surf_m = p.surf_m ? reinterpret_cast<T*>(p.surf_m) + (size_t)pair * ny * nx : nullptr;
    for (int idx = tid; idx < nx * nl; idx += nthr) {
        int x = (int)(((float)idx + 0.5f) * inv_nl);
        int l = idx - x * nl;
        cx<T> v = s[(size_t)x * pitch + l];
        if (mirror) {
            int y = row0 + l;
            T c = v.x, m = v.y < T(0) ? -v.y : v.y;
            if (surf) surf[(size_t)y * nx + x] = c;
            if (surf_m) surf_m[(size_t)y * nx + x] = m;
            if (norm) c = c / norm[(size_t)y * nx + x];
            if (norm_m) m = m / norm_m[(size_t)y * nx + x];
            acc_take(acc, c, y * nx + x);
            acc.mir = m > acc.mir ? m : acc.mir;
            if (want_std) { acc.sum += (double)c; acc.sumsq += (double)c * (double)c; }
        } else {
            int y = row0 + 2 * l;
            T c = v.x;
            if (surf) surf[(size_t)y * nx + x] = c;
            if (norm) c = c / norm[(size_t)y * nx + x];
            acc_take(acc, c, y * nx + x);
            if (want_std) { acc.sum += (double)c; acc.sumsq += (double)c * (double)c; }
            if (y + 1 < ny) {
                T c1 = v.y;
                if (surf) surf[(size_t)(y + 1) * nx + x] = c1;
                if (norm) c1 = c1 / norm[(size_t)(y + 1) * nx + x];
                acc_take(acc, c1, (y + 1) * nx + x);
                if (want_std) { acc.sum += (double)c1; acc.sumsq += (double)c1 * (double)c1; }
            }
        }
    }
    FB_SYNC();
}

// ----------------------------------------------------------------------------
// Phase D: sub-pixel refinement (matcher.py:84-106), centre / wrap correction
// (:107-110) and confidence (:111-134) for one pair.  Recomputes the three
// surface rows around the peak from the P (and Q) rows.
// ----------------------------------------------------------------------------
template <typename T, int MAXR = 8>
FB_DEV void finalize_pair(const XcParams& p, int pair, const Acc<T>& best, const cx<T>* Pb, const cx<T>* Qb,
                          int rpitch, bool mirror, cx<T>* s, int tid, int nthr, size_t ks = 1)
{
    const int nx = p.nx, ny = p.ny, kp = p.kp;
    const int py = best.idx / nx;
    int px = best.idx - py * nx;
    const int pitch = p.fin_narrow ? 1 : 4;
    T* keep = reinterpret_cast<T*>(s + (size_t)nx * pitch);                 // narrow: c[line][x - 1, x, x + 1]
    Acc<T>* red = reinterpret_cast<Acc<T>*>(keep + 16);
    if ((p.subpixel || p.gt_layout) && !p.fin_narrow) {
        for (int idx = tid; idx < kp * 3; idx += nthr) {
            int l = idx / kp, k = idx - l * kp;
            int y = py - 1 + l;
            y = y < 0 ? y + ny : (y >= ny ? y - ny : y);
            // tiled fast-path layout: row y starts at (y / R) * 2 kp R + y % R, R = gt_layout
            const size_t ro = p.gt_layout ? (size_t)(y / p.gt_layout) * 2 * kp * p.gt_layout + (y % p.gt_layout)
                                          : (size_t)y * rpitch;
            rows_inverse_fill<T>(p, Pb + ro, mirror ? Qb + ro : nullptr, s, pitch, l, k, kp, ks);
        }
        FB_SYNC();
        fft_lines<T, true, MAXR>(p.px, s, pitch, 3, tid, nthr);
        if (p.norm) {                                        // sub-pixel fit runs on the normalised surface
            const T* norm = reinterpret_cast<const T*>(p.norm);
            for (int idx = tid; idx < nx * 3; idx += nthr) {
                int l = idx / nx, x = idx - l * nx;
                int y = py - 1 + l;
                y = y < 0 ? y + ny : (y >= ny ? y - ny : y);
                s[(size_t)x * pitch + l].x = s[(size_t)x * pitch + l].x / norm[(size_t)y * nx + x];
            }
            FB_SYNC();
        }
        if (p.gt_layout) {
            // locate the maximum inside row py (np.argmax: first occurrence)
            Acc<T> a; acc_init(a);
            for (int x = tid; x < nx; x += nthr) acc_take(a, s[(size_t)x * pitch + 1].x, x);
            a = block_reduce<T>(a, red, tid, nthr);
            px = a.idx;
        }
    } else if (p.subpixel || p.gt_layout) {
        // one line at a time, the peak's own row first (it yields px on the tiled layout)
        for (int step = 0; step < 3; ++step) {
            const int l = step == 0 ? 1 : (step == 1 ? 0 : 2);
            if (l != 1 && !p.subpixel) break;
            int y = py - 1 + l;
            y = y < 0 ? y + ny : (y >= ny ? y - ny : y);
            const size_t ro = p.gt_layout ? (size_t)(y / p.gt_layout) * 2 * kp * p.gt_layout + (y % p.gt_layout)
                                          : (size_t)y * rpitch;
            for (int k = tid; k < kp; k += nthr) rows_inverse_fill<T>(p, Pb + ro, mirror ? Qb + ro : nullptr, s, 1, 0, k, kp, ks);
            FB_SYNC();
            fft_lines<T, true, MAXR>(p.px, s, 1, 1, tid, nthr);
            if (p.norm) {
                const T* norm = reinterpret_cast<const T*>(p.norm);
                for (int x = tid; x < nx; x += nthr) s[x].x = s[x].x / norm[(size_t)y * nx + x];
                FB_SYNC();
            }
            if (l == 1 && p.gt_layout) {
                Acc<T> a; acc_init(a);
                for (int x = tid; x < nx; x += nthr) acc_take(a, s[x].x, x);
                a = block_reduce<T>(a, red, tid, nthr);
                px = a.idx;
            }
            if (tid == 0) {
                const int xm = px == 0 ? nx - 1 : px - 1, xp = px == nx - 1 ? 0 : px + 1;
                keep[l * 3 + 0] = s[xm].x; keep[l * 3 + 1] = s[px].x; keep[l * 3 + 2] = s[xp].x;
            }
            FB_SYNC();
        }
    }
    if (tid == 0) {
        float ox = 0.f, oy = 0.f;
        if (p.subpixel) {
            const int xm = px == 0 ? nx - 1 : px - 1, xp = px == nx - 1 ? 0 : px + 1;
#define FB_C(j, xi) (p.fin_narrow ? keep[((j) + 1) * 3 + ((xi) == px ? 1 : ((xi) == xm ? 0 : 2))] : s[(xi) * pitch + ((j) + 1)].x)
            T c00 = FB_C(0, px);
            T gx = (FB_C(0, xp) - FB_C(0, xm)) / T(2);
            T gy = (FB_C(1, px) - FB_C(-1, px)) / T(2);
            T hxx = FB_C(0, xm) + FB_C(0, xp) - T(2) * c00;
            T hyy = FB_C(1, px) + FB_C(-1, px) - T(2) * c00;
            T hxy = (FB_C(-1, xm) + FB_C(1, xp) - FB_C(-1, xp) - FB_C(1, xm)) / T(4);
#undef FB_C
            T det = hxx * hyy - hxy * hxy;
            if (det > T(0)) {
                ox = (float)(-(hyy / det) * gx - (-hxy / det) * gy);
                oy = (float)(-(-hxy / det) * gx - (hxx / det) * gy);
            }
            ox = ox < -0.5f ? -0.5f : (ox > 0.5f ? 0.5f : ox);
            oy = oy < -0.5f ? -0.5f : (oy > 0.5f ? 0.5f : oy);
        }
        double dx = (double)px + (double)ox + (double)(p.w0 - p.w1) / 2.0;
        double dy = (double)py + (double)oy + (double)(p.h0 - p.h1) / 2.0;
        dy -= rint(dy / ny) * ny;
        dx -= rint(dx / nx) * nx;
        double conf;
        if (p.conf_mode == CONF_NONE) {
            conf = 1.0;
        } else if (p.conf_mode == CONF_MIRROR) {
            conf = 0.0;
            if (best.val > T(0)) {
                T c = T(1) - best.mir / best.val;
                c = c < T(0) ? T(0) : (c > T(1) ? T(1) : c);
                conf = (double)(float)c;
            }
        } else {
            double cnt = (double)ny * (double)nx;
            double mean = best.sum / cnt;
            double var = best.sumsq / cnt - mean * mean;
            T sd = (T)sqrt(var > 0.0 ? var : 0.0);
            T r = best.val / sd;
            T base = T(1) - (T)exp((double)(-r));
            // the exponent np.prod(fftshp) is an int64 scalar, so numpy evaluates the power in float64
            double c = pow((double)base, cnt);
            if (c == c) c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
            conf = c;
        }
        p.dx[pair] = dx; p.dy[pair] = dy; p.conf[pair] = conf;
        if (p.peak) p.peak[pair] = (double)best.val * p.out_scale;
        if (p.mir) p.mir[pair] = (double)best.mir * p.out_scale;
    }
    FB_SYNC();
}

// ============================================================================
// kernel bodies
// ============================================================================
template <typename T> FB_HD int row_tiles(int h, int tl) { return ((h + 1) / 2 + tl - 1) / tl; }

// K1: grid = n * (tiles(img0) + tiles(img1))
template <typename T, typename TI>
FB_DEV void k1_rows_forward(const XcParams& p, int bid, int tid, int nthr, unsigned char* smem)
{
    const int t0 = row_tiles<T>(p.h0, p.tl), t1 = row_tiles<T>(p.h1, p.tl);
    const int pair = bid / (t0 + t1);
    int t = bid - pair * (t0 + t1);
    const bool second = t >= t0;
    if (second) t -= t0;
    const int H = second ? p.h1 : p.h0, W = second ? p.w1 : p.w0;
    const TI* img = reinterpret_cast<const TI*>(second ? p.img1 : p.img0) + (size_t)pair * H * W;
    cx<T>* dst = reinterpret_cast<cx<T>*>(second ? p.F1 : p.F0) + (size_t)pair * H * p.fpitch;
    const int lines = (H + 1) / 2;
    const int line0 = t * p.tl;
    const int nl = lines - line0 < p.tl ? lines - line0 : p.tl;
    rows_forward_tile<T, TI>(p, img, H, W, line0, nl, dst, p.fpitch, reinterpret_cast<cx<T>*>(smem), p.tl + 1, tid, nthr);
}

// K2: grid = n * ceil(kp / tc)
template <typename T>
FB_DEV void k2_columns(const XcParams& p, int bid, int tid, int nthr, unsigned char* smem)
{
    const int nct = (p.kp + p.tc - 1) / p.tc;
    const int pair = bid / nct, ct = bid - pair * nct;
    const int tc = p.tc, pitch = 2 * tc + 1, ny = p.ny;
    const bool mirror = p.conf_mode == CONF_MIRROR;
    cx<T>* S = reinterpret_cast<cx<T>*>(smem);
    const cx<T>* F0 = reinterpret_cast<const cx<T>*>(p.F0) + (size_t)pair * p.h0 * p.fpitch;
    const cx<T>* F1 = reinterpret_cast<const cx<T>*>(p.F1) + (size_t)pair * p.h1 * p.fpitch;
    const int w2 = 2 * tc;
    const float inv_w2 = 1.0f / (float)w2;
    for (int idx = tid; idx < ny * w2; idx += nthr) {
        int y = fdiv(idx, inv_w2), c2 = idx - y * w2;
        int second = c2 >= tc;
        int col = ct * tc + (second ? c2 - tc : c2);
        cx<T> v = mk<T>(T(0), T(0));
        if (col < p.kp && y < (second ? p.h1 : p.h0)) v = ldg((second ? F1 : F0) + (size_t)y * p.fpitch + col);
        S[(size_t)y * pitch + c2] = v;
    }
    FB_SYNC();
    cols_stage<T>(p, S, pitch, tc, mirror, tid, nthr);
    cx<T>* G = reinterpret_cast<cx<T>*>(p.G) + (size_t)pair * ny * 2 * p.fpitch;
    const int wout = mirror ? w2 : tc;
    const float inv_wout = 1.0f / (float)wout;
    for (int idx = tid; idx < ny * wout; idx += nthr) {
        int y = fdiv(idx, inv_wout), c2 = idx - y * wout;
        int second = c2 >= tc;
        int col = ct * tc + (second ? c2 - tc : c2);
        if (col < p.kp) G[(size_t)y * 2 * p.fpitch + (second ? p.fpitch : 0) + col] = S[(size_t)y * pitch + c2];
    }
}

// K3: grid = n * nrt, nrt = ceil(ny / rows_per_tile), rows_per_tile = mirror ? tl : 2 tl
template <typename T>
FB_DEV void k3_rows_inverse(const XcParams& p, int bid, int tid, int nthr, unsigned char* smem)
{
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int rpt = mirror ? p.tl : 2 * p.tl;
    const int pair = bid / p.nrt, rt = bid - pair * p.nrt;
    const int row0 = rt * rpt;
    int rows = p.ny - row0 < rpt ? p.ny - row0 : rpt;
    const int nl = mirror ? rows : (rows + 1) / 2;
    const int pitch = p.tl + 1;
    cx<T>* s = reinterpret_cast<cx<T>*>(smem);
    Acc<T>* red = reinterpret_cast<Acc<T>*>(smem + (size_t)p.nx * pitch * sizeof(cx<T>));
    const cx<T>* Pb = reinterpret_cast<const cx<T>*>(p.G) + (size_t)pair * p.ny * 2 * p.fpitch;
    Acc<T> acc; acc_init(acc);
    rows_inverse_tile<T>(p, Pb, Pb + p.fpitch, 2 * p.fpitch, row0, nl, mirror, acc, s, pitch, tid, nthr, pair);
    Acc<T> r = block_reduce<T>(acc, red, tid, nthr);
    if (tid == 0) {
        Partial& o = p.part[(size_t)pair * p.nrt + rt];
        o.val = (double)r.val; o.mir = (double)r.mir; o.sum = r.sum; o.sumsq = r.sumsq; o.idx = r.idx; o.pad = 0;
    }
}

// K4: grid = n
template <typename T>
FB_DEV void k4_finalize(const XcParams& p, int bid, int tid, int nthr, unsigned char* smem)
{
    const bool mirror = p.conf_mode == CONF_MIRROR;
    cx<T>* s = reinterpret_cast<cx<T>*>(smem);
    Acc<T>* red = reinterpret_cast<Acc<T>*>(smem + (size_t)p.nx * (p.fin_narrow ? 1 : 4) * sizeof(cx<T>) + 16 * sizeof(T));
    Acc<T> acc; acc_init(acc);
    for (int i = tid; i < p.nrt; i += nthr) {
        const Partial& q = p.part[(size_t)bid * p.nrt + i];
        Acc<T> b; b.val = (T)q.val; b.mir = (T)q.mir; b.sum = q.sum; b.sumsq = q.sumsq; b.idx = q.idx; b.any = 1;
        acc_merge(acc, b);
    }
    Acc<T> best = block_reduce<T>(acc, red, tid, nthr);
    if (p.gt_layout) {
        // fast path: conjugated, tiled G^T[pair][ny / R][P|Q][kx][R]
        const size_t plane = (size_t)p.kp * p.ny;
        const cx<T>* Pb = reinterpret_cast<const cx<T>*>(p.G) + (size_t)bid * 2 * plane;
        finalize_pair<T>(p, bid, best, Pb, Pb + (size_t)p.kp * p.gt_layout, 0, mirror, s, tid, nthr, (size_t)p.gt_layout);
        return;
    }
    const cx<T>* Pb = reinterpret_cast<const cx<T>*>(p.G) + (size_t)bid * p.ny * 2 * p.fpitch;
    finalize_pair<T>(p, bid, best, Pb, Pb + p.fpitch, 2 * p.fpitch, mirror, s, tid, nthr);
}

// Fused: grid = n, shared memory = S[ny][spitch] | scratch[nx][tl+1] | reduction scratch
template <typename T, typename TI>
FB_DEV void kf_fused(const XcParams& p, int bid, int tid, int nthr, unsigned char* smem)
{
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int ny = p.ny, nx = p.nx, kp = p.kp, sp = p.spitch, tl = p.tl, pitch = tl + 1;
    cx<T>* S = reinterpret_cast<cx<T>*>(smem);
    cx<T>* s = S + (size_t)ny * sp;
    Acc<T>* red = reinterpret_cast<Acc<T>*>(s + (size_t)nx * pitch);
    const int pair = bid;
#if defined(__CUDA_ARCH__)
    if (tid < 2) {                                  // both images of the pair -> L2 ahead of the row tiles that read them
        const int H = tid ? p.h1 : p.h0, W = tid ? p.w1 : p.w0;
        size_t a = reinterpret_cast<size_t>(reinterpret_cast<const TI*>(tid ? p.img1 : p.img0) + (size_t)pair * H * W);
        size_t e = (a + (size_t)H * W * sizeof(TI)) & ~(size_t)15;
        a = (a + 15) & ~(size_t)15;
        if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
    }
#endif
    const float inv_kp = 1.0f / (float)kp;
    for (int second = 0; second < 2; ++second) {
        const int H = second ? p.h1 : p.h0, W = second ? p.w1 : p.w0;
        const TI* img = reinterpret_cast<const TI*>(second ? p.img1 : p.img0) + (size_t)pair * H * W;
        cx<T>* dst = S + (second ? kp : 0);
        const int lines = (H + 1) / 2;
        for (int line0 = 0; line0 < lines; line0 += tl) {
            int nl = lines - line0 < tl ? lines - line0 : tl;
            rows_forward_tile<T, TI, 16>(p, img, H, W, line0, nl, dst, sp, s, pitch, tid, nthr);
        }
        for (int idx = tid; idx < (ny - H) * kp; idx += nthr) {
            int y = fdiv(idx, inv_kp), k = idx - y * kp;
            dst[(size_t)(H + y) * sp + k] = mk<T>(T(0), T(0));
        }
    }
    FB_SYNC();
    cols_stage<T, 16>(p, S, sp, kp, mirror, tid, nthr);
    Acc<T> acc; acc_init(acc);
    const int rpt = mirror ? tl : 2 * tl;
    for (int row0 = 0; row0 < ny; row0 += rpt) {
        int rows = ny - row0 < rpt ? ny - row0 : rpt;
        int nl = mirror ? rows : (rows + 1) / 2;
        rows_inverse_tile<T, 16>(p, S, S + kp, sp, row0, nl, mirror, acc, s, pitch, tid, nthr, pair);
    }
    Acc<T> best = block_reduce<T>(acc, red, tid, nthr);
    finalize_pair<T, 16>(p, pair, best, S, S + kp, sp, mirror, s, tid, nthr);
}

}  // namespace fb
