// fb_xcorr_fast.cuh -- register-resident fast path for power-of-two FFT grids (float32 compute).
//
// A line of N = E * T points is transformed by T lanes of a warp holding E points each
// (T = 16: two lines per warp, T = 32: one): radix-E in registers over the stride-T
// elements, twiddle by w_N^(k1 t), one transpose through a private shared-memory region,
// radix-T in registers.  Input and output are both in natural order, lane-contiguous, so
// global loads / stores of whole lines are coalesced without staging.
//
//   n = n1 T + t,  k = k1 + E k2:
//   A[k1][t]   = sum_n1 x[n1 T + t] w_E^(n1 k1)           (stage A, lane t)
//   A[k1][t]  *= w_N^(k1 t)
//   X[k1+E k2] = sum_t  A[k1][t] w_T^(t k2)               (stage B, lane k1 mod T)
//
// The three stages keep the HBM intermediates transposed so that every kernel reads and
// writes whole lines:  K1 rows -> FT[img][pair][kx][y];  K2 columns: FT -> GT[pair][P|Q][kx][y];
// K3 rows: GT -> per-row arg-max partials.  Same arithmetic as the generic path
// (feabas/matcher.py:63-68,82,114-125).  CUDA only (warp shuffles, cp.async).
#pragma once
#include "fb_regfft.cuh"
#include "fb_xcorr.cuh"

namespace fb {

template <int E, int T> struct WarpFFT {
    static_assert(T == 16 || T == 32, "lanes per line");
    static_assert(E % T == 0 && E / T <= 2, "E/T stage-B transforms per lane");
    static constexpr int N = E * T;
    static constexpr int LPW = 32 / T;        // lines per warp
    static constexpr int M = E / T;           // stage-B transforms per lane
    static constexpr int RS = N + E + 1;      // minimum slots per line region
    // smallest region stride >= RS with stride % 16 == m: makes accesses by (line-minor, slot-major)
    // thread groups of 16/m lines conflict free
    static constexpr __host__ __device__ int stride_mod16(int m) { return RS + ((m - RS % 16) + 16) % 16; }

    // k (natural index) of register j after run(): see out_index()
    static __device__ __forceinline__ int out_k(int t, int j) { return t + T * (j % M) + E * (j / M); }
    // register that holds output j (j-th in increasing k for this lane)
    static constexpr __host__ __device__ int out_reg(int j) { return (j % M) * T + brev<T>(j / M); }

    // v: E registers; in: v[n1] = x[n1 T + t]; out: X[out_k(t, j)] = v[out_reg(j)].
    // Forward transform only: inverse transforms are taken as conj(fft(conj(.))) with the
    // conjugations folded into the neighbouring point-wise steps, so that every kernel
    // runs ONE butterfly body (instruction-cache footprint).
    template <bool PRUNED>
    static __device__ __forceinline__ void run(cx<float>* v, cx<float>* region, const cx<float>* tw, int t,
                                               bool pruned_now = true)
    {
        // first radix-2 level: skipped arithmetic when the upper half of the input is zero; the
        // remaining levels are one shared body
        if (PRUNED && pruned_now) DifLevelPruned<float, E, false>::run(v); else DifLevel<float, E, false>::run(v);
        RegFFT<float, E / 2, false>::run(v);
        RegFFT<float, E / 2, false>::run(v + E / 2);
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) {
            cx<float> a = v[brev<E>(k1)];
            if (k1) a = cmul(a, tw[k1 * T + t]);
            region[k1 * (T + 1) + t] = a;
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; ++m) {
#pragma unroll
            for (int n2 = 0; n2 < T; ++n2) v[m * T + n2] = region[(t + T * m) * (T + 1) + n2];
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; ++m) RegFFT<float, T, false>::run(v + m * T);
    }
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct FastParams {
    XcParams x;
    const cx<float>* twx;   // [EX][TX] w_nx^(k1 t)
    const cx<float>* twy;   // [EY][TY]
    cx<float>* FT0;         // [n][kp][hp0]
    cx<float>* FT1;         // [n][kp][hp1]
    cx<float>* GT;          // [n][2][kp][ny]: conjugate of the column-stage output, y contiguous
    int hp0, hp1;
};

// ---------------------------------------------------------------------------------------------
// K1: forward row transforms, two image rows per complex line, transposed half-spectrum out.
// grid-stride over (pair, image, tile of TR = 2*LPW*NW rows).  NW warps per CTA.
// ---------------------------------------------------------------------------------------------
template <int E, int T, int NW, typename TI, bool PRUNED>
__device__ void kfast_rows_forward(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPW = W::LPW, LPC = LPW * NW, TR = 2 * LPC, NT = 32 * NW;
    constexpr int RS = W::stride_mod16(LPC >= 16 ? 1 : 16 / LPC);
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = regions + LPC * RS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = lane % T, lw = lane / T;                 // lane within line, line within warp
    for (int i = tid; i < N; i += NT) tw[i] = fp.twx[i];
    __syncthreads();
    const int tiles0 = fp.hp0 / TR, tiles1 = fp.hp1 / TR, tpp = tiles0 + tiles1;
    const int kp = p.kp;
    for (int work = blockIdx.x; work < p.n * tpp; work += gridDim.x) {
        const int pair = work / tpp;
        int tile = work - pair * tpp;
        const bool second = tile >= tiles0;
        if (second) tile -= tiles0;
        const int H = second ? p.h1 : p.h0, Wd = second ? p.w1 : p.w0, hp = second ? fp.hp1 : fp.hp0;
        const TI* img = reinterpret_cast<const TI*>(second ? p.img1 : p.img0) + (size_t)pair * H * Wd;
        cx<float>* FT = (second ? fp.FT1 : fp.FT0) + (size_t)pair * kp * hp;
        const int row0 = tile * TR;
        const int line = warp * LPW + lw;                  // line within the tile
        const int rA = row0 + 2 * line, rB = rA + 1;
        cx<float>* region = regions + line * RS;
        cx<float> v[E];
        if (rA < H) {                                      // uniform per T-lane group
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) {
                const int xx = n1 * T + t;
                float a = 0.f, b = 0.f;
                if ((!PRUNED || n1 < E / 2) && xx < Wd) {
                    a = (float)__ldg(img + (size_t)rA * Wd + xx);
                    if (rB < H) b = (float)__ldg(img + (size_t)rB * Wd + xx);
                }
                v[n1] = mk<float>(a, b);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) v[n1] = mk<float>(0.f, 0.f);
        }
        W::template run<PRUNED>(v, region, tw, t);
#pragma unroll
        for (int j = 0; j < E; ++j) region[W::out_k(t, j)] = v[W::out_reg(j)];
        __syncthreads();
        // separation + transposed store: thread -> (line l minor, k major); both rows of a line are
        // adjacent in FT, so one 16-byte store writes the pair (A[k], B[k])
        {
            constexpr int KSTEP = NT / LPC;
            const int l = tid % LPC;
            const cx<float>* reg = regions + l * RS;
            float4* dst = reinterpret_cast<float4*>(FT + row0 + 2 * l);
            for (int k = tid / LPC; k < kp; k += KSTEP) {
                const cx<float> zk = reg[k], zm = reg[k ? N - k : 0];
                float4 o;
                o.x = 0.5f * (zk.x + zm.x); o.y = 0.5f * (zk.y - zm.y);      // row 2l   : (Z[k] + conj Z[N-k]) / 2
                o.z = 0.5f * (zk.y + zm.y); o.w = 0.5f * (zm.x - zk.x);      // row 2l+1 : (Z[k] - conj Z[N-k]) / 2i
                dst[(size_t)k * (hp / 2)] = o;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K2: column stage.  Warp w < NW/2 transforms the F0 column(s), warp w + NW/2 the same F1
// column(s); they swap spectra through shared memory, form conj(P) = F0 conj(F1) and
// conj(Q) = conj(F0) conj(F1), and transform again (forward transform of the conjugate =
// conjugate of the inverse transform; consumers undo the conjugation).  The two transforms of a
// column run through ONE copy of the butterfly code (rolled two-phase loop).
// grid-stride over (pair, column group).
// ---------------------------------------------------------------------------------------------
template <int E, int T, int NW, bool PRUNED0>
__device__ void kfast_columns(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPW = W::LPW, RS = W::RS, CPG = LPW * (NW / 2), NT = 32 * NW;   // CPG columns per CTA
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = regions + NW * LPW * RS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = lane % T, lw = lane / T;
    const bool roleB = warp >= NW / 2;
    const int pw = roleB ? warp - NW / 2 : warp;           // pair-of-warps index
    for (int i = tid; i < N; i += NT) tw[i] = fp.twy[i];
    __syncthreads();
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int kp = p.kp, groups = (kp + CPG - 1) / CPG;
    const float sc = (float)p.scale;
    const float sgn = roleB ? -sc : sc;
    cx<float>* mine = regions + (warp * LPW + lw) * RS;
    cx<float>* other = regions + ((roleB ? pw : pw + NW / 2) * LPW + lw) * RS;
    const int hp = roleB ? fp.hp1 : fp.hp0;
    const bool second_phase = !roleB || mirror;
    for (int work = blockIdx.x; work < p.n * groups; work += gridDim.x) {
        const int pair = work / groups, grp = work - pair * groups;
        const int col = grp * CPG + pw * LPW + lw;
        const bool live = col < kp;
        const cx<float>* src = (roleB ? fp.FT1 : fp.FT0) + ((size_t)pair * kp + (live ? col : 0)) * hp;
        cx<float> v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int y = n1 * T + t;
            cx<float> a = mk<float>(0.f, 0.f);
            if ((!PRUNED0 || n1 < E / 2) && y < hp) a = ldg(src + y);
            v[n1] = a;
        }
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            W::template run<PRUNED0>(v, mine, tw, t, phase == 0);
            if (phase == 0) {
#pragma unroll
                for (int j = 0; j < E; ++j) mine[W::out_k(t, j)] = v[W::out_reg(j)];
                asm volatile("bar.sync %0, 64;" ::"r"(pw + 1) : "memory");
                // the output of lane t, k = t + T j, is exactly the input element n1 = j of the
                // next transform: a register permutation, no exchange needed
                cx<float> u[E];
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    const cx<float> o = other[W::out_k(t, j)], m = v[W::out_reg(j)];
                    u[j] = cmulc(mk<float>(m.x * sc, m.y * sgn), o);
                }
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = u[j];
                asm volatile("bar.sync %0, 64;" ::"r"(pw + 1) : "memory");
                if (!second_phase) break;
            } else if (live) {
                cx<float>* dst = fp.GT + (((size_t)pair * 2 + (roleB ? 1 : 0)) * kp + col) * N;
#pragma unroll
                for (int j = 0; j < E; ++j) dst[W::out_k(t, j)] = v[W::out_reg(j)];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: inverse row transforms + per-line maxima, row-batched.  A line is (P row y, Q row y) with
// the mirror term, else (P row y, P row y+1).  The CTA (256 threads) owns R = 256 / T consecutive
// lines.  Stage A: thread (t, r) -- r minor, so a lane group reads R consecutive y of one GT
// column, contiguous -- assembles conj(Z)[n1 T + t] of line r from the (conjugated) P / Q
// columns by Hermitian extension, radix-E in registers, twiddles, writes X[k1][t][r].
// Stage B: thread (k1, r) reads X[k1][.][r], radix-T, reduces maxima per line.
// The forward transform of conj(Z) is conj(surface): C = Re, mirror = -Im (|.| only), second
// row = -Im.  Only per-line maxima are produced; the finalize kernel locates x inside the
// winning row (np.argmax order: lowest row, then lowest x).
// ---------------------------------------------------------------------------------------------
template <int E, int T>
__device__ void kfast_rows_inverse(const FastParams& fp, unsigned char* smem)
{
    constexpr int N = E * T, R = 256 / T, M = E / T;
    constexpr int XS = T * R + (R == 8 ? 8 : 0);              // k1 stride of the exchange tile
    const XcParams& p = fp.x;
    cx<float>* X = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = X + E * XS;
    float* red = reinterpret_cast<float*>(tw + N);            // [8 warps][R][2] floats + doubles after
    double* redd = reinterpret_cast<double*>(red + 8 * R * 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid % R, tq = tid / R;                      // stage A: t = tq; stage B: k1 = tq + T m
    for (int i = tid; i < N; i += 256) tw[i] = fp.twx[i];
    const bool mirror = p.conf_mode == CONF_MIRROR, want_std = p.conf_mode == CONF_STD;
    const int ny = p.ny, kp = p.kp;
    const int lines_pp = mirror ? ny : (ny + 1) / 2;          // lines per pair
    const int tiles = (lines_pp + R - 1) / R;
    const size_t plane = (size_t)kp * ny;
    for (int work = blockIdx.x; work < p.n * tiles; work += gridDim.x) {
        const int pair = work / tiles, tile = work - pair * tiles;
        const int gl = tile * R + r;                           // line index inside the pair
        const bool live = gl < lines_pp;
        const int y0 = mirror ? gl : 2 * gl;
        const bool have2 = live && (mirror || (y0 + 1 < ny));
        const float h1 = live ? 1.f : 0.f, h2 = have2 ? 1.f : 0.f;
        const cx<float>* A = fp.GT + (size_t)pair * 2 * plane + (live ? y0 : 0);
        const cx<float>* B = have2 ? (mirror ? A + plane : A + 1) : A;
        __syncthreads();                                       // X free (and tw visible)
        {
            const int t = tq;
            cx<float> v[E];
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) {
                const int k = n1 * T + t;
                cx<float> z;
                if (n1 < E / 2 || (n1 == E / 2 && t == 0)) {
                    cx<float> a = ldg(A + (size_t)k * ny), b = ldg(B + (size_t)k * ny);
                    a = mk<float>(a.x * h1, a.y * h1); b = mk<float>(b.x * h2, b.y * h2);
                    // conj(P + iQ) with stored a = conj(P), b = conj(Q):  a - i b
                    z = (k == 0 || 2 * k == N) ? mk<float>(a.x, -b.x) : mk<float>(a.x + b.y, a.y - b.x);
                } else {
                    const int m = N - k;
                    cx<float> a = ldg(A + (size_t)m * ny), b = ldg(B + (size_t)m * ny);
                    a = mk<float>(a.x * h1, a.y * h1); b = mk<float>(b.x * h2, b.y * h2);
                    // conj(conj(P) + i conj(Q)) = P - i Q = conj(a) - i conj(b)
                    z = mk<float>(a.x - b.y, -a.y - b.x);
                }
                v[n1] = z;
            }
            RegFFT<float, E, false>::run(v);
#pragma unroll
            for (int k1 = 0; k1 < E; ++k1) {
                cx<float> a = v[brev<E>(k1)];
                if (k1) a = cmul(a, tw[k1 * T + t]);
                X[k1 * XS + t * R + r] = a;
            }
        }
        __syncthreads();
        float best, second;
        double sum = 0.0, sumsq = 0.0;
        {
            cx<float> u[E];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int k1 = tq + T * m;
#pragma unroll
                for (int n2 = 0; n2 < T; ++n2) u[m * T + n2] = X[k1 * XS + n2 * R + r];
            }
#pragma unroll
            for (int m = 0; m < M; ++m) RegFFT<float, T, false>::run(u + m * T);
            best = u[0].x; second = mirror ? fabsf(u[0].y) : -u[0].y;
#pragma unroll
            for (int j = 1; j < E; ++j) {
                best = fmaxf(best, u[j].x);
                second = fmaxf(second, mirror ? fabsf(u[j].y) : -u[j].y);
            }
            if (want_std) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    sum += (double)u[j].x; sumsq += (double)u[j].x * (double)u[j].x;
                    if (have2 && !mirror) { sum -= (double)u[j].y; sumsq += (double)u[j].y * (double)u[j].y; }
                }
            }
        }
        // lanes with equal r inside the warp, then the 8 warps through shared memory
#pragma unroll
        for (int off = R; off < 32; off <<= 1) {
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
            second = fmaxf(second, __shfl_xor_sync(0xffffffffu, second, off));
            if (want_std) {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
                sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
            }
        }
        if ((tid & 31) < R) {
            red[(warp * R + r) * 2] = best; red[(warp * R + r) * 2 + 1] = second;
            if (want_std) { redd[(warp * R + r) * 2] = sum; redd[(warp * R + r) * 2 + 1] = sumsq; }
        }
        __syncthreads();
        if (tid < R && live) {
            for (int w = 1; w < 8; ++w) {
                best = fmaxf(best, red[(w * R + r) * 2]); second = fmaxf(second, red[(w * R + r) * 2 + 1]);
                if (want_std) { sum += redd[(w * R + r) * 2]; sumsq += redd[(w * R + r) * 2 + 1]; }
            }
            // idx = first flat index of the row that holds the maximum; x is resolved by the finalize kernel
            int row = y0;
            float mir = 0.f;
            if (mirror) mir = second;
            else if (have2 && second > best) { best = second; row += 1; }
            Partial& o = p.part[(size_t)pair * p.nrt + gl];
            o.val = (double)best; o.mir = (double)mir; o.sum = sum; o.sumsq = sumsq; o.idx = row * N; o.pad = 0;
        }
    }
}

}  // namespace fb
