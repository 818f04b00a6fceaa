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
    static constexpr int RS = N + E + 1;      // slots per line region (odd)

    // k (natural index) of register j after run(): see out_index()
    static __device__ __forceinline__ int out_k(int t, int j) { return t + T * (j % M) + E * (j / M); }
    // register that holds output j (j-th in increasing k for this lane)
    static constexpr __host__ __device__ int out_reg(int j) { return (j % M) * T + brev<T>(j / M); }

    // v: E registers; in: v[n1] = x[n1 T + t]; out: X[out_k(t, j)] = v[out_reg(j)]
    template <bool INV, bool PRUNED>
    static __device__ __forceinline__ void run(cx<float>* v, cx<float>* region, const cx<float>* tw, int t)
    {
        if (PRUNED) RegFFT<float, E, INV>::run_pruned(v); else RegFFT<float, E, INV>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) {
            cx<float> a = v[brev<E>(k1)];
            if (k1) {
                cx<float> w = tw[k1 * T + t];
                a = INV ? cmulc(a, w) : cmul(a, w);
            }
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
        for (int m = 0; m < M; ++m) RegFFT<float, T, INV>::run(v + m * T);
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
    cx<float>* GT;          // [n][2][kp][ny]
    int hp0, hp1;
};

constexpr int kFastWarps = 8;

// ---------------------------------------------------------------------------------------------
// K1: forward row transforms, two image rows per complex line, transposed half-spectrum out.
// grid-stride over (pair, image, tile of TR = 16*LPW rows).
// ---------------------------------------------------------------------------------------------
template <int E, int T, typename TI, bool PRUNED>
__device__ void kfast_rows_forward(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPW = W::LPW, RS = W::RS, TR = 2 * LPW * kFastWarps;
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = regions + kFastWarps * LPW * RS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = lane % T, lw = lane / T;                 // lane within line, line within warp
    for (int i = tid; i < N; i += blockDim.x) tw[i] = fp.twx[i];
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
        W::template run<false, PRUNED>(v, region, tw, t);
#pragma unroll
        for (int j = 0; j < E; ++j) region[W::out_k(t, j)] = v[W::out_reg(j)];
        __syncthreads();
        // separation + transposed store: thread -> (row r minor, k major)
        for (int idx = tid; idx < kp * TR; idx += blockDim.x) {
            const int r = idx % TR, k = idx / TR;
            const cx<float>* reg = regions + (r >> 1) * RS;
            const cx<float> zk = reg[k], zm = reg[k ? N - k : 0];
            cx<float> o = (r & 1) ? mk<float>(0.5f * (zk.y + zm.y), 0.5f * (zm.x - zk.x))
                                  : mk<float>(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
            FT[(size_t)k * hp + row0 + r] = o;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// K2: column stage.  Warp w < 4 transforms the F0 column(s), warp w + 4 the same F1 column(s);
// they swap spectra through shared memory, form P = conj(F0) F1 and Q = F0 F1, inverse
// transform and store the P / Q columns contiguously.  grid-stride over (pair, column group).
// ---------------------------------------------------------------------------------------------
template <int E, int T, bool PRUNED0, bool PRUNED1>
__device__ void kfast_columns(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPW = W::LPW, RS = W::RS, CPG = LPW * (kFastWarps / 2);   // columns per CTA
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = regions + kFastWarps * LPW * RS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = lane % T, lw = lane / T;
    const bool roleB = warp >= kFastWarps / 2;
    const int pw = roleB ? warp - kFastWarps / 2 : warp;   // pair-of-warps index
    for (int i = tid; i < N; i += blockDim.x) tw[i] = fp.twy[i];
    __syncthreads();
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int kp = p.kp, groups = (kp + CPG - 1) / CPG;
    const float sc = (float)p.scale;
    cx<float>* mine = regions + (warp * LPW + lw) * RS;
    cx<float>* other = regions + ((roleB ? pw : pw + kFastWarps / 2) * LPW + lw) * RS;
    const int hp = roleB ? fp.hp1 : fp.hp0;
    for (int work = blockIdx.x; work < p.n * groups; work += gridDim.x) {
        const int pair = work / groups, grp = work - pair * groups;
        const int col = grp * CPG + pw * LPW + lw;
        const bool live = col < kp;
        const cx<float>* src = (roleB ? fp.FT1 : fp.FT0) + ((size_t)pair * kp + (live ? col : 0)) * hp;
        cx<float> v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int y = n1 * T + t;
            const bool pruned = roleB ? PRUNED1 : PRUNED0;
            cx<float> a = mk<float>(0.f, 0.f);
            if ((!pruned || n1 < E / 2) && y < hp) a = ldg(src + y);
            v[n1] = a;
        }
        if (roleB) W::template run<false, PRUNED1>(v, mine, tw, t); else W::template run<false, PRUNED0>(v, mine, tw, t);
#pragma unroll
        for (int j = 0; j < E; ++j) mine[W::out_k(t, j)] = v[W::out_reg(j)];
        asm volatile("bar.sync %0, 64;" ::"r"(pw + 1) : "memory");
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const cx<float> o = other[W::out_k(t, j)], m = v[W::out_reg(j)];
            // role A: m = F0, o = F1 -> conj(F0) F1 ; role B: m = F1, o = F0 -> F0 F1
            v[W::out_reg(j)] = cscale(roleB ? cmul(o, m) : cmulc(o, m), sc);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(pw + 1) : "memory");
        if (!roleB || mirror) {
            // natural order in -> registers n1: value at y-frequency k = n1 T + t
            cx<float> u[E];
#pragma unroll
            for (int j = 0; j < E; ++j) mine[W::out_k(t, j)] = v[W::out_reg(j)];
            __syncwarp();
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) u[n1] = mine[n1 * T + t];
            __syncwarp();
            W::template run<true, false>(u, mine, tw, t);
            if (live) {
                cx<float>* dst = fp.GT + (((size_t)pair * 2 + (roleB ? 1 : 0)) * kp + col) * N;
#pragma unroll
                for (int j = 0; j < E; ++j) dst[W::out_k(t, j)] = u[W::out_reg(j)];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: inverse row transforms + per-line arg-max partials.  A line is (P row y, Q row y) with the
// mirror term, else (P row y, P row y+1).  grid-stride over (pair, tile of 8*LPW lines).
// ---------------------------------------------------------------------------------------------
template <int E, int T>
__device__ void kfast_rows_inverse(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPW = W::LPW, RS = W::RS, LPC = LPW * kFastWarps;   // lines per CTA
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    cx<float>* tw = regions + LPC * RS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = lane % T, lw = lane / T;
    for (int i = tid; i < N; i += blockDim.x) tw[i] = fp.twx[i];
    const bool mirror = p.conf_mode == CONF_MIRROR, want_std = p.conf_mode == CONF_STD;
    const int kp = p.kp, ny = p.ny;
    const int lines_pp = mirror ? ny : (ny + 1) / 2;          // lines per pair
    const int tiles = (lines_pp + LPC - 1) / LPC;
    const int line = warp * LPW + lw;
    cx<float>* region = regions + line * RS;
    for (int work = blockIdx.x; work < p.n * tiles; work += gridDim.x) {
        const int pair = work / tiles, tile = work - pair * tiles;
        const int line0 = tile * LPC;
        const cx<float>* GP = fp.GT + (size_t)pair * 2 * kp * ny;
        __syncthreads();                                       // previous tile fully consumed (and tw visible)
        if (mirror) {
            // slots [0,kp): P^T[kx][y], slots [kp,2kp): Q^T[kx][y]
            for (int idx = tid; idx < 2 * kp * LPC; idx += blockDim.x) {
                const int r = idx % LPC, c = idx / LPC;        // c in [0, 2kp)
                const int y = line0 + r;
                if (y < ny) cp_async8(regions + r * RS + c, GP + (size_t)c * ny + y);
            }
        } else {
            for (int idx = tid; idx < 2 * kp * LPC; idx += blockDim.x) {
                const int r2 = idx % (2 * LPC), kx = idx / (2 * LPC);
                const int y = 2 * line0 + r2;
                if (y < ny) cp_async8(regions + (r2 >> 1) * RS + (r2 & 1) * kp + kx, GP + (size_t)kx * ny + y);
            }
        }
        cp_async_wait_all();
        __syncthreads();
        const int gl = line0 + line;                           // global line index
        const bool live = gl < lines_pp;
        const bool have2 = mirror || (2 * gl + 1 < ny);        // second half of the line present
        cx<float> v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int k = n1 * T + t;
            cx<float> z;
            if (n1 < E / 2 || (n1 == E / 2 && t == 0)) {
                cx<float> a = region[k];
                cx<float> b = have2 ? region[kp + k] : mk<float>(0.f, 0.f);
                z = (k == 0 || 2 * k == N) ? mk<float>(a.x, b.x) : mk<float>(a.x - b.y, a.y + b.x);
            } else {
                cx<float> a = region[N - k];
                cx<float> b = have2 ? region[kp + N - k] : mk<float>(0.f, 0.f);
                z = mk<float>(a.x + b.y, b.x - a.y);
            }
            v[n1] = live ? z : mk<float>(0.f, 0.f);
        }
        __syncwarp();
        W::template run<true, false>(v, region, tw, t);
        // lane-local scan in increasing x
        float best = 0.f, mir = 0.f, best2 = 0.f;
        int bx = 0, bx2 = 0;
        double sum = 0.0, sumsq = 0.0;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const cx<float> c = v[W::out_reg(j)];
            const int x = W::out_k(t, j);
            if (j == 0 || c.x > best) { best = c.x; bx = x; }
            if (mirror) {
                mir = fmaxf(mir, fabsf(c.y));
            } else {
                if (j == 0 || c.y > best2) { best2 = c.y; bx2 = x; }
            }
            if (want_std) {
                sum += (double)c.x; sumsq += (double)c.x * (double)c.x;
                if (have2) { sum += (double)c.y; sumsq += (double)c.y * (double)c.y; }
            }
        }
        int idx1, idx2 = 0x7fffffff;
        if (mirror) {
            idx1 = gl * N + bx;
        } else {
            idx1 = 2 * gl * N + bx;
            if (have2) {
                idx2 = (2 * gl + 1) * N + bx2;
                if (best2 > best) { best = best2; idx1 = idx2; }     // row y+1 has the larger flat index
            }
        }
        // reduce across the T lanes of the line
#pragma unroll
        for (int off = T / 2; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, idx1, off);
            const float om = __shfl_xor_sync(0xffffffffu, mir, off);
            if (ov > best || (ov == best && oi < idx1)) { best = ov; idx1 = oi; }
            mir = fmaxf(mir, om);
            if (want_std) {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
                sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
            }
        }
        if (t == 0 && live) {
            Partial& o = p.part[(size_t)pair * p.nrt + gl];
            o.val = (double)best; o.mir = (double)mir; o.sum = sum; o.sumsq = sumsq; o.idx = idx1; o.pad = 0;
        }
    }
}

}  // namespace fb
