// fb_xcorr_fast.cuh -- register-resident fast path (float32 compute) for FFT grids whose line lengths
// are in the size table of fb_fast_groups.h: powers of two 256 .. 4096 and 5-smooth lengths with
// factors 3 / 5 (576, 288, 384, 768, 1152, 300, 200, 400, 800).
//
// A line of N = E * T points is transformed by T lanes holding E points each (T < 32: several lines per
// warp, lanes beyond LPW * T shadow the first ones; T = 32: one line per warp; T = 64: a line spans two
// warps): radix-E in registers over the stride-T elements, twiddle by w_N^(k1 t), one transpose
// through a private shared-memory region, radix-T in registers (E / T transforms per lane).  Input and
// output are both in natural order, lane-contiguous, so global loads / stores of whole lines are
// coalesced without staging.  The per-lane transforms are the packed decimation-in-time radix-2 code of
// fb_regfft.cuh for powers of two and the mixed-radix (2, 3, 5) code of fb_gfft.cuh otherwise.
//
//   n = n1 T + t,  k = k1 + E k2:
//   A[k1][t]   = sum_n1 x[n1 T + t] w_E^(n1 k1)           (stage A, lane t)
//   A[k1][t]  *= w_N^(k1 t)
//   X[k1+E k2] = sum_t  A[k1][t] w_T^(t k2)               (stage B, lane k1 mod T)
//
// The three stages keep the HBM intermediates transposed so that every kernel reads and
// writes whole lines:  K1 rows -> FT[img][pair][kx][y];  K2 columns: FT -> GT[pair][y / R][P|Q][kx][y % R];
// K3 rows: GT -> per-row arg-max partials.  Same arithmetic as the generic path
// (feabas/matcher.py:63-68,82,114-125).  CUDA only (warp shuffles, cp.async.bulk, TMA tensor stores).
#pragma once
#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched at run time, no libcuda link)
#include "fb_regfft.cuh"
#include "fb_xcorr.cuh"

namespace fb {

// The stage-A twiddles w_N^(k1 t) of a lane (fixed t, k1 = 0..E-1) held in REGISTERS for the life of
// the kernel: k1 = 8 a + b, w^(k1 t) = w^(8 a t) w^(b t), i.e. 7 + (E/8 - 1) complex values per lane
// instead of an E-entry shared-memory row per transform.  The kernels are bound by the LSU /
// shared-memory pipe (profiles/r1: l1tex 70-78 % busy, FP32 40 %): this trades 2 E wavefronts per
// transform for ~E extra complex multiplies.
template <int E> struct LaneTw {
    static constexpr int NA = E / 8;
    cx<float> b[8];      // b[j] = w^(j t), j = 1..7
    cx<float> a[NA];     // a[i] = w^(8 i t), i = 1..NA-1
    // tab: [E / 2][T][2] paired table (fb_xcorr.cu get_warp_table)
    template <int T> __device__ __forceinline__ void load(const cx<float>* tab, int t)
    {
#pragma unroll
        for (int j = 1; j < 8; ++j) b[j] = ldg(tab + ((j / 2) * T + t) * 2 + (j & 1));
#pragma unroll
        for (int i = 1; i < NA; ++i) a[i] = ldg(tab + ((8 * i / 2) * T + t) * 2);
    }
    __device__ __forceinline__ cx<float> apply(cx<float> z, int k1) const
    {
        const int ia = k1 >> 3, ib = k1 & 7;       // compile-time after unrolling
        if (ib) z = cmul(z, b[ib]);
        if (ia) z = cmul(z, a[ia]);
        return z;
    }
};

// The same twiddles read from a shared-memory copy of the table instead (paired rows, one LDS.128 per
// two outputs, fetched a batch ahead of use).  Cheaper in FP32 work, dearer in LSU wavefronts; the
// two variants measure within 1 % of each other on B200 (profiles/r1/sweeps.md), this one slightly ahead.
// TBMAX: most table rows fetched per batch (two batches are in flight: 8 TBMAX registers)
template <int E, int T, int TBMAX = 4> struct SmemTw {
    const float4* tw4;   // + t applied
    __device__ __forceinline__ void init(const cx<float>* smem_table, int t) { tw4 = reinterpret_cast<const float4*>(smem_table) + t; }
    // out(k1, value) stores the twiddled value of output k1
    template <typename F> __device__ __forceinline__ void apply_all(const cx<float>* v, F out) const
    {
        constexpr int H = E / 2, TB0 = H % 4 == 0 ? 4 : (H % 3 == 0 ? 3 : (H % 5 == 0 ? 5 : 1));
        constexpr int TB = TB0 <= TBMAX ? TB0 : (H % 2 == 0 && TBMAX >= 2 ? 2 : 1), NB = H / TB;
        float4 wq[2][TB];
#pragma unroll
        for (int i = 0; i < TB; ++i) wq[0][i] = tw4[i * T];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b + 1 < NB) {
#pragma unroll
                for (int i = 0; i < TB; ++i) wq[(b + 1) & 1][i] = tw4[((b + 1) * TB + i) * T];
            }
#pragma unroll
            for (int i = 0; i < TB; ++i) {
                const int k1 = 2 * (b * TB + i);
                const float4 w = wq[b & 1][i];
                cx<float> a0 = v[gpos<E>(k1)], a1 = v[gpos<E>(k1 + 1)];
                if (k1) a0 = cmul(a0, mk<float>(w.x, w.y));
                a1 = cmul(a1, mk<float>(w.z, w.w));
                out(k1, a0);
                out(k1 + 1, a1);
            }
        }
    }
};

constexpr bool kLaneTwiddles = false;     // true: LaneTw (registers), false: SmemTw (shared-memory table)
#ifndef FB_LANE_TW_K2
#define FB_LANE_TW_K2 0
#endif
// LANE: this kernel keeps the stage twiddles of a lane in registers (needs E % 8 == 0)
template <int E, int T, bool LANE = kLaneTwiddles, int TBMAX = 4> struct StageTw {
    LaneTw<E> lane;
    SmemTw<E, T, TBMAX> sm;
    // table: global [E/2][T][2]; smem_table: room for E * T entries (unused with lane twiddles)
    __device__ __forceinline__ void init(const cx<float>* table, cx<float>* smem_table, int t, int tid, int nthr)
    {
        if constexpr (LANE) {
            lane.template load<T>(table, t);
        } else {
            for (int i = tid; i < E * T; i += nthr) smem_table[i] = table[i];
            sm.init(smem_table, t);
            __syncthreads();
        }
    }
    template <typename F> __device__ __forceinline__ void apply_all(const cx<float>* v, F out) const
    {
        if constexpr (LANE) {
#pragma unroll
            for (int k1 = 0; k1 < E; ++k1) out(k1, lane.apply(v[gpos<E>(k1)], k1));
        } else {
            sm.apply_all(v, out);
        }
    }
    static constexpr __host__ __device__ int smem_entries() { return kLaneTwiddles ? 0 : E * T; }
};

// named barrier over `nthreads` threads (a multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void named_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- exchange-tile layout for lane groups that are not powers of two ---------------------------------------------
// 8-byte shared-memory accesses are served one half-warp at a time, 16 slots of 8 bytes per wavefront: a half-warp needs
// as many wavefronts as its busiest slot holds DISTINCT addresses.  With T = 10, 12, 20 or 24 lanes per line the lines
// of a warp straddle the half-warps, and the plain layout (tile pitch T + 1, every line's region starting on a 128-byte
// boundary) puts the lanes of two lines into the same slots: ncu shows 30 - 50 % of the shared-memory wavefronts of the
// 300 / 600 / 288 families as bank conflicts.  wfft_layout() picks, at compile time (and again on the host for the
// shared-memory size), the tile pitch and a per-line shift of the tile inside its region with the fewest wavefronts over
// the two access patterns of WarpFFT::run (stage-A store: lanes consecutive; stage-B load: lanes pitch apart).
constexpr __host__ __device__ int wfft_wavefronts(int T, int al, int sh0, int a, int b)
{
    int total = 0;
    for (int half = 0; half < 2; ++half) {
        int addr[16] = {0};
        for (int i = 0; i < 16; ++i) {
            int l = half * 16 + i;
            l = l >= al ? sh0 + (l - al) % (al - sh0) : l;
            addr[i] = (l / T) * b + (l % T) * a;
        }
        int worst = 0;
        for (int slot = 0; slot < 16; ++slot) {
            int distinct = 0;
            for (int i = 0; i < 16; ++i) {
                if ((addr[i] & 15) != slot) continue;
                bool seen = false;
                for (int k = 0; k < i; ++k) seen = seen || addr[k] == addr[i];
                distinct += seen ? 0 : 1;
            }
            worst = distinct > worst ? distinct : worst;
        }
        total += worst;
    }
    return total;
}
struct WfftLayout { int pitch, shift; };
constexpr __host__ __device__ WfftLayout wfft_layout(int E, int T)
{
    if (T >= 32 || (T & (T - 1)) == 0) return WfftLayout{T + 1, 0};        // powers of two: lines and half-warps line up
    const int lpw = 32 / T, al = lpw * T, sh0 = al > 16 ? 16 : 0;
    WfftLayout best{T + 1, 0};
    int cost = 1 << 30;
    for (int p = T; p < T + 8; ++p)
        for (int s = 0; s < 16; ++s) {
            // a line's region starts on a multiple of 16 slots, its tile s * (line in warp) slots further
            const int c = (2 * wfft_wavefronts(T, al, sh0, 1, 16 * 64 + s) + wfft_wavefronts(T, al, sh0, p, 16 * 64 + s)) * 1024 + (p - T) * 16 + s;
            if (c < cost) { cost = c; best = WfftLayout{p, s}; }
        }
    return best;
}
// complex elements a line's region needs: the tile [E][pitch] behind its shift, or the line itself in natural order
constexpr __host__ __device__ int wfft_region(int E, int T)
{
    const WfftLayout l = wfft_layout(E, T);
    const int lpw = T >= 32 ? 1 : 32 / T;
    const int tile = E * l.pitch + l.shift * (lpw - 1);
    return (tile > E * T ? tile : E * T) + 1;
}

template <int E, int T> struct WarpFFT {
    static_assert(T <= 32 || T == 64, "lanes per line");
    static_assert(E % T == 0 && E / T <= 8 && E % 2 == 0, "E/T stage-B transforms per lane");
    static constexpr int N = E * T;
    static constexpr int WPL = T > 32 ? T / 32 : 1;   // warps per line (T = 64: the line's lanes span two adjacent warps)
    static constexpr int LPW = T >= 32 ? 1 : 32 / T;  // lines per warp
    static constexpr int M = E / T;           // stage-B transforms per lane
    // T not a divisor of 32 (24, 12, 10, ...): the lanes >= AL SHADOW the first lanes of the warp -- same line,
    // same t, same loads and (duplicate, identical) stores -- so no code path diverges on them; only
    // one-per-line actions (TMA issue) and sums must skip them
    static constexpr int AL = T >= 32 ? 32 : LPW * T;
    static __device__ __forceinline__ bool is_shadow(int lane) { return lane >= AL; }
    // the lane a shadow lane repeats: one of the SAME half-warp (8-byte shared-memory accesses are served per half-warp:
    // a repeated address inside a half is a broadcast, the same address from the other half costs another wavefront)
    static constexpr int SH0 = AL > 16 ? 16 : 0;
    static __device__ __forceinline__ int real_lane(int lane) { return lane >= AL ? SH0 + (lane - AL) % (AL - SH0) : lane; }
    // lines a CTA of NW warps works on at a time
    static constexpr __host__ __device__ int lines_per_cta(int nw) { return WPL > 1 ? nw / WPL : nw * LPW; }
    // lane within the line / line within the CTA of this thread
    static __device__ __forceinline__ int lane_in_line(int warp, int lane)
    {
        return WPL > 1 ? (warp % WPL) * 32 + lane : real_lane(lane) % T;
    }
    static __device__ __forceinline__ int line_in_warp(int lane) { return WPL > 1 ? 0 : real_lane(lane) / T; }
    static __device__ __forceinline__ int line_in_cta(int warp, int lane) { return WPL > 1 ? warp / WPL : warp * LPW + line_in_warp(lane); }
    // all lanes of a line (bar: a named barrier id private to the line, used only when the line spans warps)
    static __device__ __forceinline__ void line_sync(int bar)
    {
        if constexpr (WPL > 1) named_barrier(bar, 32 * WPL); else __syncwarp();
    }
    static constexpr WfftLayout TL = wfft_layout(E, T);
    static constexpr int TSH = TL.shift;                  // exchange tile (tuned layout): shift per line of the warp, pitch TL.pitch
    static constexpr int RS = wfft_region(E, T);          // minimum slots per line region
    // smallest region stride >= RS with stride % 16 == m: makes accesses by (line-minor, slot-major)
    // thread groups of 16/m lines conflict free
    static constexpr __host__ __device__ int stride_mod16(int m) { return RS + ((m - RS % 16) + 16) % 16; }

    // k (natural index) of register j after run(): see out_index()
    static __device__ __forceinline__ int out_k(int t, int j) { return t + T * (j % M) + E * (j / M); }
    // register that holds output j (j-th in increasing k for this lane)
    static constexpr __host__ __device__ int out_reg(int j) { return (j % M) * T + gpos<T>(j / M); }

    // v: E registers; in: v[n1] = x[n1 T + t]; out: X[out_k(t, j)] = v[out_reg(j)].
    // Forward transform only: inverse transforms are taken as conj(fft(conj(.))) with the
    // conjugations folded into the neighbouring point-wise steps, so that every kernel
    // runs ONE butterfly body (instruction-cache footprint).
    // TUNED: the regions of the lines start on 128-byte boundaries (column kernel: TMA source) and the tile uses the
    // searched pitch and per-line shift (lw: line of the warp this lane works on); otherwise the plain pitch T + 1
    template <bool PRUNED, class TW, bool TUNED = false>
    static __device__ __forceinline__ void run(cx<float>* v, cx<float>* region_base, const TW& tw, int t,
                                               bool pruned_now = true, int bar = 1, int lw = 0)
    {
        constexpr int TP = TUNED ? TL.pitch : T + 1;
        cx<float>* region = region_base + (TUNED ? lw * TSH : 0);
        // first radix-2 level: skipped arithmetic when the upper half of the input is zero; the
        // remaining levels are one shared body
        LaneFFT<E>::template run_first<PRUNED>(v, pruned_now);
        tw.apply_all(v, [&](int k1, cx<float> a) { region[k1 * TP + t] = a; });
        line_sync(bar);
#pragma unroll
        for (int m = 0; m < M; ++m) {
#pragma unroll
            for (int n2 = 0; n2 < T; ++n2) v[m * T + n2] = region[(t + T * m) * TP + n2];
        }
        line_sync(bar);
#pragma unroll
        for (int m = 0; m < M; ++m) LaneFFT<T>::run(v + m * T);
    }
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// one instruction: pull a contiguous, 16-byte aligned chunk (bytes % 16 == 0) into L2 ahead of use
__device__ __forceinline__ void prefetch_l2_bulk(const void* gsrc, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) store of one column: shared memory [ny] complex, natural y order ->
// the tiled G^T layout; ONE instruction per column instead of 32 scattered STG.64 per lane.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n"
                 ::"l"(map), "r"(s), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// all bulk stores of this thread have finished READING shared memory (the buffer may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
// make preceding generic-proxy shared-memory writes visible to the async (TMA) proxy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- mbarrier + 1-D bulk copy (TMA) global -> shared ------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
// bytes % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// threads of a K3 CTA: T * R rounded up to whole warps
constexpr __host__ __device__ int k3_threads(int T, int R) { return (T * R + 31) / 32 * 32; }

struct FastParams {
    XcParams x;
    const cx<float>* twx;   // [EX / 2][TX][2]: w_nx^(k1 t) with rows (k1, k1 + 1) paired per lane
    const cx<float>* twy;   // [EY / 2][TY][2]
    cx<float>* FT0;         // [n][kp][hp0]
    cx<float>* FT1;         // [n][kp][hp1]
    cx<float>* GT;          // [n][ny / rblk][P|Q][kp][rblk]: conjugate of the column-stage output, tiled so
                            // that the rblk rows one K3 CTA owns are ONE contiguous chunk (all kx)
    int hp0, hp1;
    int rblk;               // rows per K3 tile, a power of two <= 16
    int use_tma;            // K2 stores its columns with gt_map (cp.async.bulk.tensor)
    int gt_tiles, gt_pieces; // ny / rblk; TMA stores per column (a box holds at most 256 tiles)
    alignas(64) CUtensorMap gt_map;   // G^T as a 4-D tensor of 8-byte elements: [n * ny / rblk][P|Q][kp][rblk], box = one column
    int flags;              // experiment switches (fb_set_option "fast_flags"): 1/2/4 = no L2 prefetch in
                            // K1/K2/K3, 8 = K2 partners are warps (w, w + NW/2) instead of (2w, 2w+1)
};

// element offset of (row y, column 0, plane P) inside a pair's GT block
__device__ __forceinline__ size_t gt_row_offset(int y, int kp, int rblk)
{
    return (size_t)(y / rblk) * 2 * kp * rblk + (y % rblk);
}

// ---------------------------------------------------------------------------------------------
// K1: forward row transforms, two image rows per complex line, transposed half-spectrum out.
// grid-stride over (pair, image, tile of TR = 2*LPW*NW rows).  NW warps per CTA.
// ---------------------------------------------------------------------------------------------
template <int E, int T, int NW, typename TI, bool PRUNED>
__device__ void kfast_rows_forward(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    constexpr int N = W::N, LPC = W::lines_per_cta(NW), TR = 2 * LPC, NT = 32 * NW;
    constexpr int RS = W::stride_mod16(LPC >= 16 ? 1 : 16 / LPC);
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = W::lane_in_line(warp, lane);             // lane within line
    StageTw<E, T> tw;
    tw.init(fp.twx, regions + LPC * RS, t, tid, NT);
    const int tiles0 = (fp.hp0 + TR - 1) / TR, tiles1 = (fp.hp1 + TR - 1) / TR, tpp = tiles0 + tiles1;   // last tile may be partial
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
        if (tid == 0 && !(fp.flags & 1)) {                 // next tile of this CTA -> L2 (rows are contiguous)
            const int nw_ = work + gridDim.x;
            if (nw_ < p.n * tpp) {
                const int np = nw_ / tpp;
                int nt = nw_ - np * tpp;
                const bool ns = nt >= tiles0;
                if (ns) nt -= tiles0;
                const int nH = ns ? p.h1 : p.h0, nW = ns ? p.w1 : p.w0;
                const TI* nimg = reinterpret_cast<const TI*>(ns ? p.img1 : p.img0) + (size_t)np * nH * nW + (size_t)nt * TR * nW;
                int rows = nH - nt * TR; rows = rows > TR ? TR : rows;
                if (rows > 0) {
                    size_t a = reinterpret_cast<size_t>(nimg);
                    size_t e = (a + (size_t)rows * nW * sizeof(TI)) & ~(size_t)15;
                    a = (a + 15) & ~(size_t)15;
                    if (e > a) prefetch_l2_bulk(reinterpret_cast<const void*>(a), (unsigned)(e - a));
                }
            }
        }
        const int line = W::line_in_cta(warp, lane);       // line within the tile
        const int rA = row0 + 2 * line, rB = rA + 1;
        cx<float>* region = regions + line * RS;
        cx<float> v[E];
        if (rA < H) {                                      // uniform per T-lane group
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) {
                const int xx = n1 * T + t;
                float a = 0.f, b = 0.f;
                if ((!PRUNED || n1 < E / 2) && xx < Wd && !(fp.flags & 512)) {
                    a = (float)__ldg(img + (size_t)rA * Wd + xx);
                    if (rB < H) b = (float)__ldg(img + (size_t)rB * Wd + xx);
                }
                v[n1] = mk<float>(a, b);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) v[n1] = mk<float>(0.f, 0.f);
        }
        W::template run<PRUNED>(v, region, tw, t, true, 1 + line);
#pragma unroll
        for (int j = 0; j < E; ++j) region[W::out_k(t, j)] = v[W::out_reg(j)];
        __syncthreads();
        // separation + transposed store: thread -> (line l minor, k major); both rows of a line are
        // adjacent in FT, so one 16-byte store writes the pair (A[k], B[k])
        {
            constexpr int KSTEP = NT / LPC;                // (threads beyond KSTEP * LPC idle when LPC does not divide NT)
            const int l = tid % LPC;
            const cx<float>* reg = regions + l * RS;
            float4* dst = reinterpret_cast<float4*>(FT + row0 + 2 * l);
            for (int k = tid / LPC; k < ((fp.flags & 1024) || tid >= KSTEP * LPC || row0 + 2 * l >= hp ? 0 : kp); k += KSTEP) {
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
// K2: column stage.  Warp 2 pw transforms the F0 column(s), warp 2 pw + 1 the same F1
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
    constexpr int N = W::N, LPW = W::LPW, WPL = W::WPL, RS = (W::RS + 15) & ~15, NT = 32 * NW;
    constexpr int CPG = W::lines_per_cta(NW) / 2;           // columns per CTA; regions are 128-byte aligned (TMA source)
    constexpr int PT = 64 * WPL;                            // threads of a pair of partner lines
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = W::lane_in_line(warp, lane), lw = W::line_in_warp(lane);
    const bool leader = t == 0 && !W::is_shadow(lane);     // one thread per line
    StageTw<E, T, (FB_LANE_TW_K2 != 0) && E % 8 == 0 && E <= 32> tw;
    tw.init(fp.twy, regions + W::lines_per_cta(NW) * RS, t, tid, NT);
    // partners are ADJACENT warps (2 pw, 2 pw + 1): they sit on different SM sub-partitions, so the
    // four warps an SMSP schedules belong to four different pairs and drift through the
    // FP-only and shared-memory-only phases independently.  (T = 64: a line spans two warps, partners
    // are adjacent warp pairs.)
    const bool split = WPL == 1 && (fp.flags & 8);
    const int wl = warp / WPL;                              // "line warp": index of the warp group that owns a line set
    const bool roleB = split ? warp >= NW / 2 : (wl & 1);
    const int pw = split ? (roleB ? warp - NW / 2 : warp) : wl >> 1;     // pair-of-lines index
    const int lbar = 1 + NW / 2 + wl;                       // named barrier of this line (used when WPL > 1)
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int kp = p.kp, groups = (kp + CPG - 1) / CPG;
    cx<float>* mine = regions + (wl * LPW + lw) * RS;
    cx<float>* other = regions + ((split ? (roleB ? pw : pw + NW / 2) : (wl ^ 1)) * LPW + lw) * RS;
    const int hp = roleB ? fp.hp1 : fp.hp0;
    const bool second_phase = !roleB || mirror;
    const bool tma = fp.use_tma && !(fp.flags & 2048);
    for (int work = blockIdx.x; work < p.n * groups; work += gridDim.x) {
        const int pair = work / groups, grp = work - pair * groups;
        const int col = grp * CPG + pw * LPW + lw;
        const bool live = col < kp;
        if (tma) {                                         // the previous column's TMA store has read `mine`
            if (leader) tma_store_wait_read();
            W::line_sync(lbar);
        }
        if ((WPL > 1 ? t == 0 : lane == 0) && pw == 0 && !(fp.flags & 2)) {     // the CTA's next column group (contiguous in FT) -> L2
            const int nw_ = work + gridDim.x;
            if (nw_ < p.n * groups) {
                const int np = nw_ / groups, ng = nw_ - np * groups;
                int nc = kp - ng * CPG; nc = nc > CPG ? CPG : nc;
                prefetch_l2_bulk((roleB ? fp.FT1 : fp.FT0) + ((size_t)np * kp + (size_t)ng * CPG) * hp, (unsigned)(nc * hp * 8));
            }
        }
        const cx<float>* src = (roleB ? fp.FT1 : fp.FT0) + ((size_t)pair * kp + (live ? col : 0)) * hp;
        cx<float> v[E];
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int y = n1 * T + t;
            cx<float> a = mk<float>(0.f, 0.f);
            if ((!PRUNED0 || n1 < E / 2) && y < hp && !(fp.flags & 128)) a = ldg(src + y);
            v[n1] = a;
        }
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            W::template run<PRUNED0, decltype(tw), true>(v, mine, tw, t, phase == 0, lbar, lw);
            if (phase == 0) {
#pragma unroll
                for (int j = 0; j < E; ++j) mine[W::out_k(t, j)] = v[W::out_reg(j)];
                named_barrier(pw + 1, PT);
                // the output of lane t, k = t + T j, is exactly the input element n1 = j of the
                // next transform: a register permutation, no exchange needed
                // unscaled: 1 / (ny nx) is applied once per pair by the finalize kernel (XcParams::out_scale);
                // confidence and sub-pixel offsets are ratios
                if constexpr (E <= 32) {
                    cx<float> u[E];
                    if (!roleB) {
#pragma unroll
                        for (int j = 0; j < E; ++j) u[j] = cmulc(v[W::out_reg(j)], other[W::out_k(t, j)]);      // conj(P) = F0 conj(F1)
                    } else {
#pragma unroll
                        for (int j = 0; j < E; ++j) {                                                         // conj(Q) = conj(F1 F0)
                            const cx<float> o = other[W::out_k(t, j)], m = v[W::out_reg(j)];
                            u[j] = mk<float>(m.x * o.x - m.y * o.y, -(m.x * o.y) - m.y * o.x);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < E; ++j) v[j] = u[j];
                } else {
                    // E = 64: no room for a second register file image -- the own spectrum is read back from
                    // the region it was just written to (natural order) instead of being permuted in registers
                    if (!roleB) {
#pragma unroll
                        for (int j = 0; j < E; ++j) v[j] = cmulc(mine[W::out_k(t, j)], other[W::out_k(t, j)]);
                    } else {
#pragma unroll
                        for (int j = 0; j < E; ++j) {
                            const cx<float> o = other[W::out_k(t, j)], m = mine[W::out_k(t, j)];
                            v[j] = mk<float>(m.x * o.x - m.y * o.y, -(m.x * o.y) - m.y * o.x);
                        }
                    }
                }
                named_barrier(pw + 1, PT);
                if (!second_phase) break;
            } else if (!(fp.flags & 64)) {
                if (tma) {
                    // natural y order into the (free) transpose region, then one bulk tensor store scatters
                    // the column into its ny / rblk tiles.  Every lane takes this path (columns past kp write
                    // their region too, nothing is stored for them): no divergence around the line barrier.
#pragma unroll
                    for (int j = 0; j < E; ++j) mine[W::out_k(t, j)] = v[W::out_reg(j)];
                    fence_proxy_async_smem();
                    W::line_sync(lbar);
                    if (leader && live) {
                        // the box of the tensor map holds at most 256 tiles: long columns leave in several pieces
                        tma_store_4d(&fp.gt_map, mine, 0, col, roleB ? 1 : 0, pair * fp.gt_tiles);
                        for (int i = 1; i < fp.gt_pieces; ++i)
                            tma_store_4d(&fp.gt_map, mine + i * 256 * fp.rblk, 0, col, roleB ? 1 : 0, pair * fp.gt_tiles + i * 256);
                    }
                } else if (live) {
                    const int R = fp.rblk;
                    cx<float>* dst = fp.GT + (size_t)pair * 2 * kp * N + ((size_t)(roleB ? kp : 0) + col) * R;
#pragma unroll
                    for (int j = 0; j < E; ++j) dst[gt_row_offset(W::out_k(t, j), kp, R)] = v[W::out_reg(j)];
                }
            }
        }
    }
    if (tma) {                                             // shared memory must outlive the last store's read
        if (leader) tma_store_wait_read();
        W::line_sync(lbar);
    }
}

// all but the `pending` most recent bulk stores of this thread have finished READING shared memory
template <int PENDING> __device__ __forceinline__ void tma_store_wait_read_but()
{
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(PENDING) : "memory");
}

// ---------------------------------------------------------------------------------------------
// K2, one line per column PAIR ("solo"): the lanes of a line transform the F0 column, keep its spectrum in
// registers, transform the F1 column, form conj(P) and conj(Q) in registers and transform both -- no spectrum
// exchange through shared memory (a fifth of the column kernel's shared-memory wavefronts) and no barrier
// between warps at all: every warp drifts through its load / butterfly / transpose phases on its own.  Costs
// registers (two 2E-register images per lane: one CTA of NW warps per SM) and two transpose regions per line (the
// TMA store of P drains from one while Q is transformed in the other).  The four transforms of a work item run
// through ONE copy of the butterfly code (rolled loop); global loads are issued a whole transform ahead of use
// into the register image that is idle at that point.  E <= 32, lines inside one warp.
// ---------------------------------------------------------------------------------------------
// the permutation j -> out_reg(j) as a list of cycles: order[i] = i-th element visited, start / last = first / last of its cycle
template <int E, int T> struct PermCycles {
    int order[E];
    bool start[E], last[E];
    constexpr __host__ __device__ PermCycles() : order{}, start{}, last{}
    {
        bool seen[E] = {};
        int n = 0;
        for (int j0 = 0; j0 < E; ++j0) {
            if (seen[j0]) continue;
            int j = j0;
            bool first = true;
            while (!seen[j]) {
                seen[j] = true; order[n] = j; start[n] = first; last[n] = false; first = false; ++n;
                j = WarpFFT<E, T>::out_reg(j);
            }
            last[n - 1] = true;
        }
    }
};
// v[j] <- conj(P)_j = F0 conj(F1), park[j] <- conj(Q)_j = conj(F0 F1) with F0 = park[out_reg(j)], F1 = v[out_reg(j)] (old values)
template <int E, int T, int I = 0> struct SoloProducts {
    static __device__ __forceinline__ void run(cx<float>* v, cx<float>* park, cx<float>& a0, cx<float>& b0)
    {
        constexpr PermCycles<E, T> pc{};
        constexpr int j = pc.order[I], src = WarpFFT<E, T>::out_reg(j);
        constexpr bool st = pc.start[I], la = pc.last[I];
        if constexpr (st) { a0 = park[j]; b0 = v[j]; }                 // the cycle's first element is overwritten first: keep it
        const cx<float> a = la ? a0 : park[src], b = la ? b0 : v[src];
        v[j] = cmulc(a, b);
        park[j] = mk<float>(a.x * b.x - a.y * b.y, -(a.x * b.y) - a.y * b.x);
        if constexpr (I + 1 < E) SoloProducts<E, T, I + 1>::run(v, park, a0, b0);
    }
};

template <int E, int T, int NW, bool PRUNED0>
__device__ void kfast_columns_solo(const FastParams& fp, unsigned char* smem)
{
    using W = WarpFFT<E, T>;
    static_assert(W::WPL == 1 && E <= 32, "solo column kernel: a line inside one warp, 2 x 2E registers per lane");
    constexpr int N = W::N, LPW = W::LPW, RS = (W::RS + 15) & ~15, NT = 32 * NW;
    constexpr int CPG = NW * LPW;                           // columns per CTA and work item
    const XcParams& p = fp.x;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = W::lane_in_line(warp, lane), lw = W::line_in_warp(lane);
    const bool leader = t == 0 && !W::is_shadow(lane);     // one thread per line
    StageTw<E, T, false, 2> tw;                             // (short twiddle batches: registers are what this kernel is short of)
    tw.init(fp.twy, regions + NW * LPW * 2 * RS, t, tid, NT);
    const bool mirror = p.conf_mode == CONF_MIRROR;
    const int kp = p.kp, groups = (kp + CPG - 1) / CPG, total = p.n * groups;
    cx<float>* regA = regions + ((warp * LPW + lw) * 2) * RS;
    cx<float>* regB = regA + RS;
    const int hp0 = fp.hp0, hp1 = fp.hp1;
    cx<float> v[E], park[E];
    // column `col` of pair `pair` of F0 / F1 -> dst (zero rows pruned)
    auto load_col = [&](cx<float>* dst, const cx<float>* FT, int hp, int pair, int col) {
        const cx<float>* src = FT + ((size_t)pair * kp + (col < kp ? col : 0)) * hp;
#pragma unroll
        for (int n1 = 0; n1 < E; ++n1) {
            const int y = n1 * T + t;
            cx<float> a = mk<float>(0.f, 0.f);
            if ((!PRUNED0 || n1 < E / 2) && y < hp) a = ldg(src + y);
            dst[n1] = a;
        }
    };
    if ((int)blockIdx.x < total) {
        const int pair = blockIdx.x / groups, grp = blockIdx.x - pair * groups;
        load_col(park, fp.FT0, hp0, pair, grp * CPG + warp * LPW + lw);
    }
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
        const int pair = work / groups, grp = work - pair * groups;
        const int col = grp * CPG + warp * LPW + lw;
        const bool live = col < kp;
        const int nw_ = work + gridDim.x;
        if (lane == 0 && warp == 0 && !(fp.flags & 2) && nw_ + (int)gridDim.x < total) {      // the work item after the next -> L2
            const int w2 = nw_ + gridDim.x;
            const int np = w2 / groups, ng = w2 - np * groups;
            int nc = kp - ng * CPG; nc = nc > CPG ? CPG : nc;
            prefetch_l2_bulk(fp.FT0 + ((size_t)np * kp + (size_t)ng * CPG) * hp0, (unsigned)(nc * hp0 * 8));
            prefetch_l2_bulk(fp.FT1 + ((size_t)np * kp + (size_t)ng * CPG) * hp1, (unsigned)(nc * hp1 * 8));
        }
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = park[j];        // the F0 column (loaded during the previous item's last transform)
        load_col(park, fp.FT1, hp1, pair, col);             // the F1 column arrives under the first transform
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            cx<float>* reg = pass == 3 ? regB : regA;
            if (pass == 0 || pass == 3) {                   // the store that last drained from this region has read it
                if (leader) tma_store_wait_read_but<1>();
                __syncwarp();
            }
            W::template run<PRUNED0, decltype(tw), true>(v, reg, tw, t, pass < 2, 1, lw);
            if (pass == 0) {
#pragma unroll
                for (int j = 0; j < E; ++j) { const cx<float> x = v[j]; v[j] = park[j]; park[j] = x; }
            } else if (pass == 1) {
                // output k = t + T j of lane t is input element n1 = j of the next transform: a register permutation.
                // unscaled: 1 / (ny nx) is applied once per pair by the finalize kernel (XcParams::out_scale)
                // (in place, cycle by cycle of the permutation: the register images sit in fixed registers across the rolled loop)
                cx<float> a0, b0;
                SoloProducts<E, T>::run(v, park, a0, b0);
            } else {
                // natural y order into the region, one bulk tensor store scatters the column into its ny / rblk tiles
#pragma unroll
                for (int j = 0; j < E; ++j) reg[W::out_k(t, j)] = v[W::out_reg(j)];
                fence_proxy_async_smem();
                __syncwarp();
                if (leader && live) tma_store_4d(&fp.gt_map, reg, 0, col, pass - 2, pair * fp.gt_tiles);
                if (pass == 2) {
                    if (!mirror) { pass = 3; }              // one surface only: skip the Q transform (park is dead)
                    else {
#pragma unroll
                        for (int j = 0; j < E; ++j) v[j] = park[j];
                    }
                    if (nw_ < total) {                      // next item's F0 column arrives under the last transform
                        const int np = nw_ / groups, ng = nw_ - np * groups;
                        load_col(park, fp.FT0, hp0, np, ng * CPG + warp * LPW + lw);
                    }
                }
            }
        }
    }
    if (leader) tma_store_wait_read();                      // shared memory must outlive the last store's read
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// K3: inverse row transforms + per-line maxima, row-batched.  A line is (P row y, Q row y) with
// the mirror term, else (P row y, P row y+1).  The CTA (T * R threads) owns R consecutive
// lines.  Stage A: thread (t, r) -- r minor, so a lane group reads R consecutive y of one GT
// column, contiguous -- assembles conj(Z)[n1 T + t] of line r from the (conjugated) P / Q
// columns by Hermitian extension, radix-E in registers, twiddles, writes X[k1][t][r].
// Stage B: thread (k1, r) reads X[k1][.][r], radix-T, reduces maxima per line.
// The forward transform of conj(Z) is conj(surface): C = Re, mirror = -Im (|.| only), second
// row = -Im.  Only per-line maxima are produced; the finalize kernel locates x inside the
// winning row (np.argmax order: lowest row, then lowest x).
// ---------------------------------------------------------------------------------------------
template <int E, int T, int R, int RB = R>
__device__ void kfast_rows_inverse(const FastParams& fp, unsigned char* smem)
{
    static_assert(R == 4 || R == 8 || R == 16, "lines per CTA");
    static_assert(RB % R == 0, "a CTA owns a whole GT tile (RB rows) or an aligned part of one");
    constexpr int N = E * T, M = E / T, NT = k3_threads(T, R), NWARP = NT / 32;   // RB == fp.rblk
    constexpr int XS = T * R + (R < 16 ? R : 0);              // k1 stride of the exchange tile (conflict free)
    const XcParams& p = fp.x;
    cx<float>* X = reinterpret_cast<cx<float>*>(smem);
    cx<float>* twsm = X + E * XS;
    float* red = reinterpret_cast<float*>(twsm + StageTw<E, T>::smem_entries());   // [NWARP][R][2] floats + doubles after
    double* redd = reinterpret_cast<double*>(red + NWARP * R * 2);
    const int warp = threadIdx.x >> 5;
    const bool shadow = (int)threadIdx.x >= T * R;            // T * R not a multiple of 32: the surplus threads repeat real
    constexpr int LASTW = (T * R) / 32 * 32, LASTN = (T * R) % 32 ? (T * R) % 32 : 32;   // threads of their own warp (identical
    const int tid = shadow ? LASTW + ((int)threadIdx.x - T * R) % LASTN : (int)threadIdx.x;   // loads / stores; sums skip them)
    const int r = tid % R, tq = tid / R;                      // stage A: t = tq; stage B: k1 = tq + T m
    StageTw<E, T> tw;
    tw.init(fp.twx, twsm, tq, threadIdx.x, NT);            // (the table copy strides over ALL threads, shadows included)
    const bool mirror = p.conf_mode == CONF_MIRROR, want_std = p.conf_mode == CONF_STD;
    const int ny = p.ny, kp = p.kp;
    const int lines_pp = mirror ? ny : (ny + 1) / 2;          // lines per pair
    const int tiles = (lines_pp + R - 1) / R;
    const size_t plane = (size_t)kp * ny;
    for (int work = blockIdx.x; work < p.n * tiles; work += gridDim.x) {
        const int pair = work / tiles, tile = work - pair * tiles;
        const int gl = tile * R + r;                           // line index inside the pair (always live:
        const int y0 = mirror ? gl : 2 * gl;                   //  R divides the power-of-two line count)
        const cx<float>* A = fp.GT + (size_t)pair * 2 * plane + gt_row_offset(y0, kp, RB);
        const cx<float>* B = mirror ? A + (size_t)kp * RB : A + 1;
        if (threadIdx.x == 0 && !(fp.flags & 4)) {             // the GT tile of the CTA's next work item is one
            const int nw_ = work + gridDim.x;                  // contiguous chunk -> L2 (once per GT tile)
            const int rows = mirror ? R : 2 * R;               // surface rows per work item
            if (nw_ < p.n * tiles && (nw_ * rows) % RB == 0) {
                const int np = nw_ / tiles, nt = nw_ - np * tiles;
                const cx<float>* nb = fp.GT + (size_t)np * 2 * plane;
                const int yb = nt * rows / RB;                 // first GT tile of that work item
                if (mirror) {
                    prefetch_l2_bulk(nb + (size_t)yb * 2 * kp * RB, (unsigned)(2 * kp * RB * 8));
                } else {
                    for (int i = 0; i < (rows + RB - 1) / RB; ++i)
                        prefetch_l2_bulk(nb + (size_t)(yb + i) * 2 * kp * RB, (unsigned)(kp * RB * 8));
                }
            }
        }
        __syncthreads();                                       // X free
        {
            const int t = tq;
            cx<float> v[E];
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) {
                const int k = n1 * T + t;
                cx<float> z;
                if (fp.flags & 256) {                          // diagnostic: no global loads
                    z = mk<float>((float)k, (float)r);
                } else if (n1 < E / 2 || (n1 == E / 2 && t == 0)) {
                    const cx<float> a = ldg(A + (size_t)k * RB), b = ldg(B + (size_t)k * RB);
                    // conj(P + iQ) with stored a = conj(P), b = conj(Q):  a - i b
                    z = (k == 0 || 2 * k == N) ? mk<float>(a.x, -b.x) : mk<float>(a.x + b.y, a.y - b.x);
                } else {
                    const int m = N - k;
                    const cx<float> a = ldg(A + (size_t)m * RB), b = ldg(B + (size_t)m * RB);
                    // conj(conj(P) + i conj(Q)) = P - i Q = conj(a) - i conj(b)
                    z = mk<float>(a.x - b.y, -a.y - b.x);
                }
                v[n1] = z;
            }
            LaneFFT<E>::run(v);
            tw.apply_all(v, [&](int k1, cx<float> a) { X[k1 * XS + t * R + r] = a; });
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
            for (int m = 0; m < M; ++m) LaneFFT<T>::run(u + m * T);
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
                    if (!mirror) { sum -= (double)u[j].y; sumsq += (double)u[j].y * (double)u[j].y; }
                }
            }
        }
        if (shadow) { sum = 0.0; sumsq = 0.0; }
        // lanes with equal r inside the warp, then the warps through shared memory
#pragma unroll
        for (int off = R; off < 32; off <<= 1) {
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
            second = fmaxf(second, __shfl_xor_sync(0xffffffffu, second, off));
            if (want_std) {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
                sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
            }
        }
        if ((threadIdx.x & 31) < R) {
            red[(warp * R + r) * 2] = best; red[(warp * R + r) * 2 + 1] = second;
            if (want_std) { redd[(warp * R + r) * 2] = sum; redd[(warp * R + r) * 2 + 1] = sumsq; }
        }
        __syncthreads();
        if (threadIdx.x < R) {
            for (int w = 1; w < NWARP; ++w) {
                best = fmaxf(best, red[(w * R + r) * 2]); second = fmaxf(second, red[(w * R + r) * 2 + 1]);
                if (want_std) { sum += redd[(w * R + r) * 2]; sumsq += redd[(w * R + r) * 2 + 1]; }
            }
            // idx = first flat index of the row that holds the maximum; x is resolved by the finalize kernel
            int row = y0;
            float mir = 0.f;
            if (mirror) mir = second;
            else if (second > best) { best = second; row += 1; }
            Partial& o = p.part[(size_t)pair * p.nrt + gl];
            o.val = (double)best; o.mir = (double)mir; o.sum = sum; o.sumsq = sumsq; o.idx = row * N; o.pad = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3, TMA-fed: same arithmetic as kfast_rows_inverse, but the GT tile of the CTA (one contiguous chunk)
// is brought into shared memory by ONE bulk copy (cp.async.bulk + mbarrier) that is issued as soon as
// the exchange buffer of the previous tile has been read, i.e. it runs under the second butterfly
// stage and the reductions of the previous tile.  The exchange tile X reuses the landing buffer in
// place: every thread pulls its inputs into registers, a barrier, then X overwrites the tile.
// MIRROR: line = (P row y, Q row y); otherwise line = (P row y, P row y + 1).
// ---------------------------------------------------------------------------------------------
template <int E, int T, int R, bool MIRROR>
__device__ void kfast_rows_inverse_tma(const FastParams& fp, unsigned char* smem)
{
    static_assert(R == 4 || R == 8 || R == 16, "lines per CTA = rows per GT tile");
    constexpr int N = E * T, M = E / T, NT = k3_threads(T, R), NWARP = NT / 32, KP = N / 2 + 1;
    constexpr int XS = T * R + (R < 16 ? R : 0);              // k1 stride of the exchange tile (conflict free)
    static_assert(E * XS >= 2 * KP * R, "the exchange tile covers the landing buffer");
    constexpr unsigned TILE_BYTES = 2u * KP * R * 8u;
    const XcParams& p = fp.x;
    cx<float>* X = reinterpret_cast<cx<float>*>(smem);
    cx<float>* twsm = X + E * XS;
    float* red = reinterpret_cast<float*>(twsm + StageTw<E, T>::smem_entries());   // [NWARP][R][2] floats, doubles, mbarrier
    double* redd = reinterpret_cast<double*>(red + NWARP * R * 2);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(redd + NWARP * R * 2);
    const int warp = threadIdx.x >> 5;
    const bool shadow = (int)threadIdx.x >= T * R;            // T * R not a multiple of 32: the surplus threads repeat real
    constexpr int LASTW = (T * R) / 32 * 32, LASTN = (T * R) % 32 ? (T * R) % 32 : 32;   // threads of their own warp (identical
    const int tid = shadow ? LASTW + ((int)threadIdx.x - T * R) % LASTN : (int)threadIdx.x;   // loads / stores; sums skip them)
    const int r = tid % R, tq = tid / R;                      // stage A: t = tq; stage B: k1 = tq + T m
    if (threadIdx.x == 0) mbar_init(bar, 1);
    StageTw<E, T> tw;
    tw.init(fp.twx, twsm, tq, threadIdx.x, NT);            // (the table copy strides over ALL threads, shadows included)
    fence_proxy_async_smem();
    __syncthreads();
    const bool want_std = !MIRROR && p.conf_mode == CONF_STD;
    const int ny = p.ny;
    const int lines_pp = MIRROR ? ny : ny / 2;                // lines per pair (R divides it: power-of-two grids)
    const int tiles = lines_pp / R;
    const int total = p.n * tiles;
    const size_t plane = (size_t)KP * ny;
    // issue the bulk copy of work item w (thread 0): MIRROR: GT tile w (P | Q planes, contiguous);
    // otherwise the P planes of GT tiles 2 t and 2 t + 1
    auto issue = [&](int w) {
        const int np = w / tiles, nt = w - np * tiles;
        const cx<float>* base = fp.GT + (size_t)np * 2 * plane;
        mbar_expect_tx(bar, TILE_BYTES);
        if (MIRROR) {
            tma_load_1d(X, base + (size_t)nt * 2 * KP * R, TILE_BYTES, bar);
        } else {
            tma_load_1d(X, base + (size_t)(2 * nt) * 2 * KP * R, TILE_BYTES / 2, bar);
            tma_load_1d(X + KP * R, base + (size_t)(2 * nt + 1) * 2 * KP * R, TILE_BYTES / 2, bar);
        }
    };
    auto prefetch = [&](int w) {
        if (w >= total || (fp.flags & 4)) return;
        const int np = w / tiles, nt = w - np * tiles;
        const cx<float>* base = fp.GT + (size_t)np * 2 * plane;
        if (MIRROR) {
            prefetch_l2_bulk(base + (size_t)nt * 2 * KP * R, TILE_BYTES);
        } else {
            prefetch_l2_bulk(base + (size_t)(2 * nt) * 2 * KP * R, TILE_BYTES / 2);
            prefetch_l2_bulk(base + (size_t)(2 * nt + 1) * 2 * KP * R, TILE_BYTES / 2);
        }
    };
    if (threadIdx.x == 0 && (int)blockIdx.x < total) { issue(blockIdx.x); prefetch(blockIdx.x + gridDim.x); }
    unsigned parity = 0;
    // this thread's inputs inside the landing buffer
    const int offA = MIRROR ? r : ((2 * r) / R) * KP * R + (2 * r) % R;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
        const int pair = work / tiles, tile = work - pair * tiles;
        const int gl = tile * R + r;                           // line index inside the pair
        const int y0 = MIRROR ? gl : 2 * gl;
        mbar_wait(bar, parity);
        parity ^= 1;
        cx<float> v[E];
        {
            const int t = tq;
            const cx<float>* A = X + offA;
#pragma unroll
            for (int n1 = 0; n1 < E; ++n1) {
                const int k = n1 * T + t;
                const bool direct = n1 < E / 2 || (n1 == E / 2 && t == 0);
                const int kk = direct ? k : N - k;
                cx<float> a, b;
                if (MIRROR) {
                    a = A[kk * R]; b = A[KP * R + kk * R];
                } else {
                    const float4 ab = *reinterpret_cast<const float4*>(A + kk * R);     // rows y0, y0 + 1 are adjacent
                    a = mk<float>(ab.x, ab.y); b = mk<float>(ab.z, ab.w);
                }
                // stored values are conj(P), conj(Q).  direct: conj(P + iQ) = a - i b (k = 0, N/2: real parts
                // only);  mirrored index: conj(conj(P) + i conj(Q)) = conj(a) - i conj(b)
                v[n1] = direct ? ((k == 0 || 2 * k == N) ? mk<float>(a.x, -b.x) : mk<float>(a.x + b.y, a.y - b.x))
                               : mk<float>(a.x - b.y, -a.y - b.x);
            }
        }
        __syncthreads();                                       // every input is in registers: X may be overwritten
        LaneFFT<E>::run(v);
        tw.apply_all(v, [&](int k1, cx<float> a) { X[k1 * XS + tq * R + r] = a; });
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
            __syncthreads();                                   // X has been read: the next tile may land
            if (threadIdx.x == 0) {
                const int nw_ = work + gridDim.x;
                if (nw_ < total) { issue(nw_); prefetch(nw_ + gridDim.x); }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) LaneFFT<T>::run(u + m * T);
            best = u[0].x; second = MIRROR ? fabsf(u[0].y) : -u[0].y;
#pragma unroll
            for (int j = 1; j < E; ++j) {
                best = fmaxf(best, u[j].x);
                second = fmaxf(second, MIRROR ? fabsf(u[j].y) : -u[j].y);
            }
            if (want_std) {
#pragma unroll
                for (int j = 0; j < E; ++j) {
                    sum += (double)u[j].x; sumsq += (double)u[j].x * (double)u[j].x;
                    sum -= (double)u[j].y; sumsq += (double)u[j].y * (double)u[j].y;
                }
            }
        }
        if (shadow) { sum = 0.0; sumsq = 0.0; }
        // lanes with equal r inside the warp, then the warps through shared memory
#pragma unroll
        for (int off = R; off < 32; off <<= 1) {
            best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, off));
            second = fmaxf(second, __shfl_xor_sync(0xffffffffu, second, off));
            if (want_std) {
                sum += __shfl_xor_sync(0xffffffffu, sum, off);
                sumsq += __shfl_xor_sync(0xffffffffu, sumsq, off);
            }
        }
        if ((threadIdx.x & 31) < R) {
            red[(warp * R + r) * 2] = best; red[(warp * R + r) * 2 + 1] = second;
            if (want_std) { redd[(warp * R + r) * 2] = sum; redd[(warp * R + r) * 2 + 1] = sumsq; }
        }
        __syncthreads();
        if (threadIdx.x < R) {
            for (int w = 1; w < NWARP; ++w) {
                best = fmaxf(best, red[(w * R + r) * 2]); second = fmaxf(second, red[(w * R + r) * 2 + 1]);
                if (want_std) { sum += redd[(w * R + r) * 2]; sumsq += redd[(w * R + r) * 2 + 1]; }
            }
            int row = y0;
            float mir = 0.f;
            if (MIRROR) mir = second;
            else if (second > best) { best = second; row += 1; }
            Partial& o = p.part[(size_t)pair * p.nrt + gl];
            o.val = (double)best; o.mir = (double)mir; o.sum = sum; o.sumsq = sumsq; o.idx = row * N; o.pad = 0;
        }
    }
}

}  // namespace fb
