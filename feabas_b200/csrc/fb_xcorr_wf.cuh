// fb_xcorr_wf.cuh -- "warp-fused" kernel: the whole xcorr of one block pair on one SM, every 1-D transform held in
// the registers of a few lanes.
//
// Replaces the arithmetic of feabas/matcher.py:22-135 for the small FFT grids of the finest pyramid levels
// (feabas/matcher.py:834-846 at the stitching / thumbnail block sizes: 74 x 67 -> FFT 150 x 135, 60 x 75 ->
// 120 x 150, 50 x 50 -> 100 x 100, ...), where both half spectra of a pair fit in one SM's shared memory.
//
// Differences to the first fused kernel (fb_xcorr.cuh kf_fused: shared-memory radix passes, a block-wide barrier
// after every pass, ~25 barriers per pair): here a line of N = E * T points is transformed by T lanes of a warp that
// hold E points each (stage A: radix-E in registers, twiddle, ONE exchange through a region private to the line,
// stage B: radix-T in registers; E need not be a multiple of T -- the E stage-B transforms are dealt to the T lanes
// round robin and the last round is partly idle).  Lines only synchronise inside their warp; the CTA meets at a handful
// of barriers per pair (images staged, rows done, columns done, maxima reduced, peak rows done).
//
//   stage      lines                         source -> destination
//   A  rows    (H0 + 1) / 2 + (H1 + 1) / 2   two image rows per complex line -> S[y][k] (F0) and S[y][kp + k] (F1)
//   B  columns kp x (2 + 1 or 2)             a line slot owns column k of both spectra: forward F0[:, k] and F1[:, k]
//                                            (rows >= H are zero, not read), conj(P) = F0 conj(F1) and conj(Q) =
//                                            conj(F0 F1) in registers (matcher.py:65,114), forward transform of the
//                                            conjugates = conjugate of the inverse, back in place
//   C  rows    ny (MIRROR) or (ny + 1) / 2   Hermitian extension of conj(P) - i conj(Q) -> surface row(s); only
//                                            maxima (value, flat index, |mirror|, sums) survive
//   D          3 lines                       rows py - 1, py, py + 1 again for the 3 x 3 sub-pixel fit (:84-106)
//
// Forward transforms only (one butterfly body per line length), conjugations folded into the point-wise steps; the
// spectra are unscaled, 1 / (ny nx) is applied to the reported maxima.  CUDA only.
#pragma once
#include "fb_regfft.cuh"
#include "fb_xcorr.cuh"

namespace fb {

// Shared memory holds real and imaginary parts in SEPARATE planes, accessed with 4-byte loads / stores: the lines of
// this kernel have 9, 10, 5 ... lanes, and 8-byte accesses (served per half-warp, 16 slots) cannot be laid out
// conflict free for such lane groups (3 - 4 wavefronts instead of 2, measured), while a 4-byte access of a whole warp
// is ONE wavefront as soon as its 32 lanes fall into 32 different banks.
// Bank model: lane l of the warp belongs to line j = l / T with lane-in-line t = l % T (lanes >= lines * T shadow lanes
// 0, 1, ... and repeat their addresses) and touches word j * b + t * a; the cost is the largest number of DISTINCT words
// in one bank.  Paddings and pitches are chosen at compile time to minimise it.
constexpr int wf_cost(int T, int a, int b)
{
    const int lpw = 32 / T, al = lpw * T;
    int addr[32] = {0};
    for (int l = 0; l < al; ++l) addr[l] = (l / T) * b + (l % T) * a;
    int worst = 0;
    for (int bank = 0; bank < 32; ++bank) {
        int distinct = 0;
        for (int i = 0; i < al; ++i) {
            if ((addr[i] & 31) != bank) continue;
            bool seen = false;
            for (int k = 0; k < i; ++k) seen = seen || addr[k] == addr[i];
            distinct += seen ? 0 : 1;
        }
        worst = distinct > worst ? distinct : worst;
    }
    return worst;
}

struct WfPitch { int p, pe, rs; };
// exchange-tile pitches (forward: rows of T entries, pitch p; reverse: rows of E entries, pitch pe) and the region
// stride rs of a line: the combination with the fewest wavefronts over the access patterns of LineFFT
constexpr WfPitch wf_pitches(int E, int T, bool reverse, int xp)
{
    WfPitch best{T + 1, E + 1, 0};
    int cost = 1 << 30;
    for (int p = T; p <= T + xp; ++p) {
        const int rs0 = E * p > E * T ? E * p : E * T;
        for (int pad = 0; pad < 32; ++pad) {
            const int rs = rs0 + pad;
            // stage-A store and natural-order store (lanes consecutive), stage-B load (lane stride p)
            const int c = ((reverse ? 3 : 2) * wf_cost(T, 1, rs) + wf_cost(T, p, rs)) * 4096 + rs;     // ties: the smaller region
            if (c < cost) { cost = c; best = WfPitch{p, E + 1, rs}; }
        }
    }
    if (reverse) {                                  // reverse exchange: store by k1 (lanes consecutive), load with lane stride pe
        int cost_pe = 1 << 30;
        for (int pe = E; pe <= E + xp; ++pe) {
            const int c = wf_cost(T, pe, best.rs) * 64 + (pe - E);
            if (c < cost_pe) { cost_pe = c; best.pe = pe; }
        }
        while (T * best.pe > best.rs) best.rs += 32;   // room for the [T][pe] tile, same stride modulo 32
    }
    return best;
}

// XP: how much wider than its payload a tile row may get in the search for conflict-free pitches (shared memory is the
// budget: a pair's spectra leave 40 - 60 KB for all exchange regions at the largest grids)
template <int E, int T, bool REV = false, int XP = 1> struct LineFFT {
    static_assert(T >= 2 && T <= 16, "lanes per line");
    static constexpr int N = E * T;
    static constexpr int M = (E + T - 1) / T;          // stage-B rounds per lane
    static constexpr int LPW = 32 / T;                  // lines per warp
    static constexpr int AL = LPW * T;                  // lanes >= AL shadow the first lanes (same line, same work)
    static constexpr int U = M * T > E ? M * T : E;     // registers of the stage-B image
    static constexpr WfPitch PP = wf_pitches(E, T, REV, XP);
    static constexpr int P = PP.p;                      // pitch of the forward exchange tile [E][T]
    static constexpr int PE = PP.pe;                    // pitch of the reverse exchange tile [T][E]
    static constexpr int RS = PP.rs;                    // exchange region of one line: RS real parts, then RS imaginary parts

    static __device__ __forceinline__ int lane_id(int lane) { return lane >= AL ? lane - AL : lane; }
    // natural index of stage-B register (m, k2): k = k1 + E k2 with k1 = t + T m (valid while k1 < E)
    static __device__ __forceinline__ bool valid(int t, int m) { return t + T * m < E; }

    static __device__ __forceinline__ void put(float* region, int i, cx<float> z) { region[i] = z.x; region[RS + i] = z.y; }
    static __device__ __forceinline__ cx<float> get(const float* region, int i) { return mk<float>(region[i], region[RS + i]); }

    // v[n1] = x[n1 T + t]  ->  u[m T + gpos<T>(k2)] = X[(t + T m) + E k2]
    static __device__ __forceinline__ void run(cx<float>* v, cx<float>* u, float* region, const cx<float>* tw, int t)
    {
        LaneFFT<E>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) {
            cx<float> a = v[gpos<E>(k1)];
            if (k1) a = cmul(a, tw[k1 * T + t]);
            put(region, k1 * P + t, a);
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; ++m) {
            int k1 = t + T * m;
            k1 = k1 < E ? k1 : E - 1;                   // idle slots of the last round recompute a valid line
#pragma unroll
            for (int n2 = 0; n2 < T; ++n2) u[m * T + n2] = get(region, k1 * P + n2);
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < M; ++m) LaneFFT<T>::run(u + m * T);
    }
    // The same transform for an input that is ALREADY in the output layout of run():  u[m T + gpos<T>(k2)] =
    // x[(t + T m) + E k2]  ->  v[gpos<E>(q1)] = Y[t + T q1], the input layout of run() and of the column stores.
    // With n = k1 + E k2 and q = q2 + T q1:  w_N^(n q) = w_T^(k2 q2) w_N^(k1 q2) w_E^(k1 q1): radix-T over the lane's
    // own registers, twiddle, ONE exchange, radix-E.  Two transforms back to back (the column stage) need no
    // natural-order round trip in between.
    static __device__ __forceinline__ void run_rev(cx<float>* u, cx<float>* v, float* region, const cx<float>* tw, int t)
    {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            cx<float> w[T];
#pragma unroll
            for (int k2 = 0; k2 < T; ++k2) w[k2] = u[m * T + gpos<T>(k2)];
            LaneFFT<T>::run(w);
            int k1 = t + T * m;
            const bool ok = k1 < E;
            k1 = ok ? k1 : E - 1;
#pragma unroll
            for (int q2 = 0; q2 < T; ++q2) {
                cx<float> a = w[gpos<T>(q2)];
                if (q2) a = cmul(a, tw[k1 * T + q2]);
                if (ok) put(region, q2 * PE + k1, a);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k1 = 0; k1 < E; ++k1) v[k1] = get(region, t * PE + k1);
        __syncwarp();
        LaneFFT<E>::run(v);
    }
    // f(k, value) for every output this lane owns
    template <typename F> static __device__ __forceinline__ void for_each(const cx<float>* u, int t, F f)
    {
#pragma unroll
        for (int m = 0; m < M; ++m) {
            if (valid(t, m)) {
#pragma unroll
                for (int k2 = 0; k2 < T; ++k2) f(t + T * m + E * k2, u[m * T + gpos<T>(k2)]);
            }
        }
    }
};

struct WfBest {
    float val; int idx; float mir; double sum, sumsq; int any;
};

__device__ __forceinline__ void wf_take(WfBest& a, float v, int idx)
{
    if (!a.any || v > a.val || (v == a.val && idx < a.idx)) { a.val = v; a.idx = idx; a.any = 1; }
}

__device__ __forceinline__ void wf_merge(WfBest& a, const WfBest& b)
{
    if (b.any) wf_take(a, b.val, b.idx);
    a.mir = fmaxf(a.mir, b.mir);
    a.sum += b.sum; a.sumsq += b.sumsq;
}

// shared-memory layout (host and device agree through this)
struct WfColumns { int sp, d, cost; };
template <int EY, int TY, int EX, int TX, int NW = 8, int XP = 1>
struct WfLayout {
    static constexpr int NY = EY * TY, NX = EX * TX, KP = NX / 2 + 1;
    using LX = LineFFT<EX, TX, false, XP>;
    using LY = LineFFT<EY, TY, true, XP>;
    // Row pitch SP of the spectra S[y][k] and the spacing D of the columns the lines of one warp take in the column
    // stage (a divisor of NW; warp w, line j -> column (w / D) D LPW + w % D + D j of the round): column accesses touch
    // word t SP + j D, row accesses k + j SP (stage C) and k + 2 j SP (stage A); the cheapest combination wins.
    static constexpr WfColumns columns()
    {
        WfColumns best{2 * KP, 1, 1 << 30};
        for (int sp = 2 * KP; sp < 2 * KP + 8; ++sp)
            for (int d = 1; d <= NW; ++d) {
                if (NW % d) continue;
                const int c = (4 * wf_cost(TY, sp, d) + 2 * wf_cost(TX, 1, sp) + wf_cost(TX, 1, 2 * sp)) * 1024 + (sp - 2 * KP) * 16 + d;
                if (c < best.cost) best = WfColumns{sp, d, c};
            }
        return best;
    }
    static constexpr WfColumns CC = columns();
    static constexpr int SP = CC.sp, D = CC.d;
    static constexpr int RW = (LX::LPW * LX::RS > LY::LPW * LY::RS ? LX::LPW * LX::RS : LY::LPW * LY::RS);   // region elements (complex) per warp
    static constexpr size_t BYTES = ((size_t)NY * SP + (size_t)NW * RW + NX + NY) * sizeof(cx<float>) + 64 * sizeof(WfBest) + 64;
};

struct WfParams {
    XcParams x;
    const cx<float>* twx;     // [EX][TX]: w_nx^(k1 t)
    const cx<float>* twy;     // [EY][TY]
};

// One line of stage C / D: surface row y (and, without the mirror term, row y2) from the conjugated spectra in S.
// Returns the lane's stage-B registers: Re = C[y][x], Im = -mirror[y][x] (MIRROR) or -C[y2][x].
template <typename LF, int E, int T, int SP, int KP, int PLANE>
__device__ __forceinline__ void wf_surface_line(const float* S, int y, int y2, bool second_is_q, cx<float>* u,
                                                float* region, const cx<float>* tw, int t)
{
    constexpr int N = E * T;
    const float* A = S + (size_t)y * SP;
    const float* B = y2 >= 0 ? S + (size_t)y2 * SP + (second_is_q ? KP : 0) : nullptr;
    cx<float> v[E];
#pragma unroll
    for (int n1 = 0; n1 < E; ++n1) {
        const int k = n1 * T + t;
        const bool direct = 2 * k <= N;
        const int kk = direct ? k : N - k;
        const cx<float> a = mk<float>(A[kk], A[PLANE + kk]);
        const cx<float> b = B ? mk<float>(B[kk], B[PLANE + kk]) : mk<float>(0.f, 0.f);
        // stored values are conj(P), conj(Q) (or conj(P) of two rows).  direct: conj(P + iQ) = a - i b (k = 0, N/2: real
        // parts only); mirrored index: conj(conj(P) + i conj(Q)) = conj(a) - i conj(b)
        v[n1] = direct ? ((k == 0 || 2 * k == N) ? mk<float>(a.x, -b.x) : mk<float>(a.x + b.y, a.y - b.x))
                       : mk<float>(a.x - b.y, -a.y - b.x);
    }
    LF::run(v, u, region, tw, t);
}

template <int EY, int TY, int EX, int TX, int NW, int XP, typename TI>
__device__ void kwf_pair(const WfParams& wp, unsigned char* smem)
{
    using L = WfLayout<EY, TY, EX, TX, NW, XP>;
    using LX = typename L::LX;
    using LY = typename L::LY;
    constexpr int NY = L::NY, NX = L::NX, KP = L::KP, SP = L::SP, NT = 32 * NW;
    const XcParams& p = wp.x;
    constexpr int PLANE = NY * SP;                                  // S: real parts [NY][SP], then imaginary parts
    float* S = reinterpret_cast<float*>(smem);
    float* regions = S + 2 * (size_t)PLANE;
    cx<float>* twx = reinterpret_cast<cx<float>*>(regions + 2 * (size_t)NW * L::RW);
    cx<float>* twy = twx + NX;
    WfBest* red = reinterpret_cast<WfBest*>(twy + NY);
    float* keep = reinterpret_cast<float*>(red + 34);              // c[3][3] of the sub-pixel fit
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mirror = p.conf_mode == CONF_MIRROR, want_std = p.conf_mode == CONF_STD;
    const int H0 = p.h0, W0 = p.w0, H1 = p.h1, W1 = p.w1;
    for (int i = tid; i < NX; i += NT) twx[i] = wp.twx[i];
    for (int i = tid; i < NY; i += NT) twy[i] = wp.twy[i];
    // lane coordinates for x lines and y lines
    const int lx = LX::lane_id(lane), tx = lx % TX, slotx = warp * LX::LPW + lx / TX;
    const int ly = LY::lane_id(lane), ty = ly % TY;
    const int sloty = (warp / L::D) * (L::D * LY::LPW) + warp % L::D + L::D * (ly / TY);     // columns of a warp are D apart
    float* regx = regions + 2 * ((size_t)warp * L::RW + (lx / TX) * LX::RS);
    float* regy = regions + 2 * ((size_t)warp * L::RW + (ly / TY) * LY::RS);
    constexpr int SLOTSX = NW * LX::LPW, SLOTSY = NW * LY::LPW;
    // the images are staged (as float) in the rows of S that stage A does not write, rows >= max(H0, H1): image 0 in the
    // plane of the real parts, image 1 in that of the imaginary parts
    const int hmax = H0 > H1 ? H0 : H1;
    float* stage0 = S + (size_t)hmax * SP;
    float* stage1 = stage0 + PLANE;
    const bool staged = (size_t)(NY - hmax) * SP >= (size_t)H0 * W0 && (size_t)(NY - hmax) * SP >= (size_t)H1 * W1;
    __syncthreads();
    for (int pair = blockIdx.x; pair < p.n; pair += gridDim.x) {
        const TI* img0 = reinterpret_cast<const TI*>(p.img0) + (size_t)pair * H0 * W0;
        const TI* img1 = reinterpret_cast<const TI*>(p.img1) + (size_t)pair * H1 * W1;
        if (tid < 2 && pair + (int)gridDim.x < p.n) {          // the CTA's next pair -> L2
            const size_t bytes = (size_t)(tid ? H1 * W1 : H0 * W0) * sizeof(TI);
            size_t a = reinterpret_cast<size_t>(tid ? img1 : img0) + (size_t)gridDim.x * bytes;
            size_t e = (a + bytes) & ~(size_t)15;
            a = (a + 15) & ~(size_t)15;
            if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
        }
        if (staged) {
            const int n0 = H0 * W0, n1 = H1 * W1;
            for (int i = tid; i < n0; i += NT) stage0[i] = (float)__ldg(img0 + i);
            for (int i = tid; i < n1; i += NT) stage1[i] = (float)__ldg(img1 + i);
            __syncthreads();
        }
        // ---- A: forward row transforms, two image rows per line
        {
            const int lines0 = (H0 + 1) / 2, lines1 = (H1 + 1) / 2, lines = lines0 + lines1;
            for (int base = 0; base < lines; base += SLOTSX) {
                int ln = base + slotx;
                const bool live = ln < lines;
                ln = live ? ln : lines - 1;
                const bool second = ln >= lines0;
                const int l = second ? ln - lines0 : ln;
                const int H = second ? H1 : H0, W = second ? W1 : W0;
                const int rA = 2 * l, rB = rA + 1;
                cx<float> v[EX], u[LX::U];
                if (staged) {
                    const float* im = second ? stage1 : stage0;
#pragma unroll
                    for (int n1 = 0; n1 < EX; ++n1) {
                        const int xx = n1 * TX + tx;
                        float a = 0.f, b = 0.f;
                        if (xx < W) { a = im[rA * W + xx]; if (rB < H) b = im[rB * W + xx]; }
                        v[n1] = mk<float>(a, b);
                    }
                } else {
                    const TI* im = second ? img1 : img0;
#pragma unroll
                    for (int n1 = 0; n1 < EX; ++n1) {
                        const int xx = n1 * TX + tx;
                        float a = 0.f, b = 0.f;
                        if (xx < W) { a = (float)__ldg(im + (size_t)rA * W + xx); if (rB < H) b = (float)__ldg(im + (size_t)rB * W + xx); }
                        v[n1] = mk<float>(a, b);
                    }
                }
                LX::run(v, u, regx, twx, tx);
                LX::for_each(u, tx, [&](int k, cx<float> z) { LX::put(regx, k, z); });
                __syncwarp();
                if (live) {
                    float* dA = S + (size_t)rA * SP + (second ? KP : 0);
                    float* dB = S + (size_t)rB * SP + (second ? KP : 0);
                    for (int k = tx; k < KP; k += TX) {
                        const cx<float> zk = LX::get(regx, k), zm = LX::get(regx, k ? NX - k : 0);
                        dA[k] = 0.5f * (zk.x + zm.x); dA[PLANE + k] = 0.5f * (zk.y - zm.y);                   // (Z[k] + conj Z[N-k]) / 2
                        if (rB < H) { dB[k] = 0.5f * (zk.y + zm.y); dB[PLANE + k] = 0.5f * (zm.x - zk.x); }   // (Z[k] - conj Z[N-k]) / 2i
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // ---- B: columns.  A line slot takes column k of BOTH spectra: forward transforms of F0[:, k] and F1[:, k]
        // (rows >= H are zero and not read), cross power(s) in registers -- a lane holds the same frequencies of both --
        // and the second transform(s) (forward transform of the conjugate = conjugate of the inverse), written back in
        // place: conj(P) column k over F0's, conj(Q) column k over F1's.  No other slot touches these two columns.
        for (int base = 0; base < KP; base += SLOTSY) {
            int c = base + sloty;
            const bool live = c < KP;
            c = live ? c : KP - 1;
            cx<float> v[EY], u0[LY::U], u1[LY::U];
#pragma unroll
            for (int n1 = 0; n1 < EY; ++n1) {
                const int y = n1 * TY + ty;
                // (idle slots of the last round transform zeros: reading the column another slot is rewriting in place would be
                // a -- harmless, the result is dropped -- data race)
                v[n1] = (live && y < H0) ? mk<float>(S[(size_t)y * SP + c], S[PLANE + (size_t)y * SP + c]) : mk<float>(0.f, 0.f);
            }
            LY::run(v, u0, regy, twy, ty);
#pragma unroll
            for (int n1 = 0; n1 < EY; ++n1) {
                const int y = n1 * TY + ty;
                v[n1] = (live && y < H1) ? mk<float>(S[(size_t)y * SP + KP + c], S[PLANE + (size_t)y * SP + KP + c]) : mk<float>(0.f, 0.f);
            }
            LY::run(v, u1, regy, twy, ty);
            // second transforms straight from the register layout of the first (LineFFT::run_rev): no exchange in between
#pragma unroll
            for (int i = 0; i < LY::U; ++i) {
                const cx<float> a = u0[i], b2 = u1[i];
                u0[i] = cmulc(a, b2);                                                    // conj(P) = F0 conj(F1)   (matcher.py:65)
                u1[i] = mk<float>(a.x * b2.x - a.y * b2.y, -(a.x * b2.y) - a.y * b2.x);  // conj(Q) = conj(F0 F1)   (matcher.py:114)
            }
            LY::run_rev(u0, v, regy, twy, ty);
            if (live) {
#pragma unroll
                for (int q1 = 0; q1 < EY; ++q1) {
                    const size_t o = (size_t)(ty + TY * q1) * SP + c;
                    S[o] = v[gpos<EY>(q1)].x; S[PLANE + o] = v[gpos<EY>(q1)].y;
                }
            }
            if (mirror) {
                LY::run_rev(u1, v, regy, twy, ty);
                if (live) {
#pragma unroll
                    for (int q1 = 0; q1 < EY; ++q1) {
                        const size_t o = (size_t)(ty + TY * q1) * SP + KP + c;
                        S[o] = v[gpos<EY>(q1)].x; S[PLANE + o] = v[gpos<EY>(q1)].y;
                    }
                }
            }
        }
        __syncthreads();
        // ---- C: inverse row transforms, maxima only
        WfBest acc;
        acc.val = 0.f; acc.idx = 0; acc.mir = 0.f; acc.sum = 0.0; acc.sumsq = 0.0; acc.any = 0;
        {
            const int lines = mirror ? NY : (NY + 1) / 2;
            for (int base = 0; base < lines; base += SLOTSX) {
                int ln = base + slotx;
                const bool live = ln < lines && lane < LX::AL;        // shadow lanes and idle slots do not vote
                ln = ln < lines ? ln : lines - 1;
                const int y = mirror ? ln : 2 * ln;
                const int y2 = mirror ? y : (y + 1 < NY ? y + 1 : -1);
                cx<float> u[LX::U];
                wf_surface_line<LX, EX, TX, SP, KP, PLANE>(S, y, y2, mirror, u, regx, twx, tx);
                if (live) {
                    LX::for_each(u, tx, [&](int x, cx<float> z) {
                        wf_take(acc, z.x, y * NX + x);
                        if (want_std) { acc.sum += (double)z.x; acc.sumsq += (double)z.x * (double)z.x; }
                        if (mirror) {
                            acc.mir = fmaxf(acc.mir, fabsf(z.y));
                        } else if (y2 >= 0) {
                            wf_take(acc, -z.y, y2 * NX + x);
                            if (want_std) { acc.sum -= (double)z.y; acc.sumsq += (double)z.y * (double)z.y; }
                        }
                    });
                }
            }
        }
        // block reduction (np.argmax order: largest value, lowest flat index)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            WfBest o;
            o.val = __shfl_xor_sync(0xffffffffu, acc.val, off);
            o.idx = __shfl_xor_sync(0xffffffffu, acc.idx, off);
            o.mir = __shfl_xor_sync(0xffffffffu, acc.mir, off);
            o.sum = __shfl_xor_sync(0xffffffffu, acc.sum, off);
            o.sumsq = __shfl_xor_sync(0xffffffffu, acc.sumsq, off);
            o.any = __shfl_xor_sync(0xffffffffu, acc.any, off);
            WfBest lo = (lane & off) ? o : acc, hi = (lane & off) ? acc : o;      // same summation order in both partners
            wf_merge(lo, hi);
            acc = lo;
        }
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            WfBest r = red[0];
            for (int w = 1; w < NW; ++w) wf_merge(r, red[w]);
            red[32] = r;
        }
        __syncthreads();
        const WfBest best = red[32];
        const int py = best.idx / NX, px = best.idx - py * NX;
        // ---- D: the three rows around the peak, one warp each
        if (p.subpixel && warp < 3) {
            int y = py - 1 + warp;
            y = y < 0 ? y + NY : (y >= NY ? y - NY : y);
            cx<float> u[LX::U];
            wf_surface_line<LX, EX, TX, SP, KP, PLANE>(S, y, -1, false, u, regx, twx, tx);
            LX::for_each(u, tx, [&](int x, cx<float> z) { regx[x] = z.x; });
            __syncwarp();
            if (lane == 0) {
                const int xm = px == 0 ? NX - 1 : px - 1, xp = px == NX - 1 ? 0 : px + 1;
                keep[warp * 3 + 0] = regx[xm]; keep[warp * 3 + 1] = regx[px]; keep[warp * 3 + 2] = regx[xp];
            }
        }
        __syncthreads();
        if (tid == 0) {
            float ox = 0.f, oy = 0.f;
            if (p.subpixel) {
#define FB_C(j, i) keep[((j) + 1) * 3 + (i) + 1]
                const float c00 = FB_C(0, 0);
                const float gx = (FB_C(0, 1) - FB_C(0, -1)) / 2.f;
                const float gy = (FB_C(1, 0) - FB_C(-1, 0)) / 2.f;
                const float hxx = FB_C(0, -1) + FB_C(0, 1) - 2.f * c00;
                const float hyy = FB_C(1, 0) + FB_C(-1, 0) - 2.f * c00;
                const float hxy = (FB_C(-1, -1) + FB_C(1, 1) - FB_C(-1, 1) - FB_C(1, -1)) / 4.f;
#undef FB_C
                const float det = hxx * hyy - hxy * hxy;
                if (det > 0.f) {
                    ox = -(hyy / det) * gx - (-hxy / det) * gy;
                    oy = -(-hxy / det) * gx - (hxx / det) * gy;
                }
                ox = ox < -0.5f ? -0.5f : (ox > 0.5f ? 0.5f : ox);
                oy = oy < -0.5f ? -0.5f : (oy > 0.5f ? 0.5f : oy);
            }
            double dx = (double)px + (double)ox + (double)(W0 - W1) / 2.0;
            double dy = (double)py + (double)oy + (double)(H0 - H1) / 2.0;
            dy -= rint(dy / NY) * NY;
            dx -= rint(dx / NX) * NX;
            double conf;
            if (p.conf_mode == CONF_NONE) {
                conf = 1.0;
            } else if (mirror) {
                conf = 0.0;
                if (best.val > 0.f) {
                    float c = 1.f - best.mir / best.val;
                    c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);
                    conf = (double)c;
                }
            } else {
                // the surface in S is unscaled: statistics of the scaled surface follow by linearity
                const double cnt = (double)NY * (double)NX, sc = p.scale;
                const double mean = best.sum * sc / cnt;
                const double var = best.sumsq * sc * sc / cnt - mean * mean;
                const float sd = (float)sqrt(var > 0.0 ? var : 0.0);
                const float r = (float)((double)best.val * sc) / sd;
                const float base = 1.f - (float)exp((double)(-r));
                double c = pow((double)base, cnt);
                if (c == c) c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
                conf = c;
            }
            p.dx[pair] = dx; p.dy[pair] = dy; p.conf[pair] = conf;
            if (p.peak) p.peak[pair] = (double)best.val * p.scale;
            if (p.mir) p.mir[pair] = (double)best.mir * p.scale;
        }
        __syncthreads();
    }
}

}  // namespace fb
