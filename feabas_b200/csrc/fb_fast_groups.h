// fb_fast_groups.h -- the register-resident fast path is compiled in several translation units (one per
// group of line lengths) so that the library builds in parallel; fb_xcorr.cu dispatches through these entry points.
#pragma once
#include <cuda_runtime.h>

#include "fb_xcorr_fast.cuh"

// fast-path line lengths: N = E * T (E points per lane, T lanes per line).  X(n, E, T)
#define FB_FAST_SIZES_POW2(X) X(256, 16, 16) X(512, 32, 16) X(1024, 32, 32)
#define FB_FAST_SIZES_BIG(X) X(2048, 64, 32) X(4096, 64, 64)
#define FB_FAST_SIZES_R3(X) X(576, 24, 24) X(288, 24, 12) X(384, 48, 8) X(768, 48, 16) X(1152, 48, 24)
#define FB_FAST_SIZES_R5(X) X(300, 30, 10) X(200, 20, 10) X(400, 40, 10) X(800, 40, 20)
#define FB_FAST_SIZES_R5B(X) X(600, 60, 10) X(500, 50, 10) X(1200, 60, 20) X(720, 60, 12)
#define FB_FAST_SIZES(X) FB_FAST_SIZES_POW2(X) FB_FAST_SIZES_BIG(X) FB_FAST_SIZES_R3(X) FB_FAST_SIZES_R5(X) FB_FAST_SIZES_R5B(X)

namespace fb {

constexpr int kNW1 = 8, kNW2 = 8;                    // warps per CTA of K1 / K2
constexpr int kNW2S = 8;                            // warps per CTA of the solo column kernel (one CTA per SM)
constexpr int kSoloRegs = 248;                       // registers per thread it may use: 65536 / (32 kNW2S), a multiple of 8
// the solo column kernel exists for lines inside one warp with at most 32 points per lane
template <int E, int T> constexpr __host__ __device__ bool kSoloOk() { return E <= 32 && T <= 32; }
// rows per GT tile = lines per K3 CTA: 8 when 8 divides the line length (square grids: the row count), else 4;
// 4 for the two-warp lines of 4096 (shared memory)
template <int E, int T> constexpr int kR3() { return T > 32 || (E * T) % 8 ? 4 : 8; }
// CTAs per SM the register budget allows: a lane holds E complex points (E > 32: one CTA)
template <int E> constexpr int kOcc(int full) { return E > 32 ? 1 : full; }
// K3 CTAs per SM: 512 threads per SM for E <= 32; for E > 32 (up to 255 registers per thread) what the register file holds
constexpr __host__ __device__ int k3_ctas_per_sm(int E, int threads)
{
    const int full = 512 / threads;
    if (E <= 32) return full;
    const int byreg = 65536 / (threads * 255);
    return byreg < 1 ? 1 : (byreg < full ? byreg : full);
}

// K3 variants: 0 = TMA-fed (default), 1 = LDG, 2 = LDG with 4 lines per CTA on 4-row tiles, 3 = same on 8-row tiles
// (2, 3: experiment switches, 1024-point lines only)
struct FastLaunch {
    int n;                 // line length of this stage
    int in_dtype;          // K1: FB_F32 / FB_U8
    bool pruned;           // K1 / K2
    bool mirror;           // K3
    int k3_variant;
    bool k2_solo;          // K2: the one-line-per-column-pair kernel
    int grid, threads;
    size_t smem;
    cudaStream_t stream;
};

// each returns false when the line length is not in the group
#define FB_FAST_GROUP_DECL(G)                                           \
    int fast_set_attrs_##G(size_t max_smem);                            \
    bool fast_k1_##G(const FastParams& fp, const FastLaunch& l);        \
    bool fast_k2_##G(const FastParams& fp, const FastLaunch& l);        \
    bool fast_k3_##G(const FastParams& fp, const FastLaunch& l);
FB_FAST_GROUP_DECL(pow2)
FB_FAST_GROUP_DECL(big)
FB_FAST_GROUP_DECL(r3)
FB_FAST_GROUP_DECL(r5)
FB_FAST_GROUP_DECL(r5b)

}  // namespace fb
