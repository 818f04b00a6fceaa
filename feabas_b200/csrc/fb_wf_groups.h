// fb_wf_groups.h -- size table and entry points of the warp-fused kernel (fb_xcorr_wf.cuh); compiled in several
// translation units (parallel build), dispatched from fb_xcorr.cu.
#pragma once
#include <cuda_runtime.h>

#include "fb_xcorr_wf.cuh"

// X(ny, nx, EY, TY, EX, TX, NW, XP): FFT grid, its line factorizations N = E * T (E points per lane, T lanes per line),
// the warps per CTA and the extra tile pitch the shared-memory budget allows (fb_xcorr_wf.cuh).  The grids of the
// shipped configurations:
//   150 x 135, 120 x 150, 120 x 135   stitching, finest level (blocks 74 x 67, 60 x 75, 60 x 67; pad)
//   75 x 72, 60 x 75                  the same without padding
//   100 x 100, 50 x 50                thumbnail alignment, 50 x 50 blocks with / without padding
//   128 x 128, 64 x 64                powers of two
#define FB_WF_SIZES_A(X) X(150, 135, 15, 10, 15, 9, 12, 1) X(75, 72, 15, 5, 12, 6, 8, 4) X(64, 64, 8, 8, 8, 8, 8, 4)
#define FB_WF_SIZES_B(X) X(120, 150, 15, 8, 15, 10, 12, 5) X(120, 135, 15, 8, 15, 9, 12, 5) X(60, 75, 15, 4, 15, 5, 8, 4)
#define FB_WF_SIZES_C(X) X(100, 100, 10, 10, 10, 10, 8, 1) X(50, 50, 10, 5, 10, 5, 8, 4) X(128, 128, 16, 8, 16, 8, 16, 1)
#define FB_WF_SIZES(X) FB_WF_SIZES_A(X) FB_WF_SIZES_B(X) FB_WF_SIZES_C(X)

namespace fb {

struct WfLaunch {
    int ny, nx;
    int in_dtype;          // FB_F32 / FB_U8 (uint8 pixels converted to float on load)
    int grid;              // 0: only report the shape (threads, smem)
    cudaStream_t stream;
    int threads;           // out
    size_t smem;           // out
};

// each returns false when the grid is not in the group
#define FB_WF_GROUP_DECL(G)                              \
    int wf_set_attrs_##G(size_t max_smem);               \
    bool wf_launch_##G(const WfParams& wp, WfLaunch& l);
FB_WF_GROUP_DECL(a)
FB_WF_GROUP_DECL(b)
FB_WF_GROUP_DECL(c)

}  // namespace fb
