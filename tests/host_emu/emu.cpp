// TEST INFRASTRUCTURE: executes the CUDA kernel bodies of feabas_b200/csrc/fb_xcorr.cuh
// on the CPU with a small team of OS threads standing in for a CTA.  Lets the CPU-only
// test tier check the kernels' index arithmetic against the oracle without a GPU.
// Never loaded by the product (feabas_b200/cuda/_lib.py only loads libfeabas_cuda.so).
#include <math.h>
#include <barrier>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../feabas_b200/csrc/fb_xcorr.cuh"
#include "../../feabas_b200/csrc/fb_host_plan.h"

namespace fb {
static thread_local std::barrier<>* tl_barrier = nullptr;
void emu_sync() { tl_barrier->arrive_and_wait(); }
}  // namespace fb

using namespace fb;

static void run_grid(int nblocks, int nthr, size_t smem_bytes,
                     const std::function<void(int, int, int, unsigned char*)>& body)
{
    std::vector<unsigned char> smem(smem_bytes + 64);
    std::barrier<> bar(nthr);
    std::vector<std::thread> team;
    for (int t = 0; t < nthr; ++t) {
        team.emplace_back([&, t]() {
            tl_barrier = &bar;
            for (int b = 0; b < nblocks; ++b) {
                body(b, t, nthr, smem.data());
                bar.arrive_and_wait();
            }
        });
    }
    for (auto& th : team) th.join();
}

template <typename T> struct HostPlan {
    std::vector<cx<T>> tw;
    std::vector<int> pos;
    Plan1D p;
    HostPlan(int n, bool wide)
    {
        auto r = wide ? radix_sequence(n) : radix_sequence_basic(n);
        pos = digit_positions(n, r);
        tw = twiddle_table<T>(n);
        p.n = n; p.npass = (int)r.size();
        for (size_t i = 0; i < r.size(); ++i) p.radix[i] = r[i];
        p.tw = tw.data(); p.pos = pos.data();
    }
};

template <typename T, typename TI>
static int run(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1, int ny, int nx,
               int conf_mode, int subpixel, int path, double* out, int nthr)
{
    if (!is_5smooth(ny) || !is_5smooth(nx)) return -2;
    Geometry g{};
    g.h0 = h0; g.w0 = w0; g.h1 = h1; g.w1 = w1; g.ny = ny; g.nx = nx;
    g.esize = (int)sizeof(cx<T>); g.mirror = conf_mode == CONF_MIRROR;
    if (!choose_tiles(g)) return -3;
    const bool fused = path == 1 || (path == 0 && g.fused);
    const bool narrow = path == 3;                 // staged pipeline, K4 recomputing one surface row at a time
    HostPlan<T> px(nx, fused), py(ny, fused);
    XcParams p{};
    p.img0 = img0; p.img1 = img1; p.n = n; p.h0 = h0; p.w0 = w0; p.h1 = h1; p.w1 = w1;
    p.ny = ny; p.nx = nx; p.kp = g.kp; p.px = px.p; p.py = py.p; p.fpitch = g.fpitch;
    p.dx = out; p.dy = out + n; p.conf = out + 2 * n; p.peak = out + 3 * n; p.mir = out + 4 * n; p.conf_mode = conf_mode; p.subpixel = subpixel; p.scale = 1.0 / ((double)ny * nx);
    if (fused) {
        if (!g.fused) return -4;
        p.tl = g.tl_fused; p.spitch = g.spitch;
        run_grid(n, nthr, g.smem_fused, [&](int b, int t, int nt, unsigned char* sm) { kf_fused<T, TI>(p, b, t, nt, sm); });
        return 1;
    }
    if (!g.tl_row) return -5;
    p.tl = g.tl_row; p.tc = g.tc_col;
    std::vector<cx<T>> F0((size_t)n * h0 * g.fpitch), F1((size_t)n * h1 * g.fpitch), G((size_t)n * ny * 2 * g.fpitch);
    const bool mirror = g.mirror;
    const int rpt = mirror ? p.tl : 2 * p.tl;
    p.nrt = (ny + rpt - 1) / rpt;
    std::vector<Partial> part((size_t)n * p.nrt);
    p.F0 = F0.data(); p.F1 = F1.data(); p.G = G.data(); p.part = part.data();
    int t0 = row_tiles<T>(h0, p.tl), t1 = row_tiles<T>(h1, p.tl);
    run_grid(n * (t0 + t1), nthr, g.smem_row, [&](int b, int t, int nt, unsigned char* sm) { k1_rows_forward<T, TI>(p, b, t, nt, sm); });
    int nct = (g.kp + p.tc - 1) / p.tc;
    run_grid(n * nct, nthr, g.smem_col, [&](int b, int t, int nt, unsigned char* sm) { k2_columns<T>(p, b, t, nt, sm); });
    run_grid(n * p.nrt, nthr, g.smem_row, [&](int b, int t, int nt, unsigned char* sm) { k3_rows_inverse<T>(p, b, t, nt, sm); });
    p.fin_narrow = narrow ? 1 : 0;
    size_t sm4 = (size_t)nx * (narrow ? 1 : 4) * sizeof(cx<T>) + 2048;
    run_grid(n, nthr, sm4, [&](int b, int t, int nt, unsigned char* sm) { k4_finalize<T>(p, b, t, nt, sm); });
    return narrow ? 3 : 2;
}

extern "C" int emu_xcorr(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1, int dtype,
                         int ny, int nx, int conf_mode, int subpixel, int path, double* out, int nthr)
{
    if (nthr < 1 || nthr > 32) return -1;
    switch (dtype) {
        case 0: return run<float, float>(img0, img1, n, h0, w0, h1, w1, ny, nx, conf_mode, subpixel, path, out, nthr);
        case 1: return run<double, unsigned char>(img0, img1, n, h0, w0, h1, w1, ny, nx, conf_mode, subpixel, path, out, nthr);
        case 2: return run<double, double>(img0, img1, n, h0, w0, h1, w1, ny, nx, conf_mode, subpixel, path, out, nthr);
    }
    return -1;
}
