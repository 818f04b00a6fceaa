// fb_xcorr.cu -- libfeabas_cuda.so: __global__ wrappers, plan / workspace caches and the
// C ABI declared in include/feabas_cuda.h.  sm_100a only; no CPU fallback.
#include <cuda_runtime.h>
#include <math.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../include/feabas_cuda.h"
#include "fb_host_plan.h"
#include "fb_xcorr.cuh"
#include "fb_fast_groups.h"
#include "fb_wf_groups.h"

using namespace fb;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_pairs{0};          // block pairs handed to the xcorr kernels (all entry points)

static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// shared with the other translation units of the library (fb_common.h)
int fb_set_error(int code, const char* msg) { g_err = msg; return code; }
void fb_count_launches(int n) { g_launches += n; }

#define CU(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail(FB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <typename T, typename TI>
__global__ void __launch_bounds__(512) fbk_rows_forward(const __grid_constant__ XcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    k1_rows_forward<T, TI>(p, blockIdx.x, threadIdx.x, blockDim.x, smem);
}
template <typename T>
__global__ void __launch_bounds__(512) fbk_columns(const __grid_constant__ XcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    k2_columns<T>(p, blockIdx.x, threadIdx.x, blockDim.x, smem);
}
template <typename T>
__global__ void __launch_bounds__(512) fbk_rows_inverse(const __grid_constant__ XcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    k3_rows_inverse<T>(p, blockIdx.x, threadIdx.x, blockDim.x, smem);
}
template <typename T>
__global__ void __launch_bounds__(256) fbk_finalize(const __grid_constant__ XcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    k4_finalize<T>(p, blockIdx.x, threadIdx.x, blockDim.x, smem);
}
template <typename T, typename TI>
__global__ void __launch_bounds__(512) fbk_fused(const __grid_constant__ XcParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    kf_fused<T, TI>(p, blockIdx.x, threadIdx.x, blockDim.x, smem);
}


// multi-channel stacks (matcher.py:66-67,115-116): the cross-power is averaged over channels before the
// inverse transform; the column stage is linear, so the mean is taken over its per-channel outputs
template <typename T>
__global__ void __launch_bounds__(256) fbk_channel_mean(const cx<T>* __restrict__ g, cx<T>* __restrict__ out, int nchan,
                                                         size_t elems, size_t total, int pitch, int kp)
{
    const T inv = T(1) / T(nchan);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pair = i / elems, e = i - pair * elems;
        if ((int)(e % (size_t)pitch) >= kp) continue;      // row padding: never written, never read downstream
        const cx<T>* src = g + (pair * nchan) * elems + e;
        cx<T> a = src[0];
        for (int c = 1; c < nchan; ++c) a = a + src[(size_t)c * elems];
        out[i] = mk<T>(a.x * inv, a.y * inv);
    }
}

// ---- register-resident fast path: kernels live in fb_fast_<group>.cu (fb_fast_groups.h) ----

// ---------------------------------------------------------------------------
// caches
// ---------------------------------------------------------------------------
struct DevTables {
    void* tw = nullptr;
    int* pos = nullptr;
    std::vector<int> radix;
};

struct GtMapKey {
    void* gt; int nb, ny, kp, rblk;
    bool operator<(const GtMapKey& o) const { return std::tie(gt, nb, ny, kp, rblk) < std::tie(o.gt, o.nb, o.ny, o.kp, o.rblk); }
};

struct StreamCtx {
    std::mutex mu;                 // serialises the users of THIS context (its workspace and staging buffers); contexts of
                                   // other devices / streams run concurrently
    int device = 0;
    void* ws = nullptr;            // staged-pipeline workspace
    std::map<GtMapKey, CUtensorMap> gt_maps;     // encoded tensor maps of the workspace's G^T region (dropped with the workspace)
    size_t ws_bytes = 0;
    // host path
    cudaStream_t copy_stream = nullptr, own_stream = nullptr;
    void* din[2] = {nullptr, nullptr};
    size_t din_bytes = 0;
    void* pin[2] = {nullptr, nullptr};
    size_t pin_bytes = 0;
    double* dout = nullptr;
    size_t dout_bytes = 0;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // fast path, pipelined schedule: side stream + ordering events (no timing)
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> order_ev;
    // per-kernel timing (option "profile"): (slot, start, stop) triples not yet read back
    std::vector<std::tuple<int, cudaEvent_t, cudaEvent_t>> prof_pending;
    std::vector<cudaEvent_t> prof_free;
    double prof_ms[5] = {0, 0, 0, 0, 0};
    long long prof_n[5] = {0, 0, 0, 0, 0};
};

// Locking: g_mu guards only the process-wide caches below (tables, the context map, attribute flags) and is never held
// across a copy, a launch sequence or a synchronisation; the work of a call runs under its StreamCtx::mu.  Host threads
// driving different devices (or different streams of one device) therefore overlap (feabas_b200/cuda/shard.py).
static std::mutex g_mu;
static std::map<std::tuple<int, int, int>, DevTables> g_tables;      // (device, n, is_double + 2 * wide radices)
static std::map<std::pair<int, void*>, StreamCtx> g_ctx;              // (device, stream the work runs on)
static std::map<int, bool> g_attr_done;
static void* const kOwnStreamKey = reinterpret_cast<void*>(~(uintptr_t)0);   // context of the host path's private stream
static std::atomic<long long> g_opt_ws_bytes{8LL << 30};
static std::atomic<long long> g_opt_host_chunk{64LL << 20};
static std::atomic<long long> g_opt_profile{0};
static std::atomic<long long> g_opt_fused_threads{0};      // experiment switch: threads per CTA of the fused kernel (0: 512 above 100 KB, else 256)
static std::atomic<long long> g_opt_pipeline_waves{1};     // ... when every kernel of a half still has this many work items per resident CTA
static std::atomic<long long> g_opt_pipeline{2};           // fast path: a chunk runs as this many independent parts on separate streams (1: serial)
static std::atomic<long long> g_opt_max_radix{16};         // largest radix of the shared-memory passes (experiment switch; set before first use)
static std::atomic<long long> g_opt_warp_fused{1};         // 0: small grids run on the first fused kernel (shared-memory radix passes)
static std::atomic<long long> g_opt_fast_flags{0};       // experiment switches, see FastParams::flags (+16: K2 unbatched twiddles, +32: K3 8 lines)

template <typename T>
static int get_tables(int device, int n, Plan1D& out, bool wide = false)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(device, n, (int)(sizeof(T) == 8) + (wide ? 2 : 0));
    auto it = g_tables.find(key);
    if (it == g_tables.end()) {
        DevTables t;
        t.radix = wide ? radix_sequence(n, (int)g_opt_max_radix.load()) : radix_sequence_basic(n);
        if ((int)t.radix.size() > kMaxPass) return fail(FB_ESIZE, "fft length %d needs too many passes", n);
        auto pos = digit_positions(n, t.radix);
        auto tw = twiddle_table<T>(n);
        CU(cudaMalloc(&t.tw, tw.size() * sizeof(cx<T>)));
        CU(cudaMalloc((void**)&t.pos, pos.size() * sizeof(int)));
        CU(cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(t.pos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
        it = g_tables.emplace(key, t).first;
    }
    const DevTables& t = it->second;
    out.n = n;
    out.npass = (int)t.radix.size();
    for (int i = 0; i < out.npass; ++i) out.radix[i] = t.radix[i];
    out.tw = t.tw;
    out.pos = t.pos;
    return FB_OK;
}

static std::map<std::tuple<int, int, int>, void*> g_wtables;           // (device, n, T) -> cx<float>[E / 2][T][2]

static int get_warp_table(int device, int n, int T, const cx<float>*& out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(device, n, T);
    auto it = g_wtables.find(key);
    if (it == g_wtables.end()) {
        const int E = n / T;
        std::vector<cx<float>> tw((size_t)n);
        for (int k1 = 0; k1 < E; ++k1)
            for (int t = 0; t < T; ++t) {
                long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)((k1 * t) % n) / (long double)n;
                const size_t slot = ((size_t)(k1 / 2) * T + t) * 2 + (k1 & 1);     // rows (k1, k1 + 1) paired per lane
                tw[slot].x = (float)std::cos(a);
                tw[slot].y = (float)std::sin(a);
            }
        void* d = nullptr;
        CU(cudaMalloc(&d, tw.size() * sizeof(cx<float>)));
        CU(cudaMemcpy(d, tw.data(), tw.size() * sizeof(cx<float>), cudaMemcpyHostToDevice));
        it = g_wtables.emplace(key, d).first;
    }
    out = reinterpret_cast<const cx<float>*>(it->second);
    return FB_OK;
}

// warp-fused kernel: plain [E][T] table of w_n^(k1 t)
static int get_wf_table(int device, int n, int T, const cx<float>*& out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(device, n, -T);
    auto it = g_wtables.find(key);
    if (it == g_wtables.end()) {
        const int E = n / T;
        std::vector<cx<float>> tw((size_t)n);
        for (int k1 = 0; k1 < E; ++k1)
            for (int t = 0; t < T; ++t) {
                long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)((k1 * t) % n) / (long double)n;
                tw[(size_t)k1 * T + t].x = (float)std::cos(a);
                tw[(size_t)k1 * T + t].y = (float)std::sin(a);
            }
        void* d = nullptr;
        CU(cudaMalloc(&d, tw.size() * sizeof(cx<float>)));
        CU(cudaMemcpy(d, tw.data(), tw.size() * sizeof(cx<float>), cudaMemcpyHostToDevice));
        it = g_wtables.emplace(key, d).first;
    }
    out = reinterpret_cast<const cx<float>*>(it->second);
    return FB_OK;
}

static int wf_lanes(int ny, int nx, int& ty, int& tx)
{
    ty = tx = 0;
#define X(NY_, NX_, EY_, TY_, EX_, TX_, NW_, XP_) if (ny == NY_ && nx == NX_) { ty = TY_; tx = TX_; }
    FB_WF_SIZES(X)
#undef X
    return ty && tx;
}

template <typename F>
static int raise_smem(F* fn)
{
    CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    return FB_OK;
}

static int set_attrs(int device)
{
    if (g_attr_done[device]) return FB_OK;
    int rc;
#define RS(k) if ((rc = raise_smem(k)) != FB_OK) return rc
    (void)0;
    RS((fbk_rows_forward<float, float>));
    RS((fbk_rows_forward<float, unsigned char>));
    RS((fbk_rows_forward<double, unsigned char>));
    RS((fbk_rows_forward<double, double>));
    RS((fbk_columns<float>));
    RS((fbk_columns<double>));
    RS((fbk_rows_inverse<float>));
    RS((fbk_rows_inverse<double>));
    RS((fbk_finalize<float>));
    RS((fbk_finalize<double>));
    RS((fbk_fused<float, float>));
    RS((fbk_fused<float, unsigned char>));
    RS((fbk_fused<double, unsigned char>));
    RS((fbk_fused<double, double>));
    if (fast_set_attrs_pow2(kMaxSmem) || fast_set_attrs_big(kMaxSmem) || fast_set_attrs_r3(kMaxSmem) || fast_set_attrs_r5(kMaxSmem) ||
        fast_set_attrs_r5b(kMaxSmem))
        return fail(FB_ECUDA, "cudaFuncSetAttribute failed for a fast-path kernel");
    if (wf_set_attrs_a(kMaxSmem) || wf_set_attrs_b(kMaxSmem) || wf_set_attrs_c(kMaxSmem))
        return fail(FB_ECUDA, "cudaFuncSetAttribute failed for a warp-fused kernel");
#undef RS
    g_attr_done[device] = true;
    return FB_OK;
}

// ---------------------------------------------------------------------------
// geometry / validation
// ---------------------------------------------------------------------------
struct Problem {
    int n, h0, w0, h1, w1, in_dtype, ny, nx, flags;
    int conf_mode, subpixel;
    bool f64;       // compute type
    int isz;        // bytes per input element
    Geometry g;
    bool fused;
    bool wf;        // fused, warp-per-line register transforms (fb_xcorr_wf.cuh)
    bool fast;      // register-resident power-of-two pipeline
    int hp0, hp1;   // fast: padded heights of the transposed row spectra
    size_t ws_per_pair;
    int nrt;
    int nchan;      // channels per image (cross-power averaged); > 1 or any fb_xcorr_ext pointer -> generic staged path
    fb_xcorr_ext ext;
};

static bool wf_shape(int ny, int nx, int& threads, size_t& smem)
{
    WfLaunch l{ny, nx, FB_F32, 0, nullptr, 0, 0};
    WfParams none{};
    if (!(wf_launch_a(none, l) || wf_launch_b(none, l) || wf_launch_c(none, l))) return false;
    threads = l.threads; smem = l.smem;
    return true;
}

static int fast_rblk_of(int nx);
static bool fast_size(int n)
{
#define X(N_, E_, T_) if (n == N_) return true;
    FB_FAST_SIZES(X)
#undef X
    return false;
}

static int make_problem(Problem& q, int n, int h0, int w0, int h1, int w1, int in_dtype, int fft_h, int fft_w, int flags,
                        const fb_xcorr_ext* ext = nullptr)
{
    q.ext = fb_xcorr_ext{};
    if (ext) q.ext = *ext;
    q.nchan = q.ext.nchan > 1 ? q.ext.nchan : 1;
    const bool ext_active = q.nchan > 1 || q.ext.norm || q.ext.norm_mirror || q.ext.surface || q.ext.surface_mirror;
    if (ext && q.ext.nchan < 0) return fail(FB_EINVAL, "bad nchan %d", q.ext.nchan);
    if (ext_active) flags = (flags & ~FB_FLAG_FORCE_FUSED) | FB_FLAG_FORCE_STAGED | FB_FLAG_FORCE_GENERIC;
    if (n < 0 || h0 < 1 || w0 < 1 || h1 < 1 || w1 < 1) return fail(FB_EINVAL, "bad shape n=%d %dx%d / %dx%d", n, h0, w0, h1, w1);
    if (in_dtype < FB_F32 || in_dtype > FB_F64) return fail(FB_EINVAL, "bad in_dtype %d", in_dtype);
    if (fft_h < (h0 > h1 ? h0 : h1) || fft_w < (w0 > w1 ? w0 : w1)) return fail(FB_EINVAL, "fft grid %dx%d smaller than the images", fft_h, fft_w);
    if (!is_5smooth(fft_h) || !is_5smooth(fft_w)) return fail(FB_ESIZE, "fft grid %dx%d is not 2^a 3^b 5^c", fft_h, fft_w);
    if ((long long)fft_h * fft_w >= (1LL << 31)) return fail(FB_ESIZE, "fft grid %dx%d too large", fft_h, fft_w);
    q.n = n; q.h0 = h0; q.w0 = w0; q.h1 = h1; q.w1 = w1; q.in_dtype = in_dtype; q.ny = fft_h; q.nx = fft_w; q.flags = flags;
    q.conf_mode = (flags >> FB_CONF_SHIFT) & 3;
    if (q.conf_mode > 2) return fail(FB_EINVAL, "bad conf_mode %d", q.conf_mode);
    q.subpixel = (flags & FB_FLAG_SUBPIXEL) ? 1 : 0;
    q.f64 = in_dtype == FB_F64 || (in_dtype == FB_U8 && !(flags & FB_FLAG_U8_AS_F32));
    q.isz = in_dtype == FB_F32 ? 4 : (in_dtype == FB_U8 ? 1 : 8);
    Geometry& g = q.g;
    g = Geometry{};
    g.h0 = h0; g.w0 = w0; g.h1 = h1; g.w1 = w1; g.ny = fft_h; g.nx = fft_w;
    g.esize = q.f64 ? 16 : 8;
    g.mirror = q.conf_mode == CONF_MIRROR;
    if (!choose_tiles(g)) return fail(FB_ESIZE, "fft grid %dx%d does not fit the kernels' shared-memory tiling", fft_h, fft_w);
    q.fused = g.fused;
    if (flags & FB_FLAG_FORCE_STAGED) {
        if (!g.tl_row || !g.tc_col) return fail(FB_ESIZE, "staged path unavailable for %dx%d", fft_h, fft_w);
        q.fused = false;
    }
    if (flags & FB_FLAG_FORCE_FUSED) {
        if (!g.fused) return fail(FB_ESIZE, "fused path unavailable for %dx%d", fft_h, fft_w);
        q.fused = true;
    }
    {
        int wt; size_t ws;
        q.wf = !q.f64 && !ext_active && !(flags & (FB_FLAG_FORCE_STAGED | FB_FLAG_FORCE_FUSED_SMEM)) && g_opt_warp_fused && wf_shape(fft_h, fft_w, wt, ws);
        if (q.wf) q.fused = true;
    }
    const int rpt = g.mirror ? g.tl_row : 2 * g.tl_row;
    q.nrt = q.fused ? 0 : (fft_h + rpt - 1) / rpt;
    q.ws_per_pair = q.fused ? 0
                            : ((size_t)(h0 + h1) * g.fpitch * q.nchan + (size_t)fft_h * 2 * g.fpitch * (q.nchan + (q.nchan > 1 ? 1 : 0))) * g.esize +
                                  (size_t)q.nrt * sizeof(Partial);
    q.fast = !q.fused && !q.f64 && fast_size(fft_h) && fast_size(fft_w) && !(flags & FB_FLAG_FORCE_GENERIC);
    if (q.fast) {
        // K3 owns whole GT tiles of rblk rows (rblk follows the x size); without the mirror term a line is two rows
        const int rb = fast_rblk_of(fft_w);
        if (fft_h % (g.mirror ? rb : 2 * rb)) q.fast = false;
    }
    if (q.fast) {
        q.hp0 = (h0 + 31) & ~31; q.hp1 = (h1 + 31) & ~31;
        q.nrt = g.mirror ? fft_h : (fft_h + 1) / 2;
        q.ws_per_pair = ((size_t)g.kp * (q.hp0 + q.hp1) + (size_t)2 * g.kp * fft_h) * 8 + (size_t)q.nrt * sizeof(Partial);
    }
    q.ws_per_pair = (q.ws_per_pair + 255) & ~(size_t)255;
    return FB_OK;
}


// ---------------------------------------------------------------------------
// optional per-kernel CUDA-event timing on the launching stream
// ---------------------------------------------------------------------------
enum { SLOT_ROWS_FWD = 0, SLOT_COLUMNS = 1, SLOT_ROWS_INV = 2, SLOT_FINALIZE = 3, SLOT_FUSED = 4 };

static cudaEvent_t prof_event(StreamCtx& c)
{
    cudaEvent_t e = nullptr;
    if (!c.prof_free.empty()) { e = c.prof_free.back(); c.prof_free.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    StreamCtx& c; cudaStream_t st; int slot; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(StreamCtx& c_, cudaStream_t st_, int slot_) : c(c_), st(st_), slot(slot_)
    {
        if (g_opt_profile) { a = prof_event(c); b = prof_event(c); cudaEventRecord(a, st); }
    }
    ~ProfScope()
    {
        if (a) { cudaEventRecord(b, st); c.prof_pending.emplace_back(slot, a, b); }
    }
};


// ---------------------------------------------------------------------------
// fast path launch
// ---------------------------------------------------------------------------
static bool fast_dispatch(int stage, const FastParams& fp, const FastLaunch& l)
{
    switch (stage) {
        case 1: return fast_k1_pow2(fp, l) || fast_k1_big(fp, l) || fast_k1_r3(fp, l) || fast_k1_r5(fp, l) || fast_k1_r5b(fp, l);
        case 2: return fast_k2_pow2(fp, l) || fast_k2_big(fp, l) || fast_k2_r3(fp, l) || fast_k2_r5(fp, l) || fast_k2_r5b(fp, l);
        default: return fast_k3_pow2(fp, l) || fast_k3_big(fp, l) || fast_k3_r3(fp, l) || fast_k3_r5(fp, l) || fast_k3_r5b(fp, l);
    }
}
static void fast_et(int n, int& E, int& T)
{
    E = T = 0;
#define X(N_, E_, T_) if (n == N_) { E = E_; T = T_; }
    FB_FAST_SIZES(X)
#undef X
}
static int fast_rblk(int n, int T) { return T > 32 || n % 8 ? 4 : 8; }                    // == kR3<E, T>()
static int fast_lines(int T, int nw) { return T > 32 ? nw / (T / 32) : nw * (32 / T); }   // WarpFFT::lines_per_cta
static int fast_rblk_of(int nx) { int E, T; fast_et(nx, E, T); return fast_rblk(nx, T); }
static size_t fast_smem(int n, int nw)
{
    int E, T; fast_et(n, E, T);
    const int lines = fast_lines(T, nw);
    return ((size_t)lines * ((wfft_region(E, T) + 15 + 16) & ~15) + (kLaneTwiddles ? 0 : n)) * sizeof(cx<float>);
}

static std::atomic<int> g_num_sms{0};
static std::atomic<int> g_opt_k2_solo{0};      // option "k2_solo": column stage on the one-line-per-column-pair kernel

// shared memory of the solo column kernel: two transpose regions per line + the stage-twiddle table
static size_t fast_smem_solo(int n)
{
    int E, T; fast_et(n, E, T);
    const int lines = fast_lines(T, kNW2S);
    return ((size_t)lines * 2 * ((wfft_region(E, T) + 15) & ~15) + n) * sizeof(cx<float>);
}
static bool fast_k2_solo(const Problem& q)
{
    int E, T; fast_et(q.ny, E, T);
    return g_opt_k2_solo && E && E <= 32 && T <= 32 && fast_smem_solo(q.ny) <= kMaxSmem;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static std::once_flag g_encode_once;

// G^T of `nb` pairs as a 4-D tensor of 8-byte elements, innermost first: [rblk][kp][2][nb * ny / rblk];
// box = one column of one plane of one pair: [rblk][1][1][ny / rblk] = ny elements in natural y order
static bool make_gt_map(StreamCtx& ctx, CUtensorMap* map, void* gt, int nb, int ny, int kp, int rblk)
{
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
        cudaGetLastError();
    });
    if (!g_encode_tiled) return false;
    // the matcher layers send the same few batch shapes over and over: encode once per (address, shape)
    const GtMapKey key{gt, nb, ny, kp, rblk};
    auto hit = ctx.gt_maps.find(key);
    if (hit != ctx.gt_maps.end()) { *map = hit->second; return true; }
    const int tiles = ny / rblk, boxt = tiles > 256 ? 256 : tiles;      // a box dimension is at most 256: K2 stores long columns in pieces
    if (tiles % boxt) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)rblk, (cuuint64_t)kp, 2, (cuuint64_t)nb * (ny / rblk)};
    const cuuint64_t strides[3] = {(cuuint64_t)rblk * 8, (cuuint64_t)kp * rblk * 8, (cuuint64_t)2 * kp * rblk * 8};
    const cuuint32_t box[4] = {(cuuint32_t)rblk, 1, 1, (cuuint32_t)boxt};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, gt, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (ctx.gt_maps.size() > 256) ctx.gt_maps.clear();
    ctx.gt_maps.emplace(key, *map);
    return true;
}

// Parameters of the sub-range [lo, lo + cnt) of a chunk of nb pairs (workspace carved by pair index).
static int fast_prepare(const Problem& q, StreamCtx& ctx, const XcParams& p, int lo, int cnt, int nb, FastParams& fp)
{
    int rc;
    int EX, TX, EY, TY;
    fast_et(q.nx, EX, TX); fast_et(q.ny, EY, TY);
    if ((rc = get_warp_table(ctx.device, q.nx, TX, fp.twx)) != FB_OK) return rc;
    if ((rc = get_warp_table(ctx.device, q.ny, TY, fp.twy)) != FB_OK) return rc;
    const Geometry& g = q.g;
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx.ws);
    const size_t f0 = (size_t)nb * g.kp * q.hp0 * 8, f1 = (size_t)nb * g.kp * q.hp1 * 8, gg = (size_t)nb * 2 * g.kp * q.ny * 8;
    fp.FT0 = reinterpret_cast<cx<float>*>(w) + (size_t)lo * g.kp * q.hp0;
    fp.FT1 = reinterpret_cast<cx<float>*>(w + f0) + (size_t)lo * g.kp * q.hp1;
    fp.GT = reinterpret_cast<cx<float>*>(w + f0 + f1) + (size_t)lo * 2 * g.kp * q.ny;
    XcParams x = p;
    x.img0 = (const char*)p.img0 + (size_t)lo * q.h0 * q.w0 * q.isz;
    x.img1 = (const char*)p.img1 + (size_t)lo * q.h1 * q.w1 * q.isz;
    x.n = cnt;
    x.dx = p.dx + lo; x.dy = p.dy + lo; x.conf = p.conf + lo;
    x.peak = p.peak ? p.peak + lo : nullptr; x.mir = p.mir ? p.mir + lo : nullptr;
    x.part = reinterpret_cast<Partial*>(w + f0 + f1 + gg) + (size_t)lo * q.nrt;
    fp.rblk = fast_rblk(q.nx, TX);                       // rows per K3 tile
    if (q.nx == 1024 && (g_opt_fast_flags & 32)) fp.rblk = 4;
    fp.flags = (int)(g_opt_fast_flags & ~(16 | 32));
    fp.use_tma = make_gt_map(ctx, &fp.gt_map, fp.GT, cnt, q.ny, g.kp, fp.rblk) ? 1 : 0;
    fp.gt_tiles = q.ny / fp.rblk;
    fp.gt_pieces = fp.gt_tiles > 256 ? fp.gt_tiles / 256 : 1;
    x.G = fp.GT; x.gt_layout = fp.rblk; x.nrt = q.nrt; x.out_scale = x.scale;
    fp.hp0 = q.hp0; fp.hp1 = q.hp1;
    fp.x = x;
    return FB_OK;
}

// K4's scratch: a 4-line tile (the three rows around the peak at once) when it fits, else one line at a time
template <typename T> static bool finalize_narrow(int nx) { return (size_t)nx * 4 * sizeof(cx<T>) + 2048 > kMaxSmem; }
template <typename T> static size_t finalize_smem(int nx) { return (size_t)nx * (finalize_narrow<T>(nx) ? 1 : 4) * sizeof(cx<T>) + 2048; }

// stage 1: K1 rows forward, 2: K2 columns, 3: K3 rows inverse, 4: K4 finalize -- of the cnt pairs described by fp
static int fast_stage(int stage, const Problem& q, StreamCtx& ctx, const FastParams& fp, int cnt, int in_dtype, cudaStream_t st)
{
    int EX, TX, EY, TY;
    fast_et(q.nx, EX, TX); fast_et(q.ny, EY, TY);
    const Geometry& g = q.g;
    if (stage == 1) {
        const int TR = 2 * fast_lines(TX, kNW1);
        const int work = cnt * ((q.hp0 + TR - 1) / TR + (q.hp1 + TR - 1) / TR);
        const int cap = g_num_sms * (EX > 32 ? 1 : 16 / kNW1);
        const int grid = work < cap ? work : cap;
        const bool pruned = q.w0 <= q.nx / 2 && q.w1 <= q.nx / 2;
        ProfScope ps(ctx, st, SLOT_ROWS_FWD);
        FastLaunch l{q.nx, in_dtype, pruned, false, 0, false, grid, 32 * kNW1, fast_smem(q.nx, kNW1), st};
        if (!fast_dispatch(1, fp, l)) return fail(FB_ESIZE, "no fast-path row kernel for %d points", q.nx);
    } else if (stage == 2) {
        const int cpg = fast_lines(TY, kNW2) / 2;
        const int work = cnt * ((g.kp + cpg - 1) / cpg);
        const int cap = g_num_sms * (EY > 32 ? 1 : 16 / kNW2);
        const int grid = work < cap ? work : cap;
        const bool pruned = q.h0 <= q.ny / 2 && q.h1 <= q.ny / 2;     // (rows >= h of the row spectra are zero)
        ProfScope ps(ctx, st, SLOT_COLUMNS);
        FastLaunch l{q.ny, 0, pruned, false, 0, false, grid, 32 * kNW2, fast_smem(q.ny, kNW2), st};
        if (fast_k2_solo(q)) {
            const int cpg_s = fast_lines(TY, kNW2S);
            const int work_s = cnt * ((g.kp + cpg_s - 1) / cpg_s);
            l.k2_solo = true; l.grid = work_s < g_num_sms ? work_s : (int)g_num_sms; l.threads = 32 * kNW2S; l.smem = fast_smem_solo(q.ny);
        }
        if (!fast_dispatch(2, fp, l)) return fail(FB_ESIZE, "no fast-path column kernel for %d points", q.ny);
    } else if (stage == 3) {
        // K3: TX * R threads own R lines (R <= rblk rows of a GT tile)
        int R = fp.rblk;
        if (q.nx == 1024 && (g_opt_fast_flags & 16)) R = 4;          // experiment: half-tile CTAs
        const int nt = k3_threads(TX, R);
        const int work = cnt * (q.nrt / R);
        const int cap = g_num_sms * k3_ctas_per_sm(EX, nt);
        const int grid = work < cap ? work : cap;
        const int XS = TX * R + (R < 16 ? R : 0);
        const size_t sm3 = ((size_t)EX * XS + (kLaneTwiddles ? 0 : q.nx)) * sizeof(cx<float>) + (nt / 32) * R * 2 * (sizeof(float) + sizeof(double));
        ProfScope ps(ctx, st, SLOT_ROWS_INV);
        const bool tma3 = R == fp.rblk && !(g_opt_fast_flags & 4096);      // TMA-fed variant (default)
        const bool mir = q.conf_mode == CONF_MIRROR;
        const size_t sm3t = sm3 + 16;                                         // + mbarrier
        const int variant = (R == 4 && q.nx == 1024) ? (fp.rblk == 8 ? 3 : 2) : (tma3 ? 0 : 1);
        FastLaunch l{q.nx, 0, false, mir, variant, false, grid, nt, variant == 0 ? sm3t : sm3, st};
        if (!fast_dispatch(3, fp, l)) return fail(FB_ESIZE, "no fast-path inverse row kernel for %d points", q.nx);
    } else {
        ProfScope ps(ctx, st, SLOT_FINALIZE);
        fbk_finalize<float><<<cnt, 256, finalize_smem<float>(q.nx), st>>>(fp.x);
    }
    g_launches += 1;
    return FB_OK;
}

// work items and resident-CTA capacity of stage 1..3 for cnt pairs (the same numbers fast_stage launches with)
static void fast_work(int stage, const Problem& q, int cnt, long long& work, int& cap)
{
    int EX, TX, EY, TY;
    fast_et(q.nx, EX, TX); fast_et(q.ny, EY, TY);
    if (stage == 1) {
        const int TR = 2 * fast_lines(TX, kNW1);
        work = (long long)cnt * ((q.hp0 + TR - 1) / TR + (q.hp1 + TR - 1) / TR);
        cap = g_num_sms * (EX > 32 ? 1 : 16 / kNW1);
    } else if (stage == 2) {
        const bool solo = fast_k2_solo(q);
        const int cpg = solo ? fast_lines(TY, kNW2S) : fast_lines(TY, kNW2) / 2;
        work = (long long)cnt * ((q.g.kp + cpg - 1) / cpg);
        cap = solo ? (int)g_num_sms : g_num_sms * (EY > 32 ? 1 : 16 / kNW2);
    } else {
        const int R = fast_rblk(q.nx, TX);
        work = (long long)cnt * (q.nrt / R);
        cap = g_num_sms * k3_ctas_per_sm(EX, k3_threads(TX, R));
    }
}

template <typename TI>
static int launch_fast(const Problem& q, StreamCtx& ctx, XcParams& p, int nb, cudaStream_t st)
{
    int rc;
    if (!g_num_sms) { int sms = 0; CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device)); g_num_sms = sms; }
    const int in_dtype = std::is_same<TI, float>::value ? FB_F32 : FB_U8;
    // Two-stream schedule: the chunk is cut into S (2) independent parts, each running its four kernels on its own
    // stream (part 0 on the caller's).  The kernels are persistent and fill the GPU, so the streams mostly
    // alternate -- but the ramp-up and tail of every kernel (CTAs waiting for the slowest one, launch gaps) are
    // filled by the other part's kernel: +6 % on the 512^2 / FFT 1024^2 workload (profiles/two_streams.py).
    // Sharing the SMs between a column kernel and row kernels deliberately (one CTA each, dependency-pipelined
    // sub-chunks) was measured and lost 4-10 %: both kinds are issue / shared-memory bound.
    int S = (int)g_opt_pipeline;
    if (S > 2) S = 2;
    if (S == 2) {                                          // each half must still fill every kernel's grid (measured: +14 % at 16-24 pairs of 512^2)
        for (int stage = 1; stage <= 3 && S == 2; ++stage) {
            long long work; int cap;
            fast_work(stage, q, nb / 2, work, cap);
            if (work < g_opt_pipeline_waves * cap) S = 1;
        }
    }
    if (S <= 1) {
        FastParams fp{};
        if ((rc = fast_prepare(q, ctx, p, 0, nb, nb, fp)) != FB_OK) return rc;
        for (int stage = 1; stage <= 4; ++stage)
            if ((rc = fast_stage(stage, q, ctx, fp, nb, in_dtype, st)) != FB_OK) return rc;
        CU(cudaGetLastError());
        return FB_OK;
    }
    if (!ctx.side) CU(cudaStreamCreateWithFlags(&ctx.side, cudaStreamNonBlocking));
    while ((int)ctx.order_ev.size() < 2) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx.order_ev.push_back(e);
    }
    cudaEvent_t fork = ctx.order_ev[0], join = ctx.order_ev[1];
    FastParams fps[2] = {};
    const int mid = nb / 2;
    if ((rc = fast_prepare(q, ctx, p, 0, mid, nb, fps[0])) != FB_OK) return rc;
    if ((rc = fast_prepare(q, ctx, p, mid, nb - mid, nb, fps[1])) != FB_OK) return rc;
    cudaStream_t sd = ctx.side;
    CU(cudaEventRecord(fork, st));                         // inputs (and the workspace's previous use) are ordered on st
    CU(cudaStreamWaitEvent(sd, fork, 0));
    for (int stage = 1; stage <= 4; ++stage) {             // interleaved enqueue: both queues fill at the same pace
        if ((rc = fast_stage(stage, q, ctx, fps[0], mid, in_dtype, st)) != FB_OK) return rc;
        if ((rc = fast_stage(stage, q, ctx, fps[1], nb - mid, in_dtype, sd)) != FB_OK) return rc;
    }
    CU(cudaEventRecord(join, sd));
    CU(cudaStreamWaitEvent(st, join, 0));                  // results are ordered on the caller's stream again
    CU(cudaGetLastError());
    return FB_OK;
}

// ---------------------------------------------------------------------------
// launch of one chunk (device pointers)
// ---------------------------------------------------------------------------
template <typename T, typename TI>
static int launch_chunk(const Problem& q, StreamCtx& ctx, const void* img0, const void* img1, int nb,
                        double* dx, double* dy, double* conf, double* peak, double* mir, cudaStream_t st)
{
    XcParams p{};
    int rc;
    if ((rc = get_tables<T>(ctx.device, q.nx, p.px, q.fused)) != FB_OK) return rc;     // fused kernel: wide pass radices
    if ((rc = get_tables<T>(ctx.device, q.ny, p.py, q.fused)) != FB_OK) return rc;
    const Geometry& g = q.g;
    p.img0 = img0; p.img1 = img1; p.n = nb;
    p.h0 = q.h0; p.w0 = q.w0; p.h1 = q.h1; p.w1 = q.w1; p.ny = q.ny; p.nx = q.nx; p.kp = g.kp;
    p.fpitch = g.fpitch; p.dx = dx; p.dy = dy; p.conf = conf; p.peak = peak; p.mir = mir;
    p.conf_mode = q.conf_mode; p.subpixel = q.subpixel; p.scale = 1.0 / ((double)q.ny * (double)q.nx);
    p.out_scale = 1.0;
    p.fin_narrow = (!q.fused && finalize_narrow<T>(q.nx)) ? 1 : 0;
    if (q.wf) {
        if constexpr (std::is_same<T, float>::value) {
            WfParams wp{};
            wp.x = p;
            int ty, tx;
            wf_lanes(q.ny, q.nx, ty, tx);
            if ((rc = get_wf_table(ctx.device, q.nx, tx, wp.twx)) != FB_OK) return rc;
            if ((rc = get_wf_table(ctx.device, q.ny, ty, wp.twy)) != FB_OK) return rc;
            if (!g_num_sms) { int sms = 0; CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx.device)); g_num_sms = sms; }
            WfLaunch l{q.ny, q.nx, std::is_same<TI, float>::value ? FB_F32 : FB_U8, 0, st, 0, 0};
            wf_shape(q.ny, q.nx, l.threads, l.smem);
            const int cap = g_num_sms * (l.smem <= 112 * 1024 ? 2 : 1);
            l.grid = nb < cap ? nb : cap;
            { ProfScope ps(ctx, st, SLOT_FUSED); wf_launch_a(wp, l) || wf_launch_b(wp, l) || wf_launch_c(wp, l); }
            g_launches += 1;
            CU(cudaGetLastError());
            return FB_OK;
        }
    }
    if (q.fused) {
        p.tl = g.tl_fused; p.spitch = g.spitch;
        int nthr = g.smem_fused > 100 * 1024 ? 512 : 256;
        if (g_opt_fused_threads) nthr = (int)g_opt_fused_threads;
        { ProfScope ps(ctx, st, SLOT_FUSED); fbk_fused<T, TI><<<nb, nthr, g.smem_fused, st>>>(p); }
        g_launches += 1;
        CU(cudaGetLastError());
        return FB_OK;
    }
    p.tl = g.tl_row; p.tc = g.tc_col; p.nrt = q.nrt;
    if (q.fast) {
        if constexpr (std::is_same<T, float>::value) return launch_fast<TI>(q, ctx, p, nb, st);
    }
    unsigned char* w = reinterpret_cast<unsigned char*>(ctx.ws);
    const int C = q.nchan;
    size_t f0 = (size_t)nb * C * q.h0 * g.fpitch * g.esize, f1 = (size_t)nb * C * q.h1 * g.fpitch * g.esize;
    const size_t gpair = (size_t)q.ny * 2 * g.fpitch;                       // complex elements of one pair's P | Q block
    size_t gg = (size_t)nb * C * gpair * g.esize, gm = C > 1 ? (size_t)nb * gpair * g.esize : 0;
    p.F0 = w; p.F1 = w + f0; p.G = w + f0 + f1; p.part = reinterpret_cast<Partial*>(w + f0 + f1 + gg + gm);
    p.norm = q.ext.norm; p.norm_m = q.ext.norm_mirror; p.surf = q.ext.surface; p.surf_m = q.ext.surface_mirror;
    const int t0 = row_tiles<T>(q.h0, p.tl), t1 = row_tiles<T>(q.h1, p.tl);
    const int nct = (g.kp + p.tc - 1) / p.tc;
    p.n = nb * C;                                                           // channels are extra pairs up to the column stage
    { ProfScope ps(ctx, st, SLOT_ROWS_FWD); fbk_rows_forward<T, TI><<<nb * C * (t0 + t1), g.nthreads_row, g.smem_row, st>>>(p); }
    { ProfScope ps(ctx, st, SLOT_COLUMNS); fbk_columns<T><<<nb * C * nct, g.nthreads_col, g.smem_col, st>>>(p); }
    if (C > 1) {
        cx<T>* gout = reinterpret_cast<cx<T>*>(w + f0 + f1 + gg);
        const size_t total = (size_t)nb * gpair;
        const int grid = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
        fbk_channel_mean<T><<<grid, 256, 0, st>>>(reinterpret_cast<const cx<T>*>(p.G), gout, C, gpair, total, g.fpitch, g.kp);
        p.G = gout;
        g_launches += 1;
    }
    p.n = nb;
    { ProfScope ps(ctx, st, SLOT_ROWS_INV); fbk_rows_inverse<T><<<nb * q.nrt, g.nthreads_row, g.smem_row, st>>>(p); }
    { ProfScope ps(ctx, st, SLOT_FINALIZE); fbk_finalize<T><<<nb, 256, finalize_smem<T>(q.nx), st>>>(p); }
    g_launches += 4;
    CU(cudaGetLastError());
    return FB_OK;
}

static int launch_chunk_any(const Problem& q, StreamCtx& ctx, const void* img0, const void* img1, int nb,
                            double* dx, double* dy, double* conf, double* peak, double* mir, cudaStream_t st)
{
    g_pairs += nb;
    if (q.in_dtype == FB_F32) return launch_chunk<float, float>(q, ctx, img0, img1, nb, dx, dy, conf, peak, mir, st);
    if (q.in_dtype == FB_U8) {
        if (q.f64) return launch_chunk<double, unsigned char>(q, ctx, img0, img1, nb, dx, dy, conf, peak, mir, st);
        return launch_chunk<float, unsigned char>(q, ctx, img0, img1, nb, dx, dy, conf, peak, mir, st);
    }
    return launch_chunk<double, double>(q, ctx, img0, img1, nb, dx, dy, conf, peak, mir, st);
}

// all pointers device; loops over workspace-sized chunks
static int run_device(const Problem& q, StreamCtx& ctx, const void* img0, const void* img1, int n,
                      double* dx, double* dy, double* conf, double* peak, double* mir, cudaStream_t st)
{
    int chunk = n;
    if (!q.fused) {
        long long fit = g_opt_ws_bytes / (long long)q.ws_per_pair;
        if (fit < 1) fit = 1;
        if (chunk > fit) chunk = (int)fit;
        size_t need = (size_t)chunk * q.ws_per_pair;
        if (need > ctx.ws_bytes) {
            if (ctx.ws) { CU(cudaStreamSynchronize(st)); CU(cudaFree(ctx.ws)); ctx.ws = nullptr; ctx.ws_bytes = 0; ctx.gt_maps.clear(); }
            // the budget is a ceiling, not a reservation: when the device cannot give that much (worker processes
            // sharing one GPU), work in smaller chunks instead of failing
            while (cudaMalloc(&ctx.ws, need) != cudaSuccess) {
                cudaGetLastError();
                ctx.ws = nullptr;
                if (chunk == 1) return fail(FB_ENOMEM, "workspace of %zu bytes", need);
                chunk = (chunk + 1) / 2;
                need = (size_t)chunk * q.ws_per_pair;
            }
            ctx.ws_bytes = need;
        }
    } else if (chunk > 65535 * 16) {
        chunk = 65535 * 16;
    }
    const size_t b0 = (size_t)q.h0 * q.w0 * q.isz * q.nchan, b1 = (size_t)q.h1 * q.w1 * q.isz * q.nchan;
    for (int lo = 0; lo < n; lo += chunk) {
        int nb = n - lo < chunk ? n - lo : chunk;
        int rc = launch_chunk_any(q, ctx, (const char*)img0 + lo * b0, (const char*)img1 + lo * b1, nb,
                                  dx + lo, dy + lo, conf + lo, peak ? peak + lo : nullptr, mir ? mir + lo : nullptr, st);
        if (rc != FB_OK) return rc;
    }
    return FB_OK;
}

static int get_ctx(int device, void* stream, StreamCtx*& out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(FB_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(FB_EINVAL, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = set_attrs(device);
    if (rc != FB_OK) return rc;
    StreamCtx& c = g_ctx[std::make_pair(device, stream)];      // map nodes are stable: the pointer outlives the lock
    c.device = device;
    out = &c;
    return FB_OK;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int fb_xcorr_batch_device(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                                     int in_dtype, int fft_h, int fft_w, int flags,
                                     double* dx, double* dy, double* conf, double* peak, double* mirror,
                                     int device, void* stream)
{
    Problem q;
    int rc = make_problem(q, n, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags);
    if (rc != FB_OK) return rc;
    if (n == 0) return FB_OK;
    if (!img0 || !img1 || !dx || !dy || !conf) return fail(FB_EINVAL, "null pointer");
    StreamCtx* ctx;
    if ((rc = get_ctx(device, stream, ctx)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_device(q, *ctx, img0, img1, n, dx, dy, conf, peak, mirror, (cudaStream_t)stream);
}

extern "C" int fb_xcorr_batch_device_ex(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                                        int in_dtype, int fft_h, int fft_w, int flags,
                                        double* dx, double* dy, double* conf, double* peak, double* mirror,
                                        int device, void* stream, const fb_xcorr_ext* ext)
{
    Problem q;
    int rc = make_problem(q, n, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags, ext);
    if (rc != FB_OK) return rc;
    if (n == 0) return FB_OK;
    if (!img0 || !img1 || !dx || !dy || !conf) return fail(FB_EINVAL, "null pointer");
    StreamCtx* ctx;
    if ((rc = get_ctx(device, stream, ctx)) != FB_OK) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    return run_device(q, *ctx, img0, img1, n, dx, dy, conf, peak, mirror, (cudaStream_t)stream);
}

static std::atomic<long long> g_opt_copy_threads{6};

// dst0 <- src0 (n0 bytes) and dst1 <- src1 (n1 bytes), cut into pieces for up to `copy_threads` host threads
static void parallel_copy(void* dst0, const void* src0, size_t n0, void* dst1, const void* src1, size_t n1)
{
    const int want = (int)g_opt_copy_threads.load();
    const size_t total = n0 + n1, piece_min = (size_t)4 << 20;
    int nt = (int)(total / piece_min);
    nt = nt < 1 ? 1 : (nt > want ? want : nt);
    if (nt <= 1) { memcpy(dst0, src0, n0); memcpy(dst1, src1, n1); return; }
    std::vector<std::thread> team;
    auto part = [&](int k) {
        // thread k takes the k-th slice of both arrays
        const size_t a0 = n0 * k / nt, e0 = n0 * (k + 1) / nt, a1 = n1 * k / nt, e1 = n1 * (k + 1) / nt;
        memcpy((char*)dst0 + a0, (const char*)src0 + a0, e0 - a0);
        memcpy((char*)dst1 + a1, (const char*)src1 + a1, e1 - a1);
    };
    for (int k = 1; k < nt; ++k) team.emplace_back(part, k);
    part(0);
    for (auto& th : team) th.join();
}

extern "C" int fb_xcorr_batch_host(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                                   int in_dtype, int fft_h, int fft_w, int flags,
                                   double* dx, double* dy, double* conf, double* peak, double* mirror,
                                   int device, void* stream)
{
    Problem q;
    int rc = make_problem(q, n, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags);
    if (rc != FB_OK) return rc;
    if (n == 0) return FB_OK;
    if (!img0 || !img1 || !dx || !dy || !conf) return fail(FB_EINVAL, "null pointer");
    // a NULL stream means "the library's own stream" here, not the legacy default stream the device entry points use
    // for NULL: the context (workspace) is keyed by the stream the kernels really run on
    StreamCtx* cp;
    if ((rc = get_ctx(device, stream ? stream : kOwnStreamKey, cp)) != FB_OK) return rc;
    StreamCtx& c = *cp;
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.copy_stream) {
        CU(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&c.ev_copied[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c.ev_done[i], cudaEventDisableTiming));
        }
    }
    cudaStream_t st = stream ? (cudaStream_t)stream : c.own_stream;
    const size_t b0 = (size_t)h0 * w0 * q.isz, b1 = (size_t)h1 * w1 * q.isz;
    long long hc = g_opt_host_chunk / (long long)(b0 + b1);
    if (hc < 1) hc = 1;
    if (hc > n) hc = n;
    // pinned or pageable?
    cudaPointerAttributes a0{}, a1{};
    bool pinned = cudaPointerGetAttributes(&a0, img0) == cudaSuccess && cudaPointerGetAttributes(&a1, img1) == cudaSuccess &&
                  a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
    cudaGetLastError();
    const size_t slot_bytes = (size_t)hc * (b0 + b1) + 512;
    if (slot_bytes > c.din_bytes) {
        CU(cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            if (c.din[i]) CU(cudaFree(c.din[i]));
            c.din[i] = nullptr;
            if (cudaMalloc(&c.din[i], slot_bytes) != cudaSuccess) { cudaGetLastError(); c.din_bytes = 0; return fail(FB_ENOMEM, "input slot of %zu bytes", slot_bytes); }
        }
        c.din_bytes = slot_bytes;
    }
    if (!pinned && slot_bytes > c.pin_bytes) {
        CU(cudaDeviceSynchronize());
        for (int i = 0; i < 2; ++i) {
            if (c.pin[i]) CU(cudaFreeHost(c.pin[i]));
            c.pin[i] = nullptr;
            if (cudaMallocHost(&c.pin[i], slot_bytes) != cudaSuccess) { cudaGetLastError(); c.pin_bytes = 0; return fail(FB_ENOMEM, "pinned slot of %zu bytes", slot_bytes); }
        }
        c.pin_bytes = slot_bytes;
    }
    const size_t ob = (size_t)n * 5 * sizeof(double);
    if (ob > c.dout_bytes) {
        CU(cudaDeviceSynchronize());
        if (c.dout) CU(cudaFree(c.dout));
        c.dout = nullptr;
        if (cudaMalloc((void**)&c.dout, ob) != cudaSuccess) { cudaGetLastError(); c.dout_bytes = 0; return fail(FB_ENOMEM, "output buffer"); }
        c.dout_bytes = ob;
    }
    double* o = c.dout;
    const size_t off1 = ((size_t)hc * b0 + 255) & ~(size_t)255;     // img1 offset inside a slot
    int ci = 0;
    for (int lo = 0; lo < n; lo += (int)hc, ++ci) {
        const int nb = n - lo < hc ? n - lo : (int)hc;
        const int s = ci & 1;
        char* d = (char*)c.din[s];
        const char* src0 = (const char*)img0 + (size_t)lo * b0;
        const char* src1 = (const char*)img1 + (size_t)lo * b1;
        if (ci >= 2) {
            if (pinned) CU(cudaStreamWaitEvent(c.copy_stream, c.ev_done[s], 0));
            else CU(cudaEventSynchronize(c.ev_done[s]));      // also implies the slot's H2D finished
        }
        if (!pinned) {
            // pageable arrays (what FEABAS callers pass): staged through the pinned slot by several host threads -- one
            // thread copies at ~10 GB/s, a fifth of what the link takes
            parallel_copy(c.pin[s], src0, (size_t)nb * b0, (char*)c.pin[s] + off1, src1, (size_t)nb * b1);
            src0 = (const char*)c.pin[s];
            src1 = (const char*)c.pin[s] + off1;
        }
        CU(cudaMemcpyAsync(d, src0, (size_t)nb * b0, cudaMemcpyHostToDevice, c.copy_stream));
        CU(cudaMemcpyAsync(d + off1, src1, (size_t)nb * b1, cudaMemcpyHostToDevice, c.copy_stream));
        CU(cudaEventRecord(c.ev_copied[s], c.copy_stream));
        CU(cudaStreamWaitEvent(st, c.ev_copied[s], 0));
        rc = run_device(q, c, d, d + off1, nb, o + lo, o + n + lo, o + 2 * (size_t)n + lo, o + 3 * (size_t)n + lo, o + 4 * (size_t)n + lo, st);
        if (rc != FB_OK) return rc;
        CU(cudaEventRecord(c.ev_done[s], st));
    }
    CU(cudaMemcpyAsync(dx, o, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(dy, o + n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(conf, o + 2 * (size_t)n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (peak) CU(cudaMemcpyAsync(peak, o + 3 * (size_t)n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (mirror) CU(cudaMemcpyAsync(mirror, o + 4 * (size_t)n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return FB_OK;
}

extern "C" int fb_xcorr_batch(const void* img0, const void* img1, int n, int h0, int w0, int h1, int w1,
                              int in_dtype, int fft_h, int fft_w, int flags,
                              double* dx, double* dy, double* conf, double* peak, double* mirror,
                              int device, void* stream)
{
    cudaPointerAttributes a{};
    bool on_device = img0 && cudaPointerGetAttributes(&a, img0) == cudaSuccess &&
                     (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    if (on_device)
        return fb_xcorr_batch_device(img0, img1, n, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags, dx, dy, conf, peak, mirror, device, stream);
    return fb_xcorr_batch_host(img0, img1, n, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags, dx, dy, conf, peak, mirror, device, stream);
}

extern "C" int fb_next_fast_len(int target)
{
    if (target <= 1) return 1;
    for (int n = target;; ++n)
        if (is_5smooth(n)) return n;
}

extern "C" int fb_xcorr_plan_info(int h0, int w0, int h1, int w1, int in_dtype, int fft_h, int fft_w, int flags,
                                  long long* info)
{
    Problem q;
    int rc = make_problem(q, 1, h0, w0, h1, w1, in_dtype, fft_h, fft_w, flags);
    if (rc != FB_OK) return rc;
    if (!info) return fail(FB_EINVAL, "null info");
    info[0] = q.wf ? 4 : (q.fused ? 1 : (q.fast ? 3 : 2));
    info[1] = (long long)q.ws_per_pair;
    info[2] = q.g.fused ? (long long)q.g.smem_fused : 0;
    info[3] = (long long)q.g.smem_row;
    info[4] = (long long)q.g.smem_col;
    info[5] = q.fused ? q.g.tl_fused : q.g.tl_row;
    info[6] = q.g.tc_col;
    info[7] = q.fused ? 1 : 4;
    return FB_OK;
}

extern "C" int fb_set_option(const char* name, long long value)
{
    if (!name) return fail(FB_EINVAL, "null option name");
    std::lock_guard<std::mutex> lk(g_mu);
    if (!strcmp(name, "fused_threads")) { if (value && (value < 64 || value > 1024 || value % 32)) return fail(FB_EINVAL, "fused_threads out of range"); g_opt_fused_threads = value; return FB_OK; }
    if (!strcmp(name, "pipeline_waves")) { if (value < 1 || value > 64) return fail(FB_EINVAL, "pipeline_waves out of range"); g_opt_pipeline_waves = value; return FB_OK; }
    if (!strcmp(name, "pipeline")) { if (value < 1 || value > 16) return fail(FB_EINVAL, "pipeline out of range"); g_opt_pipeline = value; return FB_OK; }
    if (!strcmp(name, "max_radix")) { if (value < 5 || value > 16) return fail(FB_EINVAL, "max_radix out of range"); g_opt_max_radix = value; return FB_OK; }
    if (!strcmp(name, "ws_bytes")) { if (value < (1 << 20)) return fail(FB_EINVAL, "ws_bytes too small"); g_opt_ws_bytes = value; return FB_OK; }
    if (!strcmp(name, "profile")) { g_opt_profile = value ? 1 : 0; return FB_OK; }
    if (!strcmp(name, "fast_flags")) { g_opt_fast_flags = value; return FB_OK; }
    if (!strcmp(name, "k2_solo")) { g_opt_k2_solo = value ? 1 : 0; return FB_OK; }
    if (!strcmp(name, "warp_fused")) { g_opt_warp_fused = value ? 1 : 0; return FB_OK; }
    if (!strcmp(name, "copy_threads")) { if (value < 1 || value > 64) return fail(FB_EINVAL, "copy_threads out of range"); g_opt_copy_threads = value; return FB_OK; }
    if (!strcmp(name, "host_chunk_bytes")) { if (value < 4096) return fail(FB_EINVAL, "host_chunk_bytes too small"); g_opt_host_chunk = value; return FB_OK; }
    return fail(FB_EINVAL, "unknown option %s", name);
}

extern "C" int fb_profile_read(int device, void* stream, double* ms5, long long* launches5, int reset)
{
    std::vector<StreamCtx*> todo;
    const bool all = stream == (void*)(intptr_t)-1;        // FB_ALL_STREAMS: every context of the device, summed
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (all) {
            for (auto& kv : g_ctx) if (kv.first.first == device) todo.push_back(&kv.second);
        } else {
            auto it = g_ctx.find(std::make_pair(device, stream));
            if (it == g_ctx.end()) it = g_ctx.find(std::make_pair(device, stream ? stream : kOwnStreamKey));
            if (it == g_ctx.end()) return fail(FB_EINVAL, "no context for device %d / stream %p", device, stream);
            todo.push_back(&it->second);
        }
    }
    for (int i = 0; i < 5; ++i) { if (ms5) ms5[i] = 0; if (launches5) launches5[i] = 0; }
    CU(cudaSetDevice(device));
    for (StreamCtx* cp : todo) {
        StreamCtx& c = *cp;
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto& t : c.prof_pending) {
            float ms = 0.f;
            CU(cudaEventSynchronize(std::get<2>(t)));
            CU(cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t)));
            c.prof_ms[std::get<0>(t)] += ms;
            c.prof_n[std::get<0>(t)] += 1;
            c.prof_free.push_back(std::get<1>(t));
            c.prof_free.push_back(std::get<2>(t));
        }
        c.prof_pending.clear();
        for (int i = 0; i < 5; ++i) {
            if (ms5) ms5[i] += c.prof_ms[i];
            if (launches5) launches5[i] += c.prof_n[i];
            if (reset) { c.prof_ms[i] = 0; c.prof_n[i] = 0; }
        }
    }
    return FB_OK;
}

extern "C" long long fb_launch_count(void) { return g_launches.load(); }
extern "C" long long fb_pair_count(void) { return g_pairs.load(); }

extern "C" int fb_release(int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_ctx.begin(); it != g_ctx.end();) {
        StreamCtx& c = it->second;
        if (device >= 0 && c.device != device) { ++it; continue; }
        c.mu.lock();                           // waits for a call in flight on this context (callers must not START one now)
        cudaSetDevice(c.device);
        cudaDeviceSynchronize();
        if (c.ws) cudaFree(c.ws);
        for (int i = 0; i < 2; ++i) {
            if (c.din[i]) cudaFree(c.din[i]);
            if (c.pin[i]) cudaFreeHost(c.pin[i]);
            if (c.ev_copied[i]) cudaEventDestroy(c.ev_copied[i]);
            if (c.ev_done[i]) cudaEventDestroy(c.ev_done[i]);
        }
        if (c.dout) cudaFree(c.dout);
        for (auto& t : c.prof_pending) { cudaEventDestroy(std::get<1>(t)); cudaEventDestroy(std::get<2>(t)); }
        for (auto e : c.prof_free) cudaEventDestroy(e);
        if (c.side) cudaStreamDestroy(c.side);
        for (auto e : c.order_ev) cudaEventDestroy(e);
        if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
        if (c.own_stream) cudaStreamDestroy(c.own_stream);
        c.mu.unlock();
        it = g_ctx.erase(it);
    }
    for (auto it = g_tables.begin(); it != g_tables.end();) {
        if (device >= 0 && std::get<0>(it->first) != device) { ++it; continue; }
        cudaSetDevice(std::get<0>(it->first));
        cudaFree(it->second.tw);
        cudaFree(it->second.pos);
        it = g_tables.erase(it);
    }
    for (auto it = g_wtables.begin(); it != g_wtables.end();) {
        if (device >= 0 && std::get<0>(it->first) != device) { ++it; continue; }
        cudaSetDevice(std::get<0>(it->first));
        cudaFree(it->second);
        it = g_wtables.erase(it);
    }
    cudaGetLastError();
    return FB_OK;
}

extern "C" int fb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* fb_last_error(void) { return g_err.c_str(); }
extern "C" const char* fb_version(void) { return "feabas_cuda 0.1 (sm_100a)"; }
