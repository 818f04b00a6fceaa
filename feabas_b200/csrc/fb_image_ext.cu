// fb_image_ext.cu -- the less travelled branches of the image operators (device pointers):
//
//   fb_resize_area_frac   cv2.resize(INTER_AREA) for ANY shrinking factor   feabas/matcher.py:254-256,321-322
//   fb_masked_dog_f64     common.masked_dog_filter on float64 images        feabas/common.py:353-377 (:363 keeps float64)
//
// Both follow the arithmetic of the third-party code the reference calls (OpenCV's ResizeArea_ with its
// DecimateAlpha tables; scipy.ndimage.correlate1d's symmetric branch) operation by operation, without
// contraction, so that the results are the reference's bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/feabas_cuda.h"
#include "fb_common.h"

namespace {

int check_device_ext(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fb_failf(FB_ECUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fb_failf(FB_EINVAL, "device %d out of range (%d devices)", device, ndev);
    FB_CU(cudaSetDevice(device));
    return FB_OK;
}

// ---------------------------------------------------------------------------------------------
// INTER_AREA, general shrinking factor.  OpenCV (imgproc/resize.cpp) builds one table per axis
// (computeResizeAreaTab): output cell d covers source [d * scale, (d + 1) * scale); a partially
// covered first pixel, the fully covered ones, a partially covered last pixel, each with a float32
// weight (covered length / cell length, the last cell clipped to the image).  ResizeArea_ then
// accumulates in float32, first along x into a row buffer (buf += S * alpha, from 0), then along y
// (sum = beta * buf for the first row of a cell, sum += beta * buf after); uint8 output is
// saturate_cast<uchar>(sum) = round half to even.
// ---------------------------------------------------------------------------------------------
struct AreaEntry { int si; float alpha; };

void area_table(int ssize, int dsize, double scale, std::vector<int>& ofs, std::vector<AreaEntry>& tab)
{
    ofs.assign(dsize + 1, 0);
    tab.clear();
    for (int d = 0; d < dsize; ++d) {
        ofs[d] = (int)tab.size();
        const double f1 = d * scale, f2 = f1 + scale;
        const double cell = fmin(scale, ssize - f1);
        int s1 = (int)ceil(f1), s2 = (int)floor(f2);
        s2 = s2 < ssize - 1 ? s2 : ssize - 1;
        s1 = s1 < s2 ? s1 : s2;
        if (s1 - f1 > 1e-3) tab.push_back(AreaEntry{s1 - 1, (float)((s1 - f1) / cell)});
        for (int s = s1; s < s2; ++s) tab.push_back(AreaEntry{s, (float)(1.0 / cell)});
        if (f2 - s2 > 1e-3) tab.push_back(AreaEntry{s2, (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell)});
    }
    ofs[dsize] = (int)tab.size();
}

template <typename TS>
__global__ void __launch_bounds__(256) fbk_resize_area_frac(const TS* src, int n, int h, int w, int oh, int ow,
                                                            const int* xofs, const AreaEntry* xtab,
                                                            const int* yofs, const AreaEntry* ytab, TS* dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= ow) return;
    const TS* s = src + (size_t)img * h * w;
    const int x0 = xofs[x], x1 = xofs[x + 1], y0 = yofs[y], y1 = yofs[y + 1];
    float sum = 0.f;
    for (int j = y0; j < y1; ++j) {
        const AreaEntry ye = ytab[j];
        const TS* row = s + (size_t)ye.si * w;
        float buf = 0.f;
        for (int k = x0; k < x1; ++k) {
            const AreaEntry xe = xtab[k];
            buf = __fadd_rn(buf, __fmul_rn((float)__ldg(row + xe.si), xe.alpha));
        }
        const float t = __fmul_rn(ye.alpha, buf);
        sum = j == y0 ? t : __fadd_rn(sum, t);
    }
    TS out;
    if (sizeof(TS) == 1) {
        const int q = __float2int_rn(sum);
        out = (TS)(q < 0 ? 0 : (q > 255 ? 255 : q));
    } else {
        out = (TS)sum;
    }
    dst[((size_t)img * oh + y) * ow + x] = out;
}

// ---------------------------------------------------------------------------------------------
// float64 band-pass.  scipy.ndimage.gaussian_filter1d on a float64 array: correlate1d, symmetric
// branch (ni_filters.c): tmp = x[0] * w[0], then tmp += (x[-i] + x[i]) * w[i] from the outermost tap inwards -- in
// double, one rounding per operation, mode='nearest'.  One kernel per 1-D pass, straight from global memory: this
// branch has no in-tree caller (FEABAS images are uint8), it is here for completeness, not speed.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRadius64 = 120;

struct Taps64 {
    int radius;
    double w[kMaxRadius64 + 1];      // w[j] = weight at offset +-(radius - j); w[radius] = centre
};

bool make_taps64(double sigma, Taps64& t)
{
    const int r = (int)(4.0 * sigma + 0.5);
    if (r < 0 || r > kMaxRadius64 || !(sigma > 0)) return false;
    std::vector<double> phi(2 * r + 1);
    const double s2 = sigma * sigma;
    double sum = 0;
    for (int x = -r; x <= r; ++x) phi[x + r] = exp(-0.5 / s2 * (double)(x * x));
    for (double v : phi) sum += v;
    t.radius = r;
    for (int j = 0; j <= r; ++j) t.w[j] = phi[j] / sum;
    return true;
}

// AXIS 0: along x (axis=-1), 1: along y (axis=-2).  MASKSRC: the source is ptp * (mask == 0) built from mask bytes.
template <int AXIS, bool MASKSRC>
__global__ void __launch_bounds__(256) fbk_gauss1d_f64(const double* src, const unsigned char* mask, long long mask_stride,
                                                       const double* span, double span_value,
                                                       int n, int h, int w, double* dst, const __grid_constant__ Taps64 t)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= w) return;
    const int r = t.radius;
    const size_t base = (size_t)img * h * w;
    const double sp = MASKSRC ? (span ? span[1] - span[0] : span_value) : 0.0;
    auto at = [&](int yy, int xx) -> double {
        yy = min(max(yy, 0), h - 1);
        xx = min(max(xx, 0), w - 1);
        if (MASKSRC) return mask[(size_t)img * mask_stride + (size_t)yy * w + xx] ? 0.0 : sp;
        return src[base + (size_t)yy * w + xx];
    };
    double acc = __dmul_rn(at(y, x), t.w[r]);
    for (int j = 0; j < r; ++j) {
        const int o = r - j;
        const double a = AXIS == 0 ? at(y, x - o) : at(y - o, x), b = AXIS == 0 ? at(y, x + o) : at(y + o, x);
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(a, b), t.w[j]));
    }
    dst[base + (size_t)y * w + x] = acc;
}

// out = a - b (optionally |.|)
__global__ void __launch_bounds__(256) fbk_sub_f64(const double* a, const double* b, long long count, int take_abs, double* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double f = __dsub_rn(a[i], b[i]);
    out[i] = take_abs ? fabs(f) : f;
}

// common.py:371-374: maskf = G(G(mask_img)) * sigma_c^2 / sigma_0^2; imgf = clip(|imgf| - maskf, 0) * sign(imgf)
__global__ void __launch_bounds__(256) fbk_mask_apply_f64(const double* gm, long long count, double sc2, double s02, int take_abs, double* out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const double mf = __ddiv_rn(__dmul_rn(gm[i], sc2), s02);
    const double f = out[i];
    double a = fmax(__dsub_rn(fabs(f), mf), 0.0);
    if (!take_abs) a = f > 0.0 ? a : (f < 0.0 ? -a : __dmul_rn(a, 0.0));
    out[i] = a;
}

__global__ void __launch_bounds__(512) fbk_minmax_f64(const double* src, long long count, double* partial)
{
    __shared__ double slo[16], shi[16];
    double lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double v = src[i];
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int off = 16; off; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 16; ++i) { lo = fmin(lo, slo[i]); hi = fmax(hi, shi[i]); }
        partial[2 * blockIdx.x] = lo; partial[2 * blockIdx.x + 1] = hi;
    }
}
__global__ void fbk_minmax_fold_f64(const double* partial, int n, double* out)
{
    double lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n; i += 32) { lo = fmin(lo, partial[2 * i]); hi = fmax(hi, partial[2 * i + 1]); }
    for (int off = 16; off; off >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (threadIdx.x == 0) { out[0] = lo; out[1] = hi; }
}

constexpr int kMinmaxBlocks = 296;

}  // namespace

extern "C" int fb_resize_area_frac(const void* src, int n, int h, int w, int in_dtype, double inv_fx, double inv_fy,
                                   void* dst, int oh, int ow, int device, void* stream)
{
    if (n < 0 || h < 1 || w < 1 || oh < 1 || ow < 1) return fb_failf(FB_EINVAL, "bad shape");
    if (in_dtype != FB_F32 && in_dtype != FB_U8) return fb_failf(FB_EINVAL, "resize_area: dtype %d not supported", in_dtype);
    if (!(inv_fx >= 1.0) || !(inv_fy >= 1.0)) return fb_failf(FB_EINVAL, "resize_area_frac shrinks: 1/fx = %g, 1/fy = %g must be >= 1", inv_fx, inv_fy);
    if ((oh - 1) * inv_fy >= h || (ow - 1) * inv_fx >= w) return fb_failf(FB_EINVAL, "output %dx%d too large for %dx%d", oh, ow, h, w);
    if (n == 0) return FB_OK;
    if (!src || !dst) return fb_failf(FB_EINVAL, "null pointer");
    int rc = check_device_ext(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<int> xofs, yofs;
    std::vector<AreaEntry> xtab, ytab;
    area_table(w, ow, inv_fx, xofs, xtab);
    area_table(h, oh, inv_fy, yofs, ytab);
    // one device buffer: xofs | yofs | xtab | ytab (8-byte entries after the 4-byte offsets, padded)
    const size_t b_xo = xofs.size() * sizeof(int), b_yo = yofs.size() * sizeof(int);
    const size_t o_xt = (b_xo + b_yo + 7) & ~(size_t)7, b_xt = xtab.size() * sizeof(AreaEntry), b_yt = ytab.size() * sizeof(AreaEntry);
    std::vector<unsigned char> host(o_xt + b_xt + b_yt);
    memcpy(host.data(), xofs.data(), b_xo);
    memcpy(host.data() + b_xo, yofs.data(), b_yo);
    memcpy(host.data() + o_xt, xtab.data(), b_xt);
    memcpy(host.data() + o_xt + b_xt, ytab.data(), b_yt);
    unsigned char* dev = nullptr;
    FB_CU(cudaMallocAsync(&dev, host.size(), st));
    // pageable source: the runtime stages it before returning, `host` may go out of scope
    cudaError_t e = cudaMemcpyAsync(dev, host.data(), host.size(), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFreeAsync(dev, st); return fb_failf(FB_ECUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e)); }
    const int* d_xo = reinterpret_cast<const int*>(dev);
    const int* d_yo = reinterpret_cast<const int*>(dev + b_xo);
    const AreaEntry* d_xt = reinterpret_cast<const AreaEntry*>(dev + o_xt);
    const AreaEntry* d_yt = reinterpret_cast<const AreaEntry*>(dev + o_xt + b_xt);
    dim3 grid((ow + 255) / 256, oh, n);
    if (in_dtype == FB_F32)
        fbk_resize_area_frac<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), n, h, w, oh, ow, d_xo, d_xt, d_yo, d_yt, reinterpret_cast<float*>(dst));
    else
        fbk_resize_area_frac<unsigned char><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(src), n, h, w, oh, ow, d_xo, d_xt, d_yo, d_yt, reinterpret_cast<unsigned char*>(dst));
    fb_count_launches(1);
    e = cudaGetLastError();
    cudaFreeAsync(dev, st);
    if (e != cudaSuccess) return fb_failf(FB_ECUDA, "fbk_resize_area_frac: %s", cudaGetErrorString(e));
    return FB_OK;
}

extern "C" long long fb_masked_dog_f64_workspace(int n, int h, int w)
{
    if (n < 0 || h < 1 || w < 1) return 0;
    return 2 * (long long)n * h * w * (long long)sizeof(double) + (long long)(kMinmaxBlocks + 1) * 2 * sizeof(double) + 256;
}

extern "C" int fb_masked_dog_f64(const double* img, const unsigned char* mask, int n, int h, int w, int mask_n,
                                 double sigma, double ptp, int flags, double* out, void* work, long long work_bytes,
                                 int device, void* stream)
{
    if (n < 0 || h < 1 || w < 1) return fb_failf(FB_EINVAL, "bad shape n=%d %dx%d", n, h, w);
    if (mask && mask_n != 1 && mask_n != n) return fb_failf(FB_EINVAL, "mask_n must be 1 or n");
    if (n == 0) return FB_OK;
    if (!img || !out || !work) return fb_failf(FB_EINVAL, "null pointer");
    if (work_bytes < fb_masked_dog_f64_workspace(n, h, w)) return fb_failf(FB_EINVAL, "workspace too small (%lld < %lld)", work_bytes, fb_masked_dog_f64_workspace(n, h, w));
    Taps64 t{};
    if (!make_taps64(sigma, t)) return fb_failf(FB_ESIZE, "sigma %g outside (0, %g]", sigma, (kMaxRadius64 + 0.49) / 4.0);
    int rc = check_device_ext(device);
    if (rc != FB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long count = (long long)n * h * w;
    double* a = reinterpret_cast<double*>(work);          // first blur
    double* b = a + count;                                // scratch
    double* table = b + count;                            // [kMinmaxBlocks][2] + [2]
    dim3 grid((w + 255) / 256, h, n);
    const int eb = (int)((count + 255) / 256);
    const int unsigned_out = (flags & FB_DOG_UNSIGNED) ? 1 : 0;
    // img0f = G_y(G_x(img)); img1f = G_y(G_x(img0f)); imgf = img0f - img1f
    fbk_gauss1d_f64<0, false><<<grid, 256, 0, st>>>(img, nullptr, 0, nullptr, 0.0, n, h, w, b, t);
    fbk_gauss1d_f64<1, false><<<grid, 256, 0, st>>>(b, nullptr, 0, nullptr, 0.0, n, h, w, a, t);
    fbk_gauss1d_f64<0, false><<<grid, 256, 0, st>>>(a, nullptr, 0, nullptr, 0.0, n, h, w, b, t);
    fbk_gauss1d_f64<1, false><<<grid, 256, 0, st>>>(b, nullptr, 0, nullptr, 0.0, n, h, w, out, t);
    fbk_sub_f64<<<eb, 256, 0, st>>>(a, out, count, (!mask && unsigned_out) ? 1 : 0, out);
    fb_count_launches(5);
    FB_CU(cudaGetLastError());
    if (!mask) return FB_OK;
    const double* span = nullptr;
    if (ptp != ptp) {
        fbk_minmax_f64<<<kMinmaxBlocks, 512, 0, st>>>(img, count, table);
        fbk_minmax_fold_f64<<<1, 32, 0, st>>>(table, kMinmaxBlocks, table + 2 * kMinmaxBlocks);
        fb_count_launches(2);
        span = table + 2 * kMinmaxBlocks;
    }
    const double sigma_c = sqrt(sigma * sigma + sigma * sigma);
    Taps64 tc{};
    if (!make_taps64(sigma_c, tc)) return fb_failf(FB_ESIZE, "sigma %g too large for the mask term", sigma);
    const long long mstride = mask_n == 1 ? 0 : (long long)h * w;
    fbk_gauss1d_f64<0, true><<<grid, 256, 0, st>>>(nullptr, mask, mstride, span, ptp, n, h, w, b, tc);
    fbk_gauss1d_f64<1, false><<<grid, 256, 0, st>>>(b, nullptr, 0, nullptr, 0.0, n, h, w, a, tc);
    fbk_mask_apply_f64<<<eb, 256, 0, st>>>(a, count, sigma_c * sigma_c, sigma * sigma, unsigned_out, out);
    fb_count_launches(3);
    FB_CU(cudaGetLastError());
    return FB_OK;
}
