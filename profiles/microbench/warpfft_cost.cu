// Microbenchmark: on-chip cost of one 1024-point warp FFT (fb_xcorr_fast.cuh WarpFFT<32,32>) split into
// its parts, 16 warps per SM, no global traffic inside the loop.
//   mode 0: 2 x radix-32 butterflies only (registers)
//   mode 1: + stage twiddles (shared-memory table)
//   mode 2: + transpose through shared memory (= WarpFFT::run)
//   mode 3: mode 2 with scalar (non-packed) butterflies
//   mode 4: 2 x radix-32 butterflies, decimation in time with FMA-fused twiddles (PDitFFT)
//   mode 5: mode 0 with the decimation-in-frequency packed butterflies (PRegFFT), for comparison when FB_DIT=1
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../feabas_b200/csrc warpfft_cost.cu -o warpfft_cost
#include <cstdio>
#include <cuda_runtime.h>
#include "fb_xcorr_fast.cuh"
using namespace fb;
constexpr int E = 32, T = 32, NW = 8, ITER = 256;

template <int MODE>
__global__ void __launch_bounds__(32 * NW, 2) k(const cx<float>* table, cx<float>* out, int iters)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using W = WarpFFT<E, T>;
    cx<float>* regions = reinterpret_cast<cx<float>*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, t = tid & 31;
    StageTw<E, T> tw;
    tw.init(table, regions + NW * W::RS, t, tid, 32 * NW);
    cx<float>* region = regions + warp * W::RS;
    cx<float> v[E];
#pragma unroll
    for (int j = 0; j < E; ++j) v[j] = mk<float>(0.001f * (tid + j), 0.002f * (tid - j));
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 2) {
            W::template run<false>(v, region, tw, t);
        } else if (MODE == 3) {
            RegFFT<float, E, false>::run(v);
            tw.apply_all(v, [&](int k1, cx<float> a) { region[k1 * (T + 1) + t] = a; });
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < T; ++n2) v[n2] = region[t * (T + 1) + n2];
            __syncwarp();
            RegFFT<float, T, false>::run(v);
        } else if (MODE == 4) {
            PDitFFT<E>::run(v);
            PDitFFT<E>::run(v);
        } else if (MODE == 5) {
            PRegFFT<E>::run(v);
            PRegFFT<E>::run(v);
        } else {
            LaneFFT<E>::run(v);
            if (MODE == 1) {
                cx<float> u[E];
                tw.apply_all(v, [&](int k1, cx<float> a) { u[k1] = a; });
#pragma unroll
                for (int j = 0; j < E; ++j) v[j] = u[j];
            }
            LaneFFT<E>::run(v);
        }
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = mk<float>(v[j].x * 0.03125f, v[j].y * 0.03125f);   // keep the values finite
    }
    cx<float> s = mk<float>(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < E; ++j) s = s + v[j];
    out[blockIdx.x * blockDim.x + tid] = s;
}

template <int MODE> void run(const char* name, const cx<float>* table, cx<float>* out)
{
    const size_t sm = ((size_t)NW * (1024 + 32 + 16) + 1024) * sizeof(cx<float>);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const int grid = 148 * 2;
    k<MODE><<<grid, 32 * NW, sm>>>(table, out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 32 * NW, sm>>>(table, out, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    const double ffts = (double)grid * NW * ITER;
    // cycles per FFT per SM sub-partition at 1.965 GHz: 4 warps share an SMSP
    printf("%-44s %8.3f ms  %7.1f ns/FFT/SM  %7.0f cycles per FFT per SMSP (@1.965 GHz)  %s\n", name, ms,
           ms * 1e6 / (ffts / 148), ms * 1e-3 * 1.965e9 / (ffts / (148 * 4)), err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main()
{
    cx<float>* table; cx<float>* out;
    cudaMalloc(&table, 1024 * sizeof(cx<float>)); cudaMalloc(&out, 148 * 2 * 256 * sizeof(cx<float>));
    cx<float> h[1024];
    for (int i = 0; i < 1024; ++i) { h[i].x = 0.9f; h[i].y = 0.1f; }
    cudaMemcpy(table, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("2 x radix-32 butterflies (LaneFFT default)", table, out);
    run<1>("+ stage twiddles (smem table)", table, out);
    run<2>("+ transpose via smem = WarpFFT::run", table, out);
    run<3>("WarpFFT with scalar butterflies", table, out);
    run<4>("2 x radix-32 DIT butterflies (FMA twiddles)", table, out);
    run<5>("2 x radix-32 DIF butterflies (packed)", table, out);
    return 0;
}
