// Tensor-core decision for the column kernel: how fast could a radix-32 DFT stage run as a GEMM?
//
// A 1024-point line = two stages of 32 DFTs of length 32.  One stage as a real GEMM: [Wr -Wi; Wi Wr] (64 x 64) times
// [Xr; Xi] (64 x 32) = 64 x 64 x 32 MACs = 128 mma.m16n8k8; float32 parity needs the 3 x TF32 split (a_hi b_hi +
// a_hi b_lo + a_lo b_hi; accuracy in dft_tf32_accuracy.py) = 384 mma per stage, 768 per line.  This microbenchmark
// measures the best case: operands already in registers (no fragment loads, no transposes, no twiddles), 16 warps / SM.
//   cycles per line per SM >= 768 / (mma per clock per SM)      vs. the SIMT warp FFT: 1208 / 4 = 302 (warpfft_cost.cu)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dft_mma dft_mma.cu && ./dft_mma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ILP independent accumulator tiles per warp, `iters` rounds
template <int ILP>
__global__ void __launch_bounds__(512) k_mma(float* out, int iters, long long* cycles)
{
    unsigned a[4], b[2];
    float d[ILP][4];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + 0.001f * (threadIdx.x + i));
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + 0.002f * (threadIdx.x + i));
    for (int j = 0; j < ILP; ++j)
        for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) mma_tf32(d[j], a, b);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < ILP; ++j)
        for (int i = 0; i < 4; ++i) s += d[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        k_mma<8><<<sms, 512>>>(out, iters, cyc);
        cudaDeviceSynchronize();
    }
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double mma_per_sm = 16.0 * 8 * iters;          // 16 warps x ILP x iters
    const double per_clk = mma_per_sm / (double)mx;
    printf("SMs %d, mma.m16n8k8.tf32 per clock per SM: %.3f  (%.0f FMA / clk / SM)\n", sms, per_clk, per_clk * 1024);
    printf("3 x TF32 radix-32 x 2 stages = 768 mma per 1024-point line -> >= %.0f cycles per line per SM with operands in registers\n", 768.0 / per_clk);
    printf("SIMT warp FFT (warpfft_cost.cu): 1208 cycles per line per SM sub-partition = 302 per SM\n");
    return cudaGetLastError() != cudaSuccess;
}
