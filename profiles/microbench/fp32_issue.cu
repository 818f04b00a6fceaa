// Microbenchmark: FP32 issue rates on sm_100a (warp-instructions per cycle per SM) for the
// instruction mixes an FFT butterfly uses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, long long* cyc)
{
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = fmaf(r[i], a, b);                 // FFMA reg,reg,reg (2 uniform-ish operands)
            if (MODE == 1) r[i] = r[i] + r[(i + 1) & 15];           // FADD
            if (MODE == 2) r[i] = r[i] * r[(i + 5) & 15];           // FMUL
            if (MODE == 3) r[i] = fmaf(r[(i + 1) & 15], r[(i + 2) & 15], r[i]);   // FFMA 3 distinct regs
            if (MODE == 4) { if (i & 1) r[i] = r[i] + r[(i + 1) & 15]; else r[i] = fmaf(r[(i + 1) & 15], r[(i + 2) & 15], r[i]); }
            if (MODE == 5) r[i] = fmaf(r[i], 0.7071067f, r[(i + 3) & 15]);   // FFMA imm
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int ctas_per_sm)
{
    float* out; long long* cyc;
    int nb = 148 * ctas_per_sm;
    cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&cyc, nb * 8);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double warp_instr = (double)ITER * 16 * 8 * ctas_per_sm;     // per SM
    printf("%-28s ctas/SM %d  cycles(cta0) %lld  -> %.2f warp-instr/cycle/SM (clock64), %.1f Ginstr/s chip (events)\n", name,
           ctas_per_sm, h[0], warp_instr / h[0], warp_instr * 148 / (ms * 1e-3) / 1e9);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int c : {1, 2, 4}) {
        run<0>("FFMA r,a,b (2 const regs)", c);
        run<1>("FADD r,r", c);
        run<2>("FMUL r,r", c);
        run<3>("FFMA 3 distinct regs", c);
        run<4>("FADD/FFMA alternating", c);
        run<5>("FFMA imm", c);
    }
    return 0;
}
