// Microbenchmark: packed FP32 (f32x2) issue rates on sm_100a next to the scalar forms.
// One "op" below = one SASS instruction per warp; the packed forms carry two FP32 lanes each.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 fp32x2_issue.cu -o fp32x2_issue
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, float2 a, float2 b, long long* cyc)
{
    float2 r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = __ffma2_rn(r[i], a, b);                                  // FFMA2 with 2 loop-invariant operands
            if (MODE == 1) r[i] = __fadd2_rn(r[i], r[(i + 1) & 15]);                        // FADD2
            if (MODE == 2) r[i] = __fmul2_rn(r[i], r[(i + 5) & 15]);                        // FMUL2
            if (MODE == 3) r[i] = __ffma2_rn(r[(i + 1) & 15], r[(i + 2) & 15], r[i]);       // FFMA2 3 distinct
            if (MODE == 4) { r[i].x = r[i].x + r[(i + 1) & 15].x; r[i].y = r[i].y + r[(i + 1) & 15].y; }   // 2 x FADD (scalar)
            if (MODE == 5) { r[i].x = fmaf(r[(i + 1) & 15].x, r[(i + 2) & 15].x, r[i].x); r[i].y = fmaf(r[(i + 1) & 15].y, r[(i + 2) & 15].y, r[i].y); }
            if (MODE == 6) { if (i & 1) r[i] = __fadd2_rn(r[i], r[(i + 1) & 15]); else r[i] = __ffma2_rn(r[(i + 1) & 15], r[(i + 2) & 15], r[i]); }
            if (MODE == 7) {   // complex multiply by a per-register twiddle w = r[(i+3)&15]: (x + iy) w, packed form
                float2 w = r[(i + 3) & 15], v = r[i];
                float2 t = __fmul2_rn(make_float2(v.x, v.x), w);                 // (vx wx, vx wy)
                r[i] = __ffma2_rn(make_float2(-v.y, v.y), make_float2(w.y, w.x), t);   // (-vy wy + vx wx, vy wx + vx wy)
            }
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 16; ++i) { s.x += r[i].x; s.y += r[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int ctas_per_sm, double flops_per_op)
{
    float2* out; long long* cyc;
    int nb = 148 * ctas_per_sm;
    cudaMalloc(&out, nb * 256 * 8); cudaMalloc(&cyc, nb * 8);
    k<MODE><<<nb, 256>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 8]; cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
    double ops = (double)ITER * 16 * 8 * ctas_per_sm;     // loop-body "element updates" per SM (warp granularity)
    printf("%-34s ctas/SM %d  %.2f updates/cycle/SM (clock64)  %.1f G updates/s chip  %.1f TFLOP/s\n", name, ctas_per_sm,
           ops / h[0], ops * 148 / (ms * 1e-3) / 1e9, ops * 148 * 32 * flops_per_op / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int c : {1, 2, 4}) {
        run<0>("FFMA2 r,a,b (2 invariant)", c, 4);
        run<1>("FADD2 r,r", c, 2);
        run<2>("FMUL2 r,r", c, 2);
        run<3>("FFMA2 3 distinct", c, 4);
        run<4>("2x FADD scalar", c, 2);
        run<5>("2x FFMA scalar 3 distinct", c, 4);
        run<6>("FADD2/FFMA2 alternating", c, 3);
        run<7>("complex mul packed (FMUL2+FFMA2)", c, 6);
    }
    return 0;
}
