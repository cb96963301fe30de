// Microbenchmark: FFMA issue rate per SM for three operand forms (register, uniform-register/constant,
// immediate) at several resident-warp counts.  Build: nvcc -arch=sm_100a -O3 ffma_rate.cu -o ffma_rate
#include <cstdio>
#include <cuda_runtime.h>
struct C { float c[16]; };
template <int MODE>
__global__ void k(const __grid_constant__ C cc, float *out, int iters, const float *regc)
{
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = threadIdx.x * 0.001f + j;
    float x = out[threadIdx.x & 7];
    float r[16];
#pragma unroll
    for (int j = 0; j < 16; j++) r[j] = regc[j];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (MODE == 0) a[j] = fmaf(r[j], x, a[j]);          // three register operands (x reused)
                else if (MODE == 1) a[j] = fmaf(cc.c[j], x, a[j]);  // constant-bank / uniform register operand
                else a[j] = fmaf(a[j], 1.0001f, x);                 // immediate operand
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int warps_per_sm)
{
    C cc; for (int j = 0; j < 16; j++) cc.c[j] = 1e-3f * j;
    float *out, *regc; cudaMalloc(&out, 148 * 2048 * 4); cudaMalloc(&regc, 64); cudaMemset(out, 0, 148 * 2048 * 4); cudaMemset(regc, 0, 64);
    int threads = warps_per_sm * 32, iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(cc, out, 10, regc);
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(cc, out, iters, regc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ffma = 148.0 * threads * iters * 64.0;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-10s warps/SM %2d: %.2f TFLOP/s  (%.1f FFMA lanes/clk/SM at %.0f MHz nominal)\n", name, warps_per_sm, 2 * ffma / ms / 1e9,
           ffma / (ms * 1e-3) / 148.0 / (clk * 1e3), clk / 1e3);
    cudaFree(out); cudaFree(regc);
}
int main()
{
    for (int w : {4, 8, 16, 32}) { run<0>("reg", w); run<1>("const/UR", w); run<2>("imm", w); }
    return 0;
}
