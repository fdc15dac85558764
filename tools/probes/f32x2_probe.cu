// Throughput probe: scalar fma.rn.f32 vs packed fma.rn.f32x2 on sm_100a (issue-bound epilogue math of the fused kernels).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu ; run on the GPU box
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    unsigned long long x[8], A, B;
    asm("mov.b64 %0, {%1, %1};" : "=l"(A) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(B) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo = threadIdx.x * 0.001f + 2 * i, hi = lo + 1; asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi)); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 512 * 4);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    const int iters = 20000;
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(s); k_scalar<<<148 * 4, 512>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e); cudaEventSynchronize(e);
        cudaEventElapsedTime(&ms, s, e);
        double fma = 148.0 * 4 * 512 * 16 * iters;
        printf("scalar fma.f32   : %.3f ms  %.1f GFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.965);
        cudaEventRecord(s); k_packed<<<148 * 4, 512>>>(out, iters, 0.999f, 0.001f); cudaEventRecord(e); cudaEventSynchronize(e);
        cudaEventElapsedTime(&ms, s, e);
        printf("packed fma.f32x2 : %.3f ms  %.1f GFMA/s  (%.1f FMA/clk/SM)\n", ms, fma / ms / 1e6, fma / ms / 1e6 / 148 / 1.965);
    }
    return 0;
}
