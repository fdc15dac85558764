// Probe: register layout of tcgen05.ld.16x256b (.x1/.x4) and whether the 16-lane window may start at lane +16 of a warp's quadrant.
// TMEM is filled with value = lane * 1000 + column through tcgen05.st.32x32b (lane = quadrant * 32 + thread), then read back.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tmem_ld16_probe tmem_ld16_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(float* out, int lane_off) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    const uint32_t trow = base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; i++) v[i] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + i));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(trow + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                     "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // 16x256b.x4: 16 lanes x 32 columns -> 16 registers per thread
    uint32_t r[16];
    const uint32_t addr = base + ((uint32_t)(warp * 32 + lane_off) << 16);
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; i++) out[tid * 16 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64) : "memory");
}
int main() {
    float* d; cudaMalloc(&d, 128 * 16 * 4);
    float h[128 * 16];
    for (int off = 0; off <= 16; off += 16) {
        k<<<1, 128>>>(d, off);
        cudaError_t e = cudaDeviceSynchronize();
        printf("lane_off=%d: %s\n", off, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        // check hypothesis: reg[4j + 2h + b] of thread t = lane (q*32 + off + t/4 + 8h), column 8j + 2(t%4) + b
        int bad = 0;
        for (int t = 0; t < 128; t++) for (int j = 0; j < 4; j++) for (int hh = 0; hh < 2; hh++) for (int b = 0; b < 2; b++) {
            int q = t >> 5, l = t & 31;
            float want = (float)((q * 32 + off + l / 4 + 8 * hh) * 1000 + 8 * j + 2 * (l % 4) + b);
            if (h[t * 16 + 4 * j + 2 * hh + b] != want) bad++;
        }
        printf("  hypothesis reg[4j+2h+b] = (lane q*32+off+t/4+8h, col 8j+2(t%%4)+b): %s (%d mismatches)\n", bad ? "NO" : "YES", bad);
        for (int t = 0; t < 6; t++) { printf("  t%d:", t); for (int i = 0; i < 16; i++) printf(" %.0f", h[t * 16 + i]); printf("\n"); }
        printf("  t37:"); for (int i = 0; i < 16; i++) printf(" %.0f", h[37 * 16 + i]); printf("\n");
    }
    return 0;
}
