// Hardware probes for the tcgen05 conventions the fused backward kernels rely on (test infrastructure, NOT part of the product
// library: built into tests/probe/libumma_probe.so, loaded only by tests/test_umma_probe_gpu.py).
//   mode 0: two M = 64 accumulators sharing the same TMEM columns, one at lane offset 0 and one at lane offset 16
//           (an M = 64 instruction writes lanes (m / 16) * 32 + m % 16; does a D address with lane field 16 fill the other half?)
//   mode 1: D[128 x Kp] = A[128 x Np] * W[Np x Kp] with A K-major and B = the FORWARD weight tile ([Np rows][Kp] K-major along k)
//           consumed as an MN-major operand (reduction over the rows): the reverse GEMM without a transposed weight copy.
#include "../../permuto_sdf_b200/csrc/fused_common.cuh"

using namespace psdf_fused;

namespace {
__global__ void __launch_bounds__(kTile) k_probe(int mode, const float* __restrict__ A, const float* __restrict__ A2,
                                                const float* __restrict__ B, float* __restrict__ dump) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + kATileBytes;
    uint8_t* a2_hi = a_lo + kATileBytes;
    uint8_t* a2_lo = a2_hi + kATileBytes;
    uint8_t* b_hi = a2_lo + kATileBytes;
    uint8_t* b_lo = b_hi + kATileBytes;
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + kATileBytes);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { umma::mbar_init(bar, 1); umma::mbar_fence_init(); }
    // all three inputs are [128][64] row-major fp32, stored as K-major core-matrix tiles (row = tid)
    for (int kc = 0; kc < 8; kc++) {
        float va[8], va2[8], vb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            va[i] = A[tid * 64 + kc * 8 + i];
            va2[i] = A2[tid * 64 + kc * 8 + i];
            vb[i] = B[tid * 64 + kc * 8 + i];
        }
        store8(a_hi, a_lo, tid, kc, va);
        store8(a2_hi, a2_lo, tid, kc, va2);
        store8(b_hi, b_lo, tid, kc, vb);
    }
    __syncthreads();
    if (warp == 0) umma::tmem_alloc(slot, 64);
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *slot;
    {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; i++) z[i] = -7.0f;        // sentinel: lanes no MMA writes keep it
        for (int c = 0; c < 4; c++) umma::tmem_st16(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16, z);
        umma::tmem_st_wait();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        umma::fence_after_sync();
        if (mode == 0) {
            const uint32_t idesc = umma::make_idesc_mn(64, 64, umma::kFmtBF16);
            for (int which = 0; which < 2; which++) {
                const uint32_t ah = umma::smem_u32(which ? a2_hi : a_hi), al = umma::smem_u32(which ? a2_lo : a_lo);
                const uint32_t bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
                const uint32_t d = tmem_base + (which ? (16u << 16) : 0u);
                for (int kk = 0; kk < kTile / 16; kk++) {
                    const uint32_t ko = kk * 2 * kSBO_A;
                    uint64_t dah = umma::make_desc(ah + ko, kSBO_A, kLBO), dal = umma::make_desc(al + ko, kSBO_A, kLBO);
                    uint64_t dbh = umma::make_desc(bh + ko, kSBO_A, kLBO), dbl = umma::make_desc(bl + ko, kSBO_A, kLBO);
                    umma::mma_bf16(d, dah, dbh, idesc, kk > 0 ? 1u : 0u);
                    umma::mma_bf16(d, dah, dbl, idesc, 1u);
                    umma::mma_bf16(d, dal, dbh, idesc, 1u);
                }
            }
        } else {
            // A = a tile [128 x 64] K-major; B = first 64 rows of the b tile = W[n = 0..63][k = 0..63], forward layout
            const uint32_t idesc = umma::make_idesc(128, 64, umma::kFmtBF16) | (1u << 16);
            const uint32_t ah = umma::smem_u32(a_hi), al = umma::smem_u32(a_lo), bh = umma::smem_u32(b_hi), bl = umma::smem_u32(b_lo);
            for (int kk = 0; kk < 4; kk++) {
                uint64_t dah = umma::make_desc(ah + kk * 2 * kLBO, kLBO, kSBO_A), dal = umma::make_desc(al + kk * 2 * kLBO, kLBO, kSBO_A);
                // reduction index n: 8-row groups of the weight tile are kSBO_A apart (K direction), the k cores kLBO (N direction)
                uint64_t dbh = umma::make_desc(bh + kk * 2 * kSBO_A, kSBO_A, kLBO), dbl = umma::make_desc(bl + kk * 2 * kSBO_A, kSBO_A, kLBO);
                umma::mma_bf16(tmem_base, dah, dbh, idesc, kk > 0 ? 1u : 0u);
                umma::mma_bf16(tmem_base, dah, dbl, idesc, 1u);
                umma::mma_bf16(tmem_base, dal, dbh, idesc, 1u);
            }
        }
        umma::commit(bar);
    }
    umma::mbar_wait(bar, 0);
    umma::fence_after_sync();
    for (int c = 0; c < 4; c++) {
        float z[16];
        umma::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 16, z);
        umma::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) dump[tid * 64 + c * 16 + i] = z[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 64);
}
}  // namespace

extern "C" int umma_probe(int mode, const float* A, const float* A2, const float* B, float* dump, void* stream) {
    size_t smem = 6 * kATileBytes + 64;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<<<1, kTile, smem, (cudaStream_t)stream>>>(mode, A, A2, B, dump);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
