// Training backward of the fused encoding + SDF MLP (sm_100a, tcgen05): given d loss / d {sdf, d sdf/dx, geom}
// per sample it produces the lattice gradient (scatter-add fused in the kernel), the bias gradients and the weight
// gradients -- ONE kernel, the weight-gradient products are formed on the tensor cores from the operand tiles while they
// are still in shared memory and accumulate in TMEM over all tiles of the CTA.
//
// What it replaces in the reference: `loss.backward()` through SDF.get_sdf_and_gradient
// (permuto_sdf_py/models/models.py:199-259, create_graph=True) -- i.e. the double backward of the
// encoding (positions-gradient -> lattice) and of the 4-layer GELU MLP, ~60 PyTorch kernels per call.
//
// Math. The forward is y = MLP(enc(x)) (sdf = y0, geom = y1..), g = d sdf/dx. For fixed upstream gradients
// (ybar, gbar) the loss depends on the parameters through y and through the scalar s = gbar . g = D_v sdf, the
// directional derivative of sdf along v = gbar. So ONE tangent stream along v is enough (the forward kernel needs
// three because it must output g itself). With a_0 = enc(x), ta_0 = D_v enc(x) and for l = 0..3 (W_3 = output layer)
//     z_{l+1} = W_l a_l + b_l,  a_{l+1} = gelu(z_{l+1});     tz_{l+1} = W_l ta_l,  ta_{l+1} = gelu'(z_{l+1}) tz_{l+1},
// reverse mode gives (zbar_4 = ybar, tzbar_4 = e_0):
//     abar_l = W_l^T zbar_{l+1},   tabar_l = W_l^T tzbar_{l+1}
//     tzbar_l = gelu'(z_l) tabar_l
//     zbar_l  = gelu'(z_l) abar_l + gelu''(z_l) tz_l tabar_l
//     dW_l = zbar_{l+1} a_l^T + tzbar_{l+1} ta_l^T,   db_l = sum zbar_{l+1}
//     lattice[lvl][idx_r] += window_lvl (B_r abar_0[lvl] + dB_r tabar_0[lvl])       (B barycentric weights, dB their tangent)
//
// Per 128-sample tile (512 threads: row = tid & 127, group g = tid >> 7 owns operand cores g, g+4, .. in the encoder phases
// and the 16-column chunk g in the epilogues). Two 64 KB operand-tile buffers X, Y ([value hi | value lo | tangent hi | tangent lo]):
//   1. encoder: a_0, ta_0 -> X; X leaves by ONE TMA store (64 KB / tile, re-read from L2 in step 4: the only spill left)
//   2. forward recompute, layers 0..2: z_1, z_2 (value + tangent) STAY in TMEM (4 x 64 columns); a_1 -> Y, a_2 -> X, a_3 -> Y;
//      gelu'(z_3), gelu''(z_3) tz_3 stay in registers (layer 3 is reversed right away)
//   3. seed zbar_4 -> X
//   4. reverse sweep l = 3..0:  abar_l = zbar_{l+1} W_l uses the FORWARD weight tile as an MN-major B operand (no transposed
//      copy, weights resident for the CTA's lifetime); while it runs, a_l / ta_l are rebuilt into Y from the TMEM-resident z_l
//      (l = 2, 1) or come back by TMA (l = 0); then dW_l += zbar_{l+1}^T a_l: the sample axis is the MMA K dimension and a
//      K-major operand tile IS an MN-major operand for that product, so X and Y are multiplied as they sit. The four dW
//      accumulators (M = 64) share 2 x 64 TMEM columns: an M = 64 instruction only writes lanes (m / 16) * 32 + m % 16, so a
//      second accumulator lives at lane offset 16 of the same columns (validated on B200: tests/test_umma_probe_gpu.py).
//   5. encoder backward: warp-aggregated red.global.add.v2.f32 into the lattice gradient.
// TMEM map (512 columns): [0,256) z_1 | tz_1 | z_2 | tz_2, [256,384) work (value | tangent), [384,512) dW_0/dW_1, dW_2/dW_3.
#include "fused_common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf_fused;

namespace {
constexpr int kBwdThreads = 512;
constexpr int kGroups = kBwdThreads / kTile;
constexpr int kSetBytes = 4 * kATileBytes;         // one operand tile set [value hi | value lo | tangent hi | tangent lo]
constexpr uint32_t kColWork = 256, kColDw = 384;

struct BwdOut {
    uint8_t* a0_spill;  // [ntiles][64 KB] encoder operand tiles, written and re-read by the same CTA
    float* gW[kNL];     // [N_l, K_l] (+=)
    float* gbias[kNL];  // [N_l]      (+=)
};

// in place for 16 accumulator columns (zz = z + bias): z <- a = gelu(zz), tz <- ta = gelu'(zz) tz; g1 = gelu'(zz), g2t = gelu''(zz) tz
__device__ __forceinline__ void activations16(float* z, float* tz, const float* bias, float* g1, float* g2t) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float zz = z[i] + bias[i];
        const GeluEval ge = gelu_eval(zz);
        const float d1 = fmaf(zz, ge.pdf, ge.cdf);
        g2t[i] = ge.pdf * (2.0f - zz * zz) * tz[i];
        z[i] = zz * ge.cdf;
        tz[i] = d1 * tz[i];
        g1[i] = d1;
    }
}
__device__ __forceinline__ void store_set16(uint8_t* set, int row, int c, const float* a, const float* ta) {
    store8(set, set + kATileBytes, row, 2 * c, a);
    store8(set, set + kATileBytes, row, 2 * c + 1, a + 8);
    store8(set + 2 * kATileBytes, set + 3 * kATileBytes, row, 2 * c, ta);
    store8(set + 2 * kATileBytes, set + 3 * kATileBytes, row, 2 * c + 1, ta + 8);
}

// abar = zbar W for both streams (single thread): A = zbar tile set [128 x Np] K-major, B = forward weight tile [Np rows][Kp] read as an
// MN-major operand (reduction over the rows: 8-row groups are sbo_w apart, the k cores kLBO), result Kp columns at tmem_d (+64: tangent)
__device__ __forceinline__ void issue_reverse(uint32_t tmem_d, const uint8_t* zset, const uint8_t* w_hi, const uint8_t* w_lo, int Kp, int Np) {
    const uint32_t idesc = umma::make_idesc(128, Kp, umma::kFmtBF16) | (1u << 16);
    const uint32_t sbo_w = (Kp / 8) * kLBO;
    const uint64_t dwh0 = umma::make_desc(umma::smem_u32(w_hi), sbo_w, kLBO), dwl0 = umma::make_desc(umma::smem_u32(w_lo), sbo_w, kLBO);
#pragma unroll 1
    for (int s = 1; s >= 0; s--) {                       // tangent stream first: its result is consumed first
        const uint64_t dzh0 = umma::make_desc(umma::smem_u32(zset + s * 2 * kATileBytes), kLBO, kSBO_A);
        const uint64_t dzl0 = umma::make_desc(umma::smem_u32(zset + s * 2 * kATileBytes + kATileBytes), kLBO, kSBO_A);
        for (int kk = 0; kk < Np / 16; kk++) {
            const uint64_t oa = (uint64_t)(kk * ((2 * kLBO) >> 4)), ow = (uint64_t)(kk * ((2 * sbo_w) >> 4));
            umma::mma_bf16(tmem_d + s * 64, dzh0 + oa, dwh0 + ow, idesc, kk > 0 ? 1u : 0u);
            umma::mma_bf16(tmem_d + s * 64, dzh0 + oa, dwl0 + ow, idesc, 1u);
            umma::mma_bf16(tmem_d + s * 64, dzl0 + oa, dwh0 + ow, idesc, 1u);
        }
    }
}
// dW (+)= zbar^T a + tzbar^T ta: both tile sets consumed as MN-major operands, K = the 128 samples, M = 64 accumulator rows
__device__ __forceinline__ void issue_dw(uint32_t tmem_d, const uint8_t* zset, const uint8_t* aset, int Kp, bool clear) {
    const uint32_t idesc = umma::make_idesc_mn(64, Kp, umma::kFmtBF16);
#pragma unroll 1
    for (int s = 0; s < 2; s++) {
        const uint64_t dzh0 = umma::make_desc(umma::smem_u32(zset + s * 2 * kATileBytes), kSBO_A, kLBO);
        const uint64_t dzl0 = umma::make_desc(umma::smem_u32(zset + s * 2 * kATileBytes + kATileBytes), kSBO_A, kLBO);
        const uint64_t dah0 = umma::make_desc(umma::smem_u32(aset + s * 2 * kATileBytes), kSBO_A, kLBO);
        const uint64_t dal0 = umma::make_desc(umma::smem_u32(aset + s * 2 * kATileBytes + kATileBytes), kSBO_A, kLBO);
        for (int kk = 0; kk < kTile / 16; kk++) {
            const uint64_t o = (uint64_t)(kk * ((2 * kSBO_A) >> 4));          // 16 samples = two 8-sample groups
            umma::mma_bf16(tmem_d, dzh0 + o, dah0 + o, idesc, (clear && s == 0 && kk == 0) ? 0u : 1u);
            umma::mma_bf16(tmem_d, dzh0 + o, dal0 + o, idesc, 1u);
            umma::mma_bf16(tmem_d, dzl0 + o, dah0 + o, idesc, 1u);
        }
    }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
k_sdf_fused_backward(FusedParams P, const float* __restrict__ pos, const float2* __restrict__ lattice, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ window, const uint8_t* __restrict__ blob,
                     const float* __restrict__ g_sdf, const float* __restrict__ g_grad, const float* __restrict__ g_geom,
                     float* __restrict__ grad_lattice, BwdOut out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_w = smem;                                   // forward operand blob (weights hi/lo + biases), resident
    uint8_t* s_X = smem + P.g.total;
    uint8_t* s_Y = s_X + kSetBytes;
    LevelC* lc = reinterpret_cast<LevelC*>(s_Y + kSetBytes);
    float* s_gb = reinterpret_cast<float*>(lc + 1);        // 4 x 64 bias-gradient accumulators of this CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_gb + kNL * 64);   // [0] TMA loads, [1] mma, [2] dW mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    float* s_x = reinterpret_cast<float*>(s_X);            // exchange tile abar_0 | tabar_0 : [128][2*Kp0+1] fp32 (aliases X, Y)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, grp = tid >> 7;
    const int level_cores = P.L / 4;
    const int K0 = P.g.Kp[0];

    if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::mbar_init(&bars[2], 1); umma::mbar_fence_init(); }
    for (int i = tid; i < P.L * 3; i += kBwdThreads) {
        lc->scale[(i / 3) * 4 + (i % 3)] = scale[i];
        lc->shift[(i / 3) * 4 + (i % 3)] = shift ? shift[i] : 0.0f;
    }
    for (int i = tid; i < P.L; i += kBwdThreads) lc->window[i] = window ? window[i] : 1.0f;
    for (int i = tid; i < kNL * 64; i += kBwdThreads) s_gb[i] = 0.0f;
    __syncthreads();
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_expect_tx(&bars[0], (uint32_t)P.g.total);
        umma::bulk_g2s(s_w, blob, (uint32_t)P.g.total, &bars[0]);
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);     // this warp's 32 TMEM lanes
    umma::mbar_wait(&bars[0], 0);
    uint32_t ld_phase = 1, mma_phase = 0, dw_phase = 0;

    const int ntiles = (P.N + kTile - 1) / kTile;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, first_tile = false) {
        const int n = tile * kTile + row;
        const bool valid = n < P.N;
        float x[3], v[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            x[i] = valid ? pos[(size_t)n * 3 + i] : 0.0f;
            v[i] = (valid && g_grad) ? g_grad[(size_t)n * 3 + i] : 0.0f;
        }
        // ---------------- 1. encoder -> X: operand cores grp, grp+4, ... (4 levels = 8 features = one 16-byte core row)
        for (int kc = grp; kc < K0 / 8; kc += kGroups) {
            float fv[8], ft[8];
            if (kc < level_cores) {
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int l = kc * 4 + q;
                    float cf[3], dcf[3], e[4];
#pragma unroll
                    for (int i = 0; i < 3; i++) {
                        cf[i] = __fmul_rn(__fadd_rn(x[i], lc->shift[l * 4 + i]), lc->scale[l * 4 + i]);
                        dcf[i] = v[i] * lc->scale[l * 4 + i];
                    }
                    elevate3(cf, e);
                    Simplex3 s;
                    locate3(e, s);
                    float db[5];
                    bary_tangent3(dcf, s, db);
                    const float2* tab = lattice + (size_t)l * P.T;
                    float2 val[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) val[r] = __ldg(tab + vindex3(s, r, P.cap_mask, (unsigned)P.T));
                    const float w = lc->window[l];
                    float a0 = 0.f, a1 = 0.f, t0 = 0.f, t1 = 0.f;
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        float wr = s.bary[r] * w, cr = db[r] * w;
                        a0 = fmaf(val[r].x, wr, a0); a1 = fmaf(val[r].y, wr, a1);
                        t0 = fmaf(val[r].x, cr, t0); t1 = fmaf(val[r].y, cr, t1);
                    }
                    fv[2 * q] = a0; fv[2 * q + 1] = a1; ft[2 * q] = t0; ft[2 * q + 1] = t1;
                }
            } else {        // concat-points columns + zero padding
                const int c0 = 2 * P.L;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    int c = kc * 8 + i - c0;
                    float a = 0.f, t = 0.f;
#pragma unroll
                    for (int d = 0; d < 3; d++) if (c == d && (c0 + c) < P.in_dim) { a = x[d] * P.points_scaling; t = v[d] * P.points_scaling; }
                    fv[i] = a; ft[i] = t;
                }
            }
            store8(s_X, s_X + kATileBytes, row, kc, fv);
            store8(s_X + 2 * kATileBytes, s_X + 3 * kATileBytes, row, kc, ft);
        }

        // ---------------- 2. forward recompute, layers 0..2. z_{l+1} | tz_{l+1} -> TMEM columns l * 128 (l = 0, 1) or the work columns (l = 2)
        float g1[16], g2t[16];                 // gelu'(z_3), gelu''(z_3) tz_3 of this thread's chunk: consumed by the first reverse step
        float seed[2][8];                      // upstream gradient columns of this thread (prefetched behind the layer-2 MMAs)
#pragma unroll 1
        for (int l = 0; l < 3; l++) {
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            const uint8_t* at = (l == 1) ? s_Y : s_X;
            uint8_t* dst = (l == 1) ? s_X : s_Y;
            const uint32_t col = (l == 2) ? kColWork : (uint32_t)(l * 128);
            if (tid == 0) {
                umma::fence_after_sync();
                if (l == 0) {                  // a_0 | ta_0 leave as they are (needed again for dW_0 at the end of the reverse sweep)
                    umma::bulk_s2g(out.a0_spill + (size_t)tile * kSetBytes, s_X, kSetBytes);
                    umma::bulk_commit();
                }
                for (int s = 0; s < 2; s++)
                    issue_gemm(tmem_base + col + s * 64, at + s * 2 * kATileBytes, at + s * 2 * kATileBytes + kATileBytes, s_w + P.g.w_hi[l],
                               s_w + P.g.w_lo[l], P.g.Kp[l], P.g.Np[l]);
                if (l == 1) umma::bulk_wait_read0();       // X is overwritten by this layer's epilogue: the store must have read it
                umma::commit(&bars[1]);
            }
            if (l == 2) {                      // zbar_4 = [g_sdf, g_geom...] columns of this thread, loaded while the MMAs run
                const int Np = P.g.Np[3], nout = P.g.N[3];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = grp + j * kGroups;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int colz = c * 8 + i;
                        float vz = 0.f;
                        if (valid && c < Np / 8) {
                            if (colz == 0) vz = g_sdf ? g_sdf[n] : 0.f;
                            else if (colz < nout) vz = g_geom ? g_geom[(size_t)n * (nout - 1) + colz - 1] : 0.f;
                        }
                        seed[j][i] = vz;
                    }
                }
            }
            umma::mbar_wait(&bars[1], mma_phase);
            mma_phase ^= 1;
            umma::fence_after_sync();
            const int c = grp;
            if (c < P.g.Np[l] / 16) {
                float z[16], tz[16];
                umma::tmem_ld16(tlane + col + c * 16, z);
                umma::tmem_ld16(tlane + col + 64 + c * 16, tz);
                umma::tmem_ld_wait();
                activations16(z, tz, reinterpret_cast<const float*>(s_w + P.g.bias[l]) + c * 16, g1, g2t);
                store_set16(dst, row, c, z, tz);
            }
        }
        // ---------------- 3. seed: zbar_4 = upstream gradient, tzbar_4 = e_0 -> X (a_2 in X is dead: the layer-2 MMAs completed)
        {
            const int Np = P.g.Np[3];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int c = grp + j * kGroups;
                if (c < Np / 8) {
                    float tb[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) tb[i] = (valid && c == 0 && i == 0) ? 1.0f : 0.0f;
                    store8(s_X, s_X + kATileBytes, row, c, seed[j]);
                    store8(s_X + 2 * kATileBytes, s_X + 3 * kATileBytes, row, c, tb);
                    float z16[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) { z16[i] = seed[j][i]; z16[i + 8] = 0.f; }
                    int colb;
                    const float cs = colsum16(z16, lane, colb);                       // bias gradient of the output layer
                    if (!(lane & 1) && colb < 8) atomicAdd(&s_gb[3 * 64 + c * 8 + colb], cs);
                }
            }
        }
        // ---------------- 4. reverse sweep, layers 3..0: X = zbar_{l+1} | tzbar_{l+1}, Y = a_l | ta_l
#pragma unroll 1
        for (int l = 3; l >= 0; l--) {
            const uint32_t dw_col = tmem_base + kColDw + (uint32_t)(l >> 1) * 64 + ((l & 1) ? (16u << 16) : 0u);
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                issue_reverse(tmem_base + kColWork, s_X, s_w + P.g.w_hi[l], s_w + P.g.w_lo[l], P.g.Kp[l], P.g.Np[l]);
                umma::commit(&bars[1]);
                if (l == 3) {                  // a_3 | ta_3 are in Y since the forward epilogue
                    issue_dw(dw_col, s_X, s_Y, P.g.Kp[l], first_tile);
                    umma::commit(&bars[2]);
                } else if (l == 0) {           // a_0 | ta_0 come back by TMA (requested after the previous step released Y)
                    umma::mbar_wait(&bars[0], ld_phase);
                    umma::fence_after_sync();
                    issue_dw(dw_col, s_X, s_Y, P.g.Kp[l], first_tile);
                    umma::commit(&bars[2]);
                }
            }
            if (l == 1 || l == 2) {
                // while the reverse MMAs run: rebuild a_l | ta_l -> Y and the GELU' / GELU'' factors from the TMEM-resident z_l, tz_l
                const int c = grp;
                if (c < P.g.Kp[l] / 16) {
                    float z[16], tz[16];
                    const uint32_t col = (uint32_t)((l - 1) * 128);
                    umma::tmem_ld16(tlane + col + c * 16, z);
                    umma::tmem_ld16(tlane + col + 64 + c * 16, tz);
                    umma::tmem_ld_wait();
                    activations16(z, tz, reinterpret_cast<const float*>(s_w + P.g.bias[l - 1]) + c * 16, g1, g2t);
                    store_set16(s_Y, row, c, z, tz);
                }
                umma::fence_async_smem();
                umma::fence_before_sync();
                __syncthreads();
                if (tid == 0) {
                    umma::fence_after_sync();
                    issue_dw(dw_col, s_X, s_Y, P.g.Kp[l], first_tile);
                    umma::commit(&bars[2]);
                }
            }
            umma::mbar_wait(&bars[1], mma_phase);
            mma_phase ^= 1;
            umma::fence_after_sync();
            const int c = grp;
            const bool have = c < P.g.Kp[l] / 16;
            float ab[16], tab_[16];
            if (have) {
                umma::tmem_ld16(tlane + kColWork + c * 16, ab);
                umma::tmem_ld16(tlane + kColWork + 64 + c * 16, tab_);
                umma::tmem_ld_wait();
                if (l > 0) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float zb = g1[i] * ab[i] + g2t[i] * tab_[i];
                        tab_[i] = g1[i] * tab_[i];
                        ab[i] = zb;                  // ab = zbar_l, tab_ = tzbar_l
                    }
                }
            }
            // X and Y are still operands of the weight-gradient MMAs
            umma::mbar_wait(&bars[2], dw_phase);
            dw_phase ^= 1;
            umma::fence_after_sync();
            if (l == 1 && tid == 0) {          // Y is free: bring a_0 | ta_0 back for the last step (the store completed long ago)
                umma::bulk_wait0();
                umma::mbar_expect_tx(&bars[0], (uint32_t)kSetBytes);
                umma::bulk_g2s(s_Y, out.a0_spill + (size_t)tile * kSetBytes, (uint32_t)kSetBytes, &bars[0]);
            }
            if (l > 0) {
                if (have) {
                    store_set16(s_X, row, c, ab, tab_);
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 16; i++) ab[i] = 0.f;
                    }
                    int colb;
                    const float cs = colsum16(ab, lane, colb);       // bias gradient: column sums over the warp's 32 rows
                    if (!(lane & 1)) atomicAdd(&s_gb[(l - 1) * 64 + c * 16 + colb], cs);
                }
            } else {
                // abar_0 | tabar_0 for the encoder backward (fp32 exchange tile over X and Y, which no MMA reads any more)
                if (have) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        s_x[row * (2 * K0 + 1) + c * 16 + i] = ab[i];
                        s_x[row * (2 * K0 + 1) + K0 + c * 16 + i] = tab_[i];
                    }
                }
            }
        }
        ld_phase ^= 1;
        umma::fence_before_sync();
        __syncthreads();

        // ---------------- 5. encoder backward: lattice[l][idx_r] += window_l (B_r abar_0[l] + dB_r tabar_0[l])
        {
            const float* xr = s_x + row * (2 * K0 + 1);
            for (int l = grp; l < P.L; l += kGroups) {
                float cf[3], dcf[3], e[4];
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    cf[i] = __fmul_rn(__fadd_rn(x[i], lc->shift[l * 4 + i]), lc->scale[l * 4 + i]);
                    dcf[i] = v[i] * lc->scale[l * 4 + i];
                }
                elevate3(cf, e);
                Simplex3 s;
                locate3(e, s);
                float db[5];
                bary_tangent3(dcf, s, db);
                const float w = lc->window[l];
                const float a0 = xr[2 * l], a1 = xr[2 * l + 1], t0 = xr[K0 + 2 * l], t1 = xr[K0 + 2 * l + 1];
                float* gtab = grad_lattice + (size_t)l * P.T * 2;
                const bool aggregate = lc->scale[l * 4] < kAggregateBelowScale;       // warp-uniform (l is)
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    unsigned idx = vindex3(s, r, P.cap_mask, (unsigned)P.T);
                    float cb = s.bary[r] * w, cd = db[r] * w;
                    float2 cv = make_float2(cb * a0 + cd * t0, cb * a1 + cd * t1);
                    if (aggregate) {
                        unsigned key = valid ? idx : 0xffffffffu;
                        unsigned peers = __match_any_sync(kFull, key);
                        cv = add_peers2(peers, cv, lane);
                        if (valid && lane == __ffs(peers) - 1) red_v2(gtab + (size_t)idx * 2, cv);
                    } else if (valid) {
                        red_v2(gtab + (size_t)idx * 2, cv);
                    }
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();     // exchange tile / TMEM free for the next tile
    }
    // ---------------- weight gradients of this CTA: TMEM -> global (+=). Warp w reads lanes 32 (w & 3) .. +31 of column chunk w >> 2:
    // lanes 0..15 hold rows 16 (w & 3) + lane of the even layer of the pair, lanes 16..31 the same rows of the odd layer
    umma::fence_after_sync();
    if (!first_tile) {
        const int q = warp & 3, c = warp >> 2;
#pragma unroll 1
        for (int p = 0; p < 2; p++) {
            float acc[16];
            umma::tmem_ld16(tlane + kColDw + p * 64 + c * 16, acc);
            umma::tmem_ld_wait();
            const int l = 2 * p + (lane >> 4), m = q * 16 + (lane & 15);
            if (m < P.g.N[l]) {
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int k = c * 16 + i;
                    if (k < P.g.K[l]) atomicAdd(out.gW[l] + (size_t)m * P.g.K[l] + k, acc[i]);
                }
            }
        }
    }
    for (int i = tid; i < kNL * 64; i += kBwdThreads) {
        int l = i >> 6, c = i & 63;
        if (c < P.g.N[l] && s_gb[i] != 0.0f) atomicAdd(out.gbias[l] + c, s_gb[i]);
    }
    if (tid == 0) umma::bulk_wait0();      // spill stores complete before the CTA retires
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {

// workspace: the encoder operand tiles (a_0 | ta_0, 64 KB per 128-sample tile), written by TMA store in the forward recompute and read
// back by the same CTA for the first layer's weight gradient
long long psdf_sdf_fused_backward_workspace_bytes(int N) {
    const int ntiles = div_up(N > 0 ? N : 1, kTile);
    return (long long)ntiles * kSetBytes;
}

// grad_lattice, grad_W_l [N_l, K_l] and grad_bias_l are accumulated (+=).
int psdf_sdf_fused_backward(int N, int L, int T, const float* pos, const float* lattice, const float* scale_factor, const float* shift,
                            const float* window, float points_scaling, int hidden, int out_dim, const uint8_t* blob, const float* g_sdf,
                            const float* g_grad, const float* g_geom, float* grad_lattice, uint8_t* workspace, float* gW0, float* gW1,
                            float* gW2, float* gW3, float* gb0, float* gb1, float* gb2, float* gb3, void* stream) {
    if (N < 0 || L < 4 || L > kMaxLevels || (L % 4) != 0 || hidden > 64 || hidden % 16 != 0 || out_dim > 64) return PSDF_ERR_UNSUPPORTED;
    if (N == 0) return PSDF_OK;
    FusedParams P;
    P.N = N; P.L = L; P.T = T;
    P.cap_mask = ((T & (T - 1)) == 0) ? (unsigned)(T - 1) : 0u;
    P.points_scaling = points_scaling;
    P.in_dim = (L + 2) * 2;
    if (P.in_dim > 64) return PSDF_ERR_UNSUPPORTED;
    P.g = make_geom(P.in_dim, hidden, out_dim);
    BwdOut out;
    float* w[kNL] = {gW0, gW1, gW2, gW3};
    float* b[kNL] = {gb0, gb1, gb2, gb3};
    for (int l = 0; l < kNL; l++) { out.gW[l] = w[l]; out.gbias[l] = b[l]; }
    out.a0_spill = workspace;
    const int ntiles = div_up(N, kTile);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)P.g.total + 2 * kSetBytes + sizeof(LevelC) + kNL * 64 * sizeof(float) + 64;
    if ((size_t)128 * (2 * P.g.Kp[0] + 1) * 4 > (size_t)2 * kSetBytes || smem > 227 * 1024) return PSDF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(k_sdf_fused_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);   // per device: cheap, not cached
    k_sdf_fused_backward<<<min(ntiles, sms), kBwdThreads, smem, ST>>>(P, pos, reinterpret_cast<const float2*>(lattice), scale_factor, shift,
                                                                      window, blob, g_sdf, g_grad, g_geom, grad_lattice, out);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}

}  // extern "C"
