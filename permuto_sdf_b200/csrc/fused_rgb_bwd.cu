// Training backward of the fused colour network (sm_100a, tcgen05): d loss / d (linear rgb) per sample in ->
// lattice gradient (scatter fused), d loss / d sdf-gradient (through the normalisation), d loss / d geom feature,
// weight + bias gradients of the 4 linear layers. Replaces loss.backward() through RGB.forward
// (permuto_sdf_py/models/models.py:395-414) -- encoding backward, 4 x (GELU backward, 2 GEMMs), normalize / cat backward.
//
// Per 128-sample tile (512 threads: row = tid & 127, group = tid >> 7):
//   1. input tile a_0 (encoder gather + SH + normal + geom) as in the forward;
//   2. forward recompute, layers 0..2 on the tensor cores; pre-activations z^(0), z^(1), z^(2) STAY in TMEM (128 + 128 + 64
//      columns); every input tile a_l is spilled (TMA store of the operand tile as it sits in shared memory) for dW;
//   3. reverse sweep l = 3..0: zbar^(l) tile -> MMA with W_l^T -> abar_l -> zbar^(l-1) = abar_l gelu'(z^(l-1)); the zbar tiles are
//      spilled the same way; bias gradients = column sums (shuffles + shared atomics);
//   4. abar_0 -> lattice scatter (warp-aggregated red.global.add.v2.f32), normal -> sdf-gradient, geom gradient.
// Weights stream per layer through one 64 KB buffer by TMA; the next block is requested as soon as the MMAs that read the
// current one have retired, so the copy hides behind the epilogue.
// dW_l = zbar^(l)^T a_l is formed by k_rgb_dw from the spilled tiles: the sample axis is the MMA K dimension and an operand tile
// in the K-major core-matrix layout is an MN-major operand for that product (see fused_sdf_bwd.cu), M = 128 / 64 accumulators
// in TMEM, one red.add of the CTA's partial sums at the end.
#include "fused_common.cuh"
#include "fused_rgb_common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf_fused;
using namespace psdf_rgb;

namespace {
constexpr int kRgbSpillBytes = 2 * kWTileBytes;      // [hi | lo] of one operand tile

struct RgbSpill {
    uint8_t* zt[kNL];    // [ntiles][64 KB] zbar^(l) tiles
    uint8_t* at[kNL];    // [ntiles][64 KB] a_l tiles
    float* gbias[kNL];   // (+=)
};

__global__ void __launch_bounds__(kRgbThreads, 1)
k_rgb_fused_backward(RgbParams P, const float* __restrict__ pos, const float* __restrict__ dirs, const float* __restrict__ sdf_grad,
                     const float* __restrict__ geom, const float2* __restrict__ lattice, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ window, const uint8_t* __restrict__ blob,
                     const float* __restrict__ g_out, float* __restrict__ grad_lattice, float* __restrict__ g_sdf_grad,
                     float* __restrict__ g_geom, RgbSpill sp) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* s_a = smem;                                   // activation tiles [hi | lo]
    uint8_t* s_z = s_a + 2 * kWTileBytes;                  // adjoint tiles [hi | lo]
    uint8_t* s_w = s_z + 2 * kWTileBytes;                  // weights of the current GEMM [hi | lo]
    LevelC* lc = reinterpret_cast<LevelC*>(s_w + 2 * kWTileBytes);
    float* s_bias = reinterpret_cast<float*>(lc + 1);      // 4 x 128
    float* s_gb = s_bias + kNL * 128;                      // 4 x 128 bias-gradient accumulators of this CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_gb + kNL * 128);   // [0] weights, [1] mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_x = reinterpret_cast<float*>(s_a);            // exchange tile abar_0: [128][K0 + 1] fp32, aliases s_a

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, grp = tid >> 7;
    const int K0 = P.g.Kp[0];
    if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::mbar_fence_init(); }
    load_level_consts(lc, P.L, scale, shift, window, tid, kRgbThreads);
    for (int i = tid; i < kNL * 128; i += kRgbThreads) {
        int l = i >> 7, c = i & 127;
        s_bias[i] = (c < P.g.Np[l]) ? reinterpret_cast<const float*>(blob + P.g.bias[l])[c] : 0.0f;
        s_gb[i] = 0.0f;
    }
    __syncthreads();
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_z[3] = {tmem_base, tmem_base + 128, tmem_base + 256};
    const uint32_t tmem_work = tmem_base + 384;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t w_phase = 0, mma_phase = 0;

    const int ntiles = (P.N + kTile - 1) / kTile;
    if (tid == 0 && blockIdx.x < ntiles) load_layer_weights(s_w, blob, P.g, 0, false, &bars[0]);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile * kTile + row;
        const bool valid = n < P.N;
        build_input_tile(P, lc, lattice, pos, dirs, sdf_grad, geom, n, valid, row, grp, s_a, s_a + kWTileBytes);

        // ---------------- forward recompute, layers 0..2
#pragma unroll 1
        for (int l = 0; l < 3; l++) {
            umma::mbar_wait(&bars[0], w_phase);
            w_phase ^= 1;
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                umma::bulk_s2g(sp.at[l] + (size_t)tile * kRgbSpillBytes, s_a, kRgbSpillBytes);      // a_l for dW_l
                umma::bulk_commit();
                issue_gemm_w(tmem_z[l], s_a, s_a + kWTileBytes, s_w, s_w + P.g.Np[l] * P.g.Kp[l] * 2, P.g.Kp[l], P.g.Np[l]);
                umma::bulk_wait_read0();
                umma::commit(&bars[1]);
            }
            umma::mbar_wait(&bars[1], mma_phase);
            mma_phase ^= 1;
            umma::fence_after_sync();
            if (tid == 0) load_layer_weights(s_w, blob, P.g, l < 2 ? l + 1 : 3, l == 2, &bars[0]);   // W_{l+1}, then W_3^T
            for (int c = grp; c < P.g.Np[l] / 16; c += kRgbGroups) {
                float z[16];
                umma::tmem_ld16(tmem_z[l] + lane_off + c * 16, z);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) { const float zz = z[i] + s_bias[l * 128 + c * 16 + i]; z[i] = zz * gelu_eval(zz).cdf; }
                store8w(s_a, s_a + kWTileBytes, row, 2 * c, z);
                store8w(s_a, s_a + kWTileBytes, row, 2 * c + 1, z + 8);
            }
            umma::fence_before_sync();
        }
        // ---------------- seed: zbar^(3) = d loss / d linear rgb (3 of 16 padded columns)
        if (grp == 0) {
            float zb[16];
#pragma unroll
            for (int i = 0; i < 16; i++) zb[i] = (i < 3 && valid) ? g_out[(size_t)n * 3 + i] : 0.f;
            store8w(s_z, s_z + kWTileBytes, row, 0, zb);
            store8w(s_z, s_z + kWTileBytes, row, 1, zb + 8);
            int col;
            const float cs = colsum16(zb, lane, col);
            if (!(lane & 1) && col < 3) atomicAdd(&s_gb[3 * 128 + col], cs);
        }
        // ---------------- reverse sweep, layers 3..0
#pragma unroll 1
        for (int l = 3; l >= 0; l--) {
            umma::mbar_wait(&bars[0], w_phase);
            w_phase ^= 1;
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                umma::bulk_s2g(sp.zt[l] + (size_t)tile * kRgbSpillBytes, s_z, kRgbSpillBytes);       // zbar^(l) for dW_l
                if (l == 3) umma::bulk_s2g(sp.at[3] + (size_t)tile * kRgbSpillBytes, s_a, kRgbSpillBytes);   // a_3
                umma::bulk_commit();
                // abar_l [128 x Kp_l] = zbar^(l) [128 x Np_l] W_l : B operand = W_l^T stored [Kp_l rows][Np_l]
                issue_gemm_w(tmem_work, s_z, s_z + kWTileBytes, s_w, s_w + P.g.Np[l] * P.g.Kp[l] * 2, P.g.Np[l], P.g.Kp[l]);
                umma::bulk_wait_read0();
                umma::commit(&bars[1]);
            }
            // while the MMAs run: gelu'(z^(l-1)) of this thread's column chunks from the TMEM-resident pre-activations
            const int Kp = P.g.Kp[l];
            float g1a[16], g1b[16];                       // chunks grp and grp + 4 (Kp <= 128 -> at most two chunks per thread)
            if (l > 0) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = grp + j * kRgbGroups;
                    if (c < Kp / 16) {
                        float z[16];
                        umma::tmem_ld16(tmem_z[l - 1] + lane_off + c * 16, z);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float zz = z[i] + s_bias[(l - 1) * 128 + c * 16 + i];
                            const GeluEval ge = gelu_eval(zz);
                            (j == 0 ? g1a : g1b)[i] = fmaf(zz, ge.pdf, ge.cdf);
                        }
                    }
                }
            }
            umma::mbar_wait(&bars[1], mma_phase);
            mma_phase ^= 1;
            umma::fence_after_sync();
            if (tid == 0) {
                if (l > 0) load_layer_weights(s_w, blob, P.g, l - 1, true, &bars[0]);
                else if (tile + (int)gridDim.x < ntiles) load_layer_weights(s_w, blob, P.g, 0, false, &bars[0]);
            }
            if (l > 0) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int c = grp + j * kRgbGroups;
                    if (c >= Kp / 16) continue;
                    float ab[16];
                    umma::tmem_ld16(tmem_work + lane_off + c * 16, ab);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) ab[i] *= (j == 0 ? g1a : g1b)[i];
                    store8w(s_z, s_z + kWTileBytes, row, 2 * c, ab);
                    store8w(s_z, s_z + kWTileBytes, row, 2 * c + 1, ab + 8);
                    if (!valid) {
#pragma unroll
                        for (int i = 0; i < 16; i++) ab[i] = 0.f;
                    }
                    int col;
                    const float cs = colsum16(ab, lane, col);
                    if (!(lane & 1)) atomicAdd(&s_gb[(l - 1) * 128 + c * 16 + col], cs);
                }
            } else {
                for (int c = grp; c < Kp / 16; c += kRgbGroups) {
                    float ab[16];
                    umma::tmem_ld16(tmem_work + lane_off + c * 16, ab);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) s_x[row * (K0 + 1) + c * 16 + i] = ab[i];
                }
            }
            umma::fence_before_sync();
        }
        __syncthreads();
        // ---------------- encoder / tail backward from abar_0
        {
            const float* xr = s_x + row * (K0 + 1);
            float x[3] = {0.f, 0.f, 0.f};
            if (valid) { x[0] = pos[(size_t)n * 3]; x[1] = pos[(size_t)n * 3 + 1]; x[2] = pos[(size_t)n * 3 + 2]; }
            for (int l = grp; l < P.L; l += kRgbGroups) {
                float cf[3], e[4];
#pragma unroll
                for (int i = 0; i < 3; i++) cf[i] = __fmul_rn(__fadd_rn(x[i], lc->shift[l * 4 + i]), lc->scale[l * 4 + i]);
                elevate3(cf, e);
                Simplex3 s;
                locate3(e, s);
                const float w = lc->window[l];
                const float a0 = xr[2 * l], a1 = xr[2 * l + 1];
                float* gtab = grad_lattice + (size_t)l * P.T * 2;
                const bool aggregate = lc->scale[l * 4] < kAggregateBelowScale;       // warp-uniform (l is)
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const unsigned idx = vindex3(s, r, P.cap_mask, (unsigned)P.T);
                    const float cb = s.bary[r] * w;
                    float2 cv = make_float2(cb * a0, cb * a1);
                    if (aggregate) {
                        const unsigned key = valid ? idx : 0xffffffffu;
                        const unsigned peers = __match_any_sync(kFull, key);
                        cv = add_peers2(peers, cv, lane);
                        if (valid && lane == __ffs(peers) - 1) red_v2(gtab + (size_t)idx * 2, cv);
                    } else if (valid) {
                        red_v2(gtab + (size_t)idx * 2, cv);
                    }
                }
            }
            if (valid) {
                const int tail = P.enc_cols;          // SH at [tail, tail+25), normal at tail+25.., geom at tail+28..
                if (g_geom) {
                    float* gg = g_geom + (size_t)n * kGeomDim + grp * 8;
#pragma unroll
                    for (int i = 0; i < 8; i++) gg[i] = xr[tail + kShCols + 3 + grp * 8 + i];
                }
                if (grp == 3 && g_sdf_grad) {
                    const float gx = sdf_grad[(size_t)n * 3], gy = sdf_grad[(size_t)n * 3 + 1], gz = sdf_grad[(size_t)n * 3 + 2];
                    const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
                    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
                    const float bx = xr[tail + kShCols], by = xr[tail + kShCols + 1], bz = xr[tail + kShCols + 2];
                    float ox = bx * inv, oy = by * inv, oz = bz * inv;
                    if (nrm > 1e-12f) {           // d (g / |g|) : (I - n n^T) / |g|
                        const float nx = gx * inv, ny = gy * inv, nz = gz * inv;
                        const float d = nx * bx + ny * by + nz * bz;
                        ox -= nx * d * inv; oy -= ny * d * inv; oz -= nz * d * inv;
                    }
                    float* o = g_sdf_grad + (size_t)n * 3;
                    if (P.acc_sdf_grad) { o[0] += ox; o[1] += oy; o[2] += oz; }
                    else { o[0] = ox; o[1] = oy; o[2] = oz; }
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();     // exchange tile / TMEM free for the next tile
    }
    __syncthreads();
    for (int i = tid; i < kNL * 128; i += kRgbThreads) {
        int l = i >> 7, c = i & 127;
        if (c < P.g.N[l] && s_gb[i] != 0.0f) atomicAdd(sp.gbias[l] + c, s_gb[i]);
    }
    if (tid == 0) umma::bulk_wait0();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ weight gradients
// stage = (tile, layer, half of the 128 samples): {zbar hi, zbar lo, a hi, a lo} halves, 16 KB each (the first / second 64
// sample rows of a tile are its first / second 16 KB)
constexpr int kDwStages = 3;
constexpr int kDwPiece = kWTileBytes / 2;
constexpr int kDwStageBytes = 4 * kDwPiece;
__global__ void __launch_bounds__(128, 1) k_rgb_dw(MlpGeom g, int ntiles, RgbSpill sp, float* __restrict__ gW0, float* __restrict__ gW1,
                                                  float* __restrict__ gW2, float* __restrict__ gW3) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kDwStages * kDwStageBytes);
    uint64_t* empty = full + kDwStages;
    uint64_t* done = empty + kDwStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kDwStages; i++) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
        umma::mbar_init(done, 1);
        umma::mbar_fence_init();
    }
    __syncthreads();
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    int my_tiles = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) my_tiles++;
    const int nstage = my_tiles * kNL * 2;            // (tile, layer, half)
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < nstage; i++) {
            const int slot = i % kDwStages, use = i / kDwStages;
            if (use > 0) umma::mbar_wait(&empty[slot], (use - 1) & 1);
            const int ti = blockIdx.x + (i / (kNL * 2)) * gridDim.x, l = (i >> 1) % kNL, h = i & 1;
            uint8_t* dst = ring + slot * kDwStageBytes;
            const uint8_t* zsrc = sp.zt[l] + (size_t)ti * kRgbSpillBytes + h * kDwPiece;
            const uint8_t* asrc = sp.at[l] + (size_t)ti * kRgbSpillBytes + h * kDwPiece;
            umma::mbar_expect_tx(&full[slot], (uint32_t)kDwStageBytes);
            umma::bulk_g2s(dst, zsrc, kDwPiece, &full[slot]);
            umma::bulk_g2s(dst + kDwPiece, zsrc + kWTileBytes, kDwPiece, &full[slot]);
            umma::bulk_g2s(dst + 2 * kDwPiece, asrc, kDwPiece, &full[slot]);
            umma::bulk_g2s(dst + 3 * kDwPiece, asrc + kWTileBytes, kDwPiece, &full[slot]);
        }
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < nstage; i++) {
            const int slot = i % kDwStages, use = i / kDwStages;
            umma::mbar_wait(&full[slot], use & 1);
            umma::fence_after_sync();
            const int l = (i >> 1) % kNL;
            const int M = g.Np[l] > 64 ? 128 : 64;
            const uint32_t idesc = umma::make_idesc_mn(M, g.Kp[l], umma::kFmtBF16);
            const uint32_t zh = umma::smem_u32(ring + slot * kDwStageBytes), zl = zh + kDwPiece, ah = zl + kDwPiece, al = ah + kDwPiece;
            const uint32_t d = tmem_base + l * 128;
            for (int kk = 0; kk < 4; kk++) {                         // 64 samples = 4 K-steps of 16
                const uint32_t ko = kk * 2 * kWSBO;
                const uint64_t dzh = umma::make_desc(zh + ko, kWSBO, kLBO), dzl = umma::make_desc(zl + ko, kWSBO, kLBO);
                const uint64_t dah = umma::make_desc(ah + ko, kWSBO, kLBO), dal = umma::make_desc(al + ko, kWSBO, kLBO);
                umma::mma_bf16(d, dzh, dah, idesc, (i >= kNL * 2 || (i & 1) || kk > 0) ? 1u : 0u);
                umma::mma_bf16(d, dzh, dal, idesc, 1u);
                umma::mma_bf16(d, dzl, dah, idesc, 1u);
            }
            umma::commit(&empty[slot]);
        }
        umma::commit(done);
    }
    __syncwarp();
    umma::mbar_wait(done, 0);
    umma::fence_after_sync();
    float* gW[kNL] = {gW0, gW1, gW2, gW3};
    if (nstage > 0) {
        for (int l = 0; l < kNL; l++) {
            const bool m128 = g.Np[l] > 64;
            // M = 128: accumulator row m in lane m; M = 64: row m in lane (m / 16) * 32 + m % 16
            const int m = m128 ? warp * 32 + lane : warp * 16 + lane;
            const bool has_row = (m128 || lane < 16) && m < g.N[l];
            for (int c = 0; c < g.Kp[l] / 16; c++) {
                float v[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + l * 128 + c * 16, v);
                umma::tmem_ld_wait();
                if (has_row) {
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int k = c * 16 + i;
                        if (k < g.K[l]) atomicAdd(gW[l] + (size_t)m * g.K[l] + k, v[i]);
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {

// chunked like the SDF backward (fused_sdf_bwd.cu): one wave of tiles per chunk so that the dW kernel reads the spill from L2
static int rgb_bwd_chunk_tiles(int ntiles) {
    static int chunk = -1;
    if (chunk < 0) {
        const char* e = getenv("PSDF_BWD_CHUNK_TILES");
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        chunk = e ? atoi(e) : 0;      // default: one chunk (measured on B200: per-wave chunks cost more in drain / refill than L2 hits give back)
        (void)sms;
    }
    return (chunk <= 0 || chunk > ntiles) ? ntiles : chunk;
}
long long psdf_rgb_fused_backward_workspace_bytes(int N) {
    const int ntiles = div_up(N > 0 ? N : 1, kTile);
    return (long long)2 * kNL * rgb_bwd_chunk_tiles(ntiles) * kRgbSpillBytes;
}

static int rgb_backward_impl(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                            const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                            int h1, int h2, int h3, const uint8_t* blob, const float* g_out, float* grad_lattice, float* g_sdf_grad,
                            float* g_geom, uint8_t* workspace, float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1,
                            float* gb2, float* gb3, int accumulate_sdf_grad, void* stream) {
    RgbParams P;
    int rc = make_rgb_params(P, N, L, T, geom_dim, points_scaling, h1, h2, h3);
    if (rc != PSDF_OK) return rc;
    P.acc_sdf_grad = accumulate_sdf_grad ? 1 : 0;
    if (N == 0) return PSDF_OK;
    if ((size_t)128 * (P.g.Kp[0] + 1) * 4 > (size_t)2 * kWTileBytes) return PSDF_ERR_UNSUPPORTED;
    RgbSpill sp;
    float* b[kNL] = {gb0, gb1, gb2, gb3};
    const int ntiles = div_up(N, kTile);
    const int chunk = rgb_bwd_chunk_tiles(ntiles);
    for (int l = 0; l < kNL; l++) {
        sp.zt[l] = workspace + (size_t)(2 * l) * chunk * kRgbSpillBytes;
        sp.at[l] = workspace + (size_t)(2 * l + 1) * chunk * kRgbSpillBytes;
        sp.gbias[l] = b[l];
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = (size_t)6 * kWTileBytes + sizeof(LevelC) + 2 * kNL * 128 * sizeof(float) + 64;
    // per device: set on every call (a second GPU needs its own opt-in)
    { static bool optin_[64]; psdf::psdf_optin_smem(k_rgb_fused_backward, 227 * 1024, optin_); }
    { static bool optin_[64]; psdf::psdf_optin_smem(k_rgb_dw, 227 * 1024, optin_); }
    const size_t smem_dw = (size_t)kDwStages * kDwStageBytes + 128;
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        const int nt = min(chunk, ntiles - t0);
        const size_t r0 = (size_t)t0 * kTile;
        RgbParams Pc = P;
        Pc.N = (int)min((size_t)nt * kTile, (size_t)N - r0);
        k_rgb_fused_backward<<<min(nt, sms), kRgbThreads, smem, ST>>>(Pc, pos + r0 * 3, dirs + r0 * 3, sdf_grad + r0 * 3, geom + r0 * kGeomDim,
                                                                      reinterpret_cast<const float2*>(lattice), scale_factor, shift, window,
                                                                      blob, g_out + r0 * 3, grad_lattice, g_sdf_grad ? g_sdf_grad + r0 * 3 : nullptr,
                                                                      g_geom ? g_geom + r0 * kGeomDim : nullptr, sp);
        PSDF_CHECK_LAUNCH();
        k_rgb_dw<<<min(nt, sms), 128, smem_dw, ST>>>(P.g, nt, sp, gW0, gW1, gW2, gW3);
        PSDF_CHECK_LAUNCH();
    }
    return PSDF_OK;
}

int psdf_rgb_fused_backward(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                            const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                            int h1, int h2, int h3, const uint8_t* blob, const float* g_out, float* grad_lattice, float* g_sdf_grad,
                            float* g_geom, uint8_t* workspace, float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1,
                            float* gb2, float* gb3, void* stream) {
    return rgb_backward_impl(N, L, T, pos, dirs, sdf_grad, geom, geom_dim, lattice, scale_factor, shift, window, points_scaling, h1, h2, h3, blob,
                             g_out, grad_lattice, g_sdf_grad, g_geom, workspace, gW0, gW1, gW2, gW3, gb0, gb1, gb2, gb3, 0, stream);
}
// same, with g_sdf_grad accumulated (+=): the buffer already holds d loss / d sdf_grad of the compositing and curvature terms
int psdf_rgb_fused_backward_acc(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                                const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                                int h1, int h2, int h3, const uint8_t* blob, const float* g_out, float* grad_lattice, float* g_sdf_grad,
                                float* g_geom, uint8_t* workspace, float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1,
                                float* gb2, float* gb3, void* stream) {
    return rgb_backward_impl(N, L, T, pos, dirs, sdf_grad, geom, geom_dim, lattice, scale_factor, shift, window, points_scaling, h1, h2, h3, blob,
                             g_out, grad_lattice, g_sdf_grad, g_geom, workspace, gW0, gW1, gW2, gW3, gb0, gb1, gb2, gb3, 1, stream);
}

}  // extern "C"
