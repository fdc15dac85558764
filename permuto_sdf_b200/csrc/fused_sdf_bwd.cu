// Training backward of the fused encoding + SDF MLP (sm_100a, tcgen05): given d loss / d {sdf, d sdf/dx, geom}
// per sample it produces the lattice gradient (scatter-add fused in the kernel), the bias gradients, and -- through the
// spilled per-sample layer adjoints / activations -- the weight gradients (second kernel, tensor cores).
//
// What it replaces in the reference: `loss.backward()` through SDF.get_sdf_and_gradient
// (permuto_sdf_py/models/models.py:199-259, create_graph=True) -- i.e. the double backward of the
// encoding (positions-gradient -> lattice) and of the 4-layer GELU MLP, ~60 PyTorch kernels per call.
//
// Math. The forward is y = MLP(enc(x)) (sdf = y0, geom = y1..), g = d sdf/dx. For fixed upstream gradients
// (ybar, gbar) the loss depends on the parameters through y and through the scalar s = gbar . g = D_v sdf, the
// directional derivative of sdf along v = gbar. So ONE tangent stream along v is enough (the forward kernel needs
// three because it must output g itself). With a_0 = enc(x), ta_0 = D_v enc(x) and for l = 1..4
//     z_l = W_l a_{l-1} + b_l,  a_l = gelu(z_l);     tz_l = W_l ta_{l-1},  ta_l = gelu'(z_l) tz_l,
// reverse mode gives (zbar_4 = ybar, tzbar_4 = e_0):
//     abar_{l-1} = W_l^T zbar_l,   tabar_{l-1} = W_l^T tzbar_l
//     tzbar_{l-1} = gelu'(z_{l-1}) tabar_{l-1}
//     zbar_{l-1}  = gelu'(z_{l-1}) abar_{l-1} + gelu''(z_{l-1}) tz_{l-1} tabar_{l-1}
//     dW_l = zbar_l a_{l-1}^T + tzbar_l ta_{l-1}^T,   db_l = zbar_l
//     lattice[l][idx_r] += window_l (B_r abar_0[l] + dB_r tabar_0[l])       (B barycentric weights, dB their tangent)
//
// Kernel structure per 128-sample tile (512 threads: row = tid & 127, group g = tid >> 7 owns operand cores g, g+4, ..
// in the encoder phases and the 16-column chunk g in the epilogues):
//   1. encoder (all four groups): a_0, ta_0 -> bf16 hi/lo operand tiles in smem (later spilled as they are, by TMA store);
//   2. forward recompute, layers 1..3, 2 streams, tcgen05 (weights from one TMA bulk copy); the pre-activations
//      z_l, tz_l STAY in TMEM (6 x 64 columns) for the reverse sweep; a_l, ta_l go to operand tiles + spill;
//   3. reverse sweep, layers 4..1: zbar/tzbar tiles -> tcgen05 with the transposed weights (second bulk copy)
//      -> abar/tabar -> elementwise with gelu', gelu'' from the TMEM-resident z;
//   4. encoder backward (both groups): warp-aggregated red.global.add.v2.f32 into the lattice gradient.
// dW is formed by the second kernel of this file (k_sdf_dw) from the operand tiles that phases 1-3 spill by TMA store:
// the sample axis becomes the MMA K dimension and the tiles are consumed as MN-major operands, exactly as stored.
#include "fused_common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf_fused;

namespace {
constexpr int kBwdThreads = 512;
constexpr int kGroups = kBwdThreads / kTile;

constexpr int kSpillTileBytes = 4 * kATileBytes;     // one 128-sample operand tile set: [value hi | value lo | tangent hi | tangent lo]
struct Spill {
    uint8_t* zt[kNL];   // [ntiles][64 KB]  zbar_l / tzbar_l operand tiles exactly as they sit in shared memory
    uint8_t* at[kNL];   // [ntiles][64 KB]  a_{l-1} / ta_{l-1} operand tiles
    float* gbias[kNL];  // [Np_l]  (+=)
};

__global__ void __launch_bounds__(kBwdThreads, 1)
k_sdf_fused_backward(FusedParams P, const float* __restrict__ pos, const float2* __restrict__ lattice, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ window, const uint8_t* __restrict__ blob,
                     const float* __restrict__ g_sdf, const float* __restrict__ g_grad, const float* __restrict__ g_geom,
                     float* __restrict__ grad_lattice, Spill sp) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int wbytes = P.g.total > P.g.total_t ? P.g.total : P.g.total_t;
    uint8_t* s_w = smem;                                   // W blob (forward) then W^T blob (reverse)
    uint8_t* s_a0 = smem + wbytes;                         // layer-0 operand tiles: [value hi, value lo, tangent hi, tangent lo]
    uint8_t* s_t = s_a0 + 4 * kATileBytes;                 // working operand tiles, same order
    LevelC* lc = reinterpret_cast<LevelC*>(s_t + 4 * kATileBytes);
    float* s_bias = reinterpret_cast<float*>(lc + 1);      // 4 x 64 biases (survive the W -> W^T swap)
    float* s_gb = s_bias + kNL * 64;                       // 4 x 64 bias-gradient accumulators of this CTA
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_gb + kNL * 64);   // [0] weights, [1] mma
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    float* s_x = reinterpret_cast<float*>(s_a0);           // exchange tile abar_0 | tabar_0 : [128][2*Kp0+1] fp32 (aliases s_a0 + s_t)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = tid & 127, grp = tid >> 7;
    const int level_cores = P.L / 4;
    const int K0 = P.g.Kp[0];

    if (tid == 0) { umma::mbar_init(&bars[0], 1); umma::mbar_init(&bars[1], 1); umma::mbar_fence_init(); }
    for (int i = tid; i < P.L * 3; i += kBwdThreads) {
        lc->scale[(i / 3) * 4 + (i % 3)] = scale[i];
        lc->shift[(i / 3) * 4 + (i % 3)] = shift ? shift[i] : 0.0f;
    }
    for (int i = tid; i < P.L; i += kBwdThreads) lc->window[i] = window ? window[i] : 1.0f;
    for (int i = tid; i < kNL * 64; i += kBwdThreads) {
        int l = i >> 6, c = i & 63;
        s_bias[i] = (c < P.g.Np[l]) ? reinterpret_cast<const float*>(blob + P.g.bias[l])[c] : 0.0f;
        s_gb[i] = 0.0f;
    }
    __syncthreads();
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_work = tmem_base + 384;            // 2 x 64 working columns after the 6 x 64 stored ones
    uint32_t w_phase = 0, mma_phase = 0;

    const int ntiles = (P.N + kTile - 1) / kTile;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int n = tile * kTile + row;
        const bool valid = n < P.N;
        // forward weights for this tile (the buffer holds W^T from the previous tile's reverse sweep)
        if (tid == 0) {
            umma::mbar_expect_tx(&bars[0], (uint32_t)P.g.total);
            umma::bulk_g2s(s_w, blob, (uint32_t)P.g.total, &bars[0]);
        }
        float x[3], v[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            x[i] = valid ? pos[(size_t)n * 3 + i] : 0.0f;
            v[i] = (valid && g_grad) ? g_grad[(size_t)n * 3 + i] : 0.0f;
        }
        // ---------------- 1. encoder: operand cores grp, grp+4, ... (4 levels = 8 features = one 16-byte core row)
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
            store8(s_a0, s_a0 + kATileBytes, row, kc, fv);
            store8(s_a0 + 2 * kATileBytes, s_a0 + 3 * kATileBytes, row, kc, ft);
        }
        umma::mbar_wait(&bars[0], w_phase);
        w_phase ^= 1;

        // ---------------- 2. forward recompute, layers 1..3 (index l = 0..2); z_l, tz_l stay in TMEM columns (2l+s)*64
#pragma unroll 1
        for (int l = 0; l < 3; l++) {
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            const uint8_t* at = (l == 0) ? s_a0 : s_t;
            if (tid == 0) {
                umma::fence_after_sync();
                // the input tiles of layer l+1 (a_l, ta_l) go to the weight-gradient spill as they are (TMA store)
                umma::bulk_s2g(sp.at[l] + (size_t)tile * kSpillTileBytes, at, kSpillTileBytes);
                umma::bulk_commit();
                for (int s = 0; s < 2; s++)
                    issue_gemm(tmem_base + (2 * l + s) * 64, at + s * 2 * kATileBytes, at + s * 2 * kATileBytes + kATileBytes, s_w + P.g.w_hi[l],
                               s_w + P.g.w_lo[l], P.g.Kp[l], P.g.Np[l]);
                umma::bulk_wait_read0();      // before the commit: whoever sees the MMAs done may overwrite the tiles
                umma::commit(&bars[1]);
            }
            umma::mbar_wait(&bars[1], mma_phase);
            umma::fence_after_sync();
            {
                const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2 * l * 64;
                const int c = grp;
                if (c < P.g.Np[l] / 16) {
                    float z[16], tz[16];
                    umma::tmem_ld16(trow + c * 16, z);
                    umma::tmem_ld16(trow + 64 + c * 16, tz);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float zz = z[i] + s_bias[l * 64 + c * 16 + i];
                        const GeluEval ge = gelu_eval(zz);
                        z[i] = zz * ge.cdf;
                        tz[i] *= fmaf(zz, ge.pdf, ge.cdf);
                    }
                    store8(s_t, s_t + kATileBytes, row, 2 * c, z);
                    store8(s_t, s_t + kATileBytes, row, 2 * c + 1, z + 8);
                    store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c, tz);
                    store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c + 1, tz + 8);
                }
            }
            mma_phase ^= 1;
        }
        // all forward MMAs are complete (last commit was waited for): swap in the transposed weights
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            umma::mbar_expect_tx(&bars[0], (uint32_t)P.g.total_t);
            umma::bulk_g2s(s_w, blob + P.g.total, (uint32_t)P.g.total_t, &bars[0]);
            // a_3 / ta_3 (input of the last layer) are only needed for dW_3: spill them before the seed overwrites the tiles
            umma::bulk_s2g(sp.at[3] + (size_t)tile * kSpillTileBytes, s_t, kSpillTileBytes);
            umma::bulk_commit();
            umma::bulk_wait_read0();
        }
        __syncthreads();

        // ---------------- 3. reverse sweep, layers 4..1 (index l = 3..0)
        // seed: zbar_4 = [g_sdf, g_geom...], tzbar_4 = e_0 ; written straight into the working tiles + spill
        {
            const int Np = P.g.Np[3], nout = P.g.N[3];
            for (int c = grp; c < Np / 8; c += kGroups) {
                float zb[8], tb[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    int col = c * 8 + i;
                    float vz = 0.f;
                    if (valid) {
                        if (col == 0) vz = g_sdf ? g_sdf[n] : 0.f;
                        else if (col < nout) vz = g_geom ? g_geom[(size_t)n * (nout - 1) + col - 1] : 0.f;
                    }
                    zb[i] = vz;
                    tb[i] = (valid && col == 0) ? 1.0f : 0.0f;
                }
                store8(s_t, s_t + kATileBytes, row, c, zb);
                store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, c, tb);
                {                                                  // bias gradient: column sums over the warp's 32 rows
                    float z16[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) { z16[i] = zb[i]; z16[i + 8] = 0.f; }
                    int col;
                    const float cs = colsum16(z16, lane, col);
                    if (!(lane & 1) && col < 8) atomicAdd(&s_gb[3 * 64 + c * 8 + col], cs);
                }
            }
        }
        umma::mbar_wait(&bars[0], w_phase);
        w_phase ^= 1;
#pragma unroll 1
        for (int l = 3; l >= 0; l--) {
            // abar_{l-1} = zbar_l W_l : A = working tiles [128 x Np_l], B = W_l^T [Kp_l rows x Np_l], result Kp_l columns
            umma::fence_async_smem();
            umma::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
                umma::bulk_s2g(sp.zt[l] + (size_t)tile * kSpillTileBytes, s_t, kSpillTileBytes);     // zbar_l, tzbar_l tiles
                umma::bulk_commit();
                for (int s = 0; s < 2; s++)
                    issue_gemm(tmem_work + s * 64, s_t + s * 2 * kATileBytes, s_t + s * 2 * kATileBytes + kATileBytes, s_w + P.g.t_hi[l],
                               s_w + P.g.t_lo[l], P.g.Np[l], P.g.Kp[l]);
                umma::bulk_wait_read0();
                umma::commit(&bars[1]);
            }
            // while the MMAs run: the GELU' / GELU'' factors of layer l-1 depend only on pre-activations that are already in TMEM
            const int cr = grp;
            const bool have_chunk = cr < P.g.Kp[l] / 16;
            float g1[16], g2t[16];
            if (l > 0 && have_chunk) {
                float z[16], tz[16];
                const uint32_t tst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2 * (l - 1) * 64;
                umma::tmem_ld16(tst + cr * 16, z);
                umma::tmem_ld16(tst + 64 + cr * 16, tz);
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float zz = z[i] + s_bias[(l - 1) * 64 + cr * 16 + i];
                    const GeluEval ge = gelu_eval(zz);
                    g1[i] = fmaf(zz, ge.pdf, ge.cdf);
                    g2t[i] = ge.pdf * (2.0f - zz * zz) * tz[i];
                }
            }
            umma::mbar_wait(&bars[1], mma_phase);
            umma::fence_after_sync();
            {
                const uint32_t twork = tmem_work + ((uint32_t)((warp & 3) * 32) << 16);
                const int Kp = P.g.Kp[l];
                const int c = grp;
                if (c < Kp / 16) {
                    float ab[16], tab_[16];
                    umma::tmem_ld16(twork + c * 16, ab);
                    umma::tmem_ld16(twork + 64 + c * 16, tab_);
                    if (l > 0) {
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const float zb = g1[i] * ab[i] + g2t[i] * tab_[i];
                            tab_[i] = g1[i] * tab_[i];
                            ab[i] = zb;
                        }
                        // ab = zbar_{l-1}, tab_ = tzbar_{l-1}
                        store8(s_t, s_t + kATileBytes, row, 2 * c, ab);
                        store8(s_t, s_t + kATileBytes, row, 2 * c + 1, ab + 8);
                        store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c, tab_);
                        store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c + 1, tab_ + 8);
                        {                                              // bias gradient: column sums over the warp's 32 rows
                            if (!valid) {
#pragma unroll
                                for (int i = 0; i < 16; i++) ab[i] = 0.f;
                            }
                            int col;
                            const float cs = colsum16(ab, lane, col);
                            if (!(lane & 1)) atomicAdd(&s_gb[(l - 1) * 64 + c * 16 + col], cs);
                        }
                    } else {
                        umma::tmem_ld_wait();
                        // abar_0 | tabar_0 for the encoder backward (fp32 exchange tile, aliases the working tiles, which
                        // the just-completed MMA no longer reads)
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            s_x[row * (2 * K0 + 1) + c * 16 + i] = ab[i];
                            s_x[row * (2 * K0 + 1) + K0 + c * 16 + i] = tab_[i];
                        }
                    }
                }
            }
            mma_phase ^= 1;
        }
        umma::fence_before_sync();
        __syncthreads();

        // ---------------- 4. encoder backward: lattice[l][idx_r] += window_l (B_r abar_0[l] + dB_r tabar_0[l])
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
    __syncthreads();
    // bias gradients of this CTA
    for (int i = tid; i < kNL * 64; i += kBwdThreads) {
        int l = i >> 6, c = i & 63;
        if (c < P.g.N[l] && s_gb[i] != 0.0f) atomicAdd(sp.gbias[l] + c, s_gb[i]);
    }
    if (tid == 0) umma::bulk_wait0();      // spill stores complete before the CTA retires
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ weight gradients
// dW_l [N_l, K_l] = sum over samples of  zbar_l^T a_{l-1} + tzbar_l^T ta_{l-1}  from the spilled operand tiles.
// The sample axis is the MMA K dimension: a [128 samples x 64 columns] tile in the K-major core-matrix layout used for the
// forward GEMMs IS an MN-major operand for this product (core matrix = 8 samples x 8 columns; LBO = the 8-sample-group stride,
// SBO = the 8-column-core stride), so the tiles are multiplied exactly as they were stored: no transposition, no conversion.
// Persistent CTAs; per (tile, layer, stream) one 64 KB stage = {zbar hi, lo, a hi, lo} arrives by TMA into a 3-deep ring,
// one thread issues 3 split products x 8 K-steps of tcgen05.mma M64 x N(Kp_l) x K16 accumulating in TMEM (4 layers x 64
// columns; accumulator row m lives in lane (m / 16) * 32 + m % 16), the epilogue adds the CTA's partial sums to dW (red.add).
constexpr int kDwStages = 3;
constexpr int kDwStageBytes = 4 * kATileBytes;
__global__ void __launch_bounds__(128, 1) k_sdf_dw(MlpGeom g, int ntiles, Spill sp, float* __restrict__ gW0, float* __restrict__ gW1,
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
    if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    int my_tiles = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) my_tiles++;
    const int nstage = my_tiles * kNL * 2;            // (tile, layer, stream)
    if (warp == 0 && lane == 0) {
        // ---- producer: TMA loads
        for (int i = 0; i < nstage; i++) {
            const int slot = i % kDwStages, use = i / kDwStages;
            if (use > 0) umma::mbar_wait(&empty[slot], (use - 1) & 1);
            const int ti = blockIdx.x + (i / (kNL * 2)) * gridDim.x, l = (i >> 1) % kNL, st = i & 1;
            uint8_t* dst = ring + slot * kDwStageBytes;
            umma::mbar_expect_tx(&full[slot], (uint32_t)kDwStageBytes);
            umma::bulk_g2s(dst, sp.zt[l] + (size_t)ti * kSpillTileBytes + st * 2 * kATileBytes, 2 * kATileBytes, &full[slot]);
            umma::bulk_g2s(dst + 2 * kATileBytes, sp.at[l] + (size_t)ti * kSpillTileBytes + st * 2 * kATileBytes, 2 * kATileBytes, &full[slot]);
        }
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        for (int i = 0; i < nstage; i++) {
            const int slot = i % kDwStages, use = i / kDwStages;
            umma::mbar_wait(&full[slot], use & 1);
            umma::fence_after_sync();
            const int l = (i >> 1) % kNL;
            const uint32_t idesc = umma::make_idesc_mn(64, g.Kp[l], umma::kFmtBF16);
            const uint32_t zh = umma::smem_u32(ring + slot * kDwStageBytes), zl = zh + kATileBytes, ah = zl + kATileBytes, al = ah + kATileBytes;
            const uint32_t d = tmem_base + l * 64;
            for (int kk = 0; kk < kTile / 16; kk++) {
                const uint32_t ko = kk * 2 * kSBO_A;                 // 16 samples = two 8-sample groups
                const uint64_t dzh = umma::make_desc(zh + ko, kSBO_A, kLBO), dzl = umma::make_desc(zl + ko, kSBO_A, kLBO);
                const uint64_t dah = umma::make_desc(ah + ko, kSBO_A, kLBO), dal = umma::make_desc(al + ko, kSBO_A, kLBO);
                umma::mma_bf16(d, dzh, dah, idesc, (i >= kNL * 2 || (i & 1) || kk > 0) ? 1u : 0u);   // first stage of a layer clears
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
    // ---- epilogue: lanes 0..15 of warp w hold accumulator rows 16 w .. 16 w + 15
    float* gW[kNL] = {gW0, gW1, gW2, gW3};
    if (nstage > 0) {
        for (int l = 0; l < kNL; l++) {
            const int m = warp * 16 + lane;
            for (int c = 0; c < g.Kp[l] / 16; c++) {
                float v[16];
                umma::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + l * 64 + c * 16, v);
                umma::tmem_ld_wait();
                if (lane < 16 && m < g.N[l]) {
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
    if (warp == 0) umma::tmem_dealloc(tmem_base, 256);
}

#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {

// The two kernels can run chunk by chunk (PSDF_BWD_CHUNK_TILES tiles per chunk, same spill buffer every time) so that a chunk's
// spill is consumed by the dW kernel while it is still in the 126 MB L2. Measured on B200 at N = 65 536: chunks of one wave (148
// tiles) make the iteration 24 % SLOWER (2.45 vs 1.98 ms) -- the kernels drain and refill the GPU at every chunk boundary and the
// persistent CTAs lose their amortisation -- so the default is one chunk; keeping the spill in L2 needs the dW product inside the
// backward kernel itself (round 2).
static int bwd_chunk_tiles(int ntiles) {
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
// workspace: 8 tile spills (zbar_l, a_{l-1} for l = 0..3) of one chunk, 64 KB per tile each, written and consumed inside this call
long long psdf_sdf_fused_backward_workspace_bytes(int N) {
    const int ntiles = div_up(N > 0 ? N : 1, kTile);
    return (long long)2 * kNL * bwd_chunk_tiles(ntiles) * kSpillTileBytes;
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
    Spill sp;
    float* b[kNL] = {gb0, gb1, gb2, gb3};
    const int ntiles = div_up(N, kTile);
    const int chunk = bwd_chunk_tiles(ntiles);
    for (int l = 0; l < kNL; l++) {
        sp.zt[l] = workspace + (size_t)(2 * l) * chunk * kSpillTileBytes;
        sp.at[l] = workspace + (size_t)(2 * l + 1) * chunk * kSpillTileBytes;
        sp.gbias[l] = b[l];
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int wbytes = P.g.total > P.g.total_t ? P.g.total : P.g.total_t;
    size_t smem = (size_t)wbytes + 8 * kATileBytes + sizeof(LevelC) + 2 * kNL * 64 * sizeof(float) + 64;
    if ((size_t)128 * (2 * P.g.Kp[0] + 1) * 4 > (size_t)8 * kATileBytes) return PSDF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(k_sdf_fused_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done = true; }
    const size_t smem_dw = (size_t)kDwStages * kDwStageBytes + 128;
    static bool attr_dw = false;
    if (!attr_dw) { cudaFuncSetAttribute(k_sdf_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_dw = true; }
    const int geom_cols = out_dim - 1;
    for (int t0 = 0; t0 < ntiles; t0 += chunk) {
        const int nt = min(chunk, ntiles - t0);
        const size_t r0 = (size_t)t0 * kTile;
        FusedParams Pc = P;
        Pc.N = (int)min((size_t)nt * kTile, (size_t)N - r0);
        k_sdf_fused_backward<<<min(nt, sms), kBwdThreads, smem, ST>>>(Pc, pos + r0 * 3, reinterpret_cast<const float2*>(lattice), scale_factor,
                                                                      shift, window, blob, g_sdf ? g_sdf + r0 : nullptr,
                                                                      g_grad ? g_grad + r0 * 3 : nullptr,
                                                                      g_geom ? g_geom + r0 * geom_cols : nullptr, grad_lattice, sp);
        PSDF_CHECK_LAUNCH();
        k_sdf_dw<<<min(nt, sms), 128, smem_dw, ST>>>(P.g, nt, sp, gW0, gW1, gW2, gW3);
        PSDF_CHECK_LAUNCH();
    }
    return PSDF_OK;
}

}  // extern "C"
