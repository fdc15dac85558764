// Training backward of the fused encoding + SDF MLP (sm_100a, tcgen05): given d loss / d {sdf, d sdf/dx, geom}
// per sample it produces the lattice gradient (scatter-add fused in the kernel), the bias gradients, and the
// per-sample layer adjoints / activations whose products are the weight gradients.
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
//   1. encoder (both groups): a_0, ta_0 -> bf16 hi/lo operand tiles in smem, and to the dW spill buffer;
//   2. forward recompute, layers 1..3, 2 streams, tcgen05 (weights from one TMA bulk copy); the pre-activations
//      z_l, tz_l STAY in TMEM (6 x 64 columns) for the reverse sweep; a_l, ta_l go to operand tiles + spill;
//   3. reverse sweep, layers 4..1: zbar/tzbar tiles -> tcgen05 with the transposed weights (second bulk copy)
//      -> abar/tabar -> elementwise with gelu', gelu'' from the TMEM-resident z;
//   4. encoder backward (both groups): warp-aggregated red.global.add.v2.f32 into the lattice gradient.
// dW itself is formed outside from the spilled [2N, 64] matrices with 4 plain library GEMMs (cuBLAS through
// torch.matmul); folding them into this kernel (M=64 MMAs on transposed tiles) is the next step.
#include "fused_common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf_fused;

namespace {
constexpr unsigned kFull = 0xffffffffu;
constexpr int kBwdThreads = 512;
constexpr int kGroups = kBwdThreads / kTile;

struct Spill {
    float* zcat[kNL];   // [2N, Np_l]   rows [0,N): zbar_l, rows [N,2N): tzbar_l
    float* acat[kNL];   // [2N, Kp_l]   rows [0,N): a_{l-1}, rows [N,2N): ta_{l-1}
    float* gbias[kNL];  // [Np_l]  (+=)
};

__device__ __forceinline__ float2 add_peers2(unsigned peers, float2 x, int lane) {
    int rel = __popc(peers << (31 - lane) << 1);
    peers &= (0xfffffffeu << lane);
    while (__any_sync(kFull, peers)) {
        int next = __ffs(peers);
        float tx = __shfl_sync(kFull, x.x, (next - 1) & 31);
        float ty = __shfl_sync(kFull, x.y, (next - 1) & 31);
        if (next) { x.x += tx; x.y += ty; }
        int done = rel & 1;
        peers &= __ballot_sync(kFull, !done);
        rel >>= 1;
    }
    return x;
}
__device__ __forceinline__ void red_v2(float* addr, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// column sums of 16 per-lane values over the 32 lanes of a warp with 16 shuffles (halving butterfly): afterwards both
// lanes of the pair {2p, 2p+1} hold the sum of column `col` = bit-reversal-free index built from lane bits 4..1.
__device__ __forceinline__ float colsum16(const float* v, int lane, int& col) {
    float a[8], b[4], c[2], d;
    bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; i++) { float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i]; a[i] = keep + __shfl_xor_sync(kFull, send, 16); }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; i++) { float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i]; b[i] = keep + __shfl_xor_sync(kFull, send, 8); }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; i++) { float send = up ? b[i] : b[i + 2], keep = up ? b[i + 2] : b[i]; c[i] = keep + __shfl_xor_sync(kFull, send, 4); }
    up = lane & 2;
    { float send = up ? c[0] : c[1], keep = up ? c[1] : c[0]; d = keep + __shfl_xor_sync(kFull, send, 2); }
    d += __shfl_xor_sync(kFull, d, 1);
    col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    return d;
}
// 8 consecutive fp32 of a row to global (two float4 stores)
__device__ __forceinline__ void st8(float* dst, const float* v) {
    reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

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
    const size_t Nn = (size_t)P.N;

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
            if (valid) {
                st8(sp.acat[0] + (size_t)n * K0 + kc * 8, fv);
                st8(sp.acat[0] + (Nn + n) * K0 + kc * 8, ft);
            }
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
                for (int s = 0; s < 2; s++)
                    issue_gemm(tmem_base + (2 * l + s) * 64, at + s * 2 * kATileBytes, at + s * 2 * kATileBytes + kATileBytes, s_w + P.g.w_hi[l],
                               s_w + P.g.w_lo[l], P.g.Kp[l], P.g.Np[l]);
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
                    if (valid) {
                        const int Kn = P.g.Kp[l + 1];
                        st8(sp.acat[l + 1] + (size_t)n * Kn + c * 16, z); st8(sp.acat[l + 1] + (size_t)n * Kn + c * 16 + 8, z + 8);
                        st8(sp.acat[l + 1] + (Nn + n) * Kn + c * 16, tz); st8(sp.acat[l + 1] + (Nn + n) * Kn + c * 16 + 8, tz + 8);
                    }
                }
            }
            mma_phase ^= 1;
        }
        // all forward MMAs are complete (last commit was waited for): swap in the transposed weights
        umma::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            umma::mbar_expect_tx(&bars[0], (uint32_t)P.g.total_t);
            umma::bulk_g2s(s_w, blob + P.g.total, (uint32_t)P.g.total_t, &bars[0]);
        }

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
                if (valid) { st8(sp.zcat[3] + (size_t)n * Np + c * 8, zb); st8(sp.zcat[3] + (Nn + n) * Np + c * 8, tb); }
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
                for (int s = 0; s < 2; s++)
                    issue_gemm(tmem_work + s * 64, s_t + s * 2 * kATileBytes, s_t + s * 2 * kATileBytes + kATileBytes, s_w + P.g.t_hi[l],
                               s_w + P.g.t_lo[l], P.g.Np[l], P.g.Kp[l]);
                umma::commit(&bars[1]);
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
                        float z[16], tz[16];
                        const uint32_t tst = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 2 * (l - 1) * 64;
                        umma::tmem_ld16(tst + c * 16, z);
                        umma::tmem_ld16(tst + 64 + c * 16, tz);
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            float zz = z[i] + s_bias[(l - 1) * 64 + c * 16 + i];
                            const GeluEval ge = gelu_eval(zz);
                            const float g1 = fmaf(zz, ge.pdf, ge.cdf), g2 = ge.pdf * (2.0f - zz * zz);
                            float zb = g1 * ab[i] + g2 * tz[i] * tab_[i];
                            tab_[i] = g1 * tab_[i];
                            ab[i] = zb;
                        }
                        // ab = zbar_{l-1}, tab_ = tzbar_{l-1}
                        store8(s_t, s_t + kATileBytes, row, 2 * c, ab);
                        store8(s_t, s_t + kATileBytes, row, 2 * c + 1, ab + 8);
                        store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c, tab_);
                        store8(s_t + 2 * kATileBytes, s_t + 3 * kATileBytes, row, 2 * c + 1, tab_ + 8);
                        if (valid) {
                            st8(sp.zcat[l - 1] + (size_t)n * Kp + c * 16, ab); st8(sp.zcat[l - 1] + (size_t)n * Kp + c * 16 + 8, ab + 8);
                            st8(sp.zcat[l - 1] + (Nn + n) * Kp + c * 16, tab_); st8(sp.zcat[l - 1] + (Nn + n) * Kp + c * 16 + 8, tab_ + 8);
                        }
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
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    unsigned idx = vindex3(s, r, P.cap_mask, (unsigned)P.T);
                    float cb = s.bary[r] * w, cd = db[r] * w;
                    float2 cv = make_float2(cb * a0 + cd * t0, cb * a1 + cd * t1);
                    unsigned key = valid ? idx : 0xffffffffu;
                    unsigned peers = __match_any_sync(kFull, key);
                    cv = add_peers2(peers, cv, lane);
                    if (valid && lane == __ffs(peers) - 1) red_v2(gtab + (size_t)idx * 2, cv);
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
    if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {

// Spill buffers (caller allocated, fp32): for l = 0..3  zcat_l [2N, Np_l], acat_l [2N, Kp_l]; Np/Kp = dims padded to 16.
// grad_lattice and grad_bias_l are accumulated (+=). Weight gradients: dW_l = (zcat_l^T @ acat_l)[:N_l, :K_l].
int psdf_sdf_fused_backward(int N, int L, int T, const float* pos, const float* lattice, const float* scale_factor, const float* shift,
                            const float* window, float points_scaling, int hidden, int out_dim, const uint8_t* blob, const float* g_sdf,
                            const float* g_grad, const float* g_geom, float* grad_lattice, float* zcat0, float* zcat1, float* zcat2,
                            float* zcat3, float* acat0, float* acat1, float* acat2, float* acat3, float* gb0, float* gb1, float* gb2,
                            float* gb3, void* stream) {
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
    float* z[kNL] = {zcat0, zcat1, zcat2, zcat3};
    float* a[kNL] = {acat0, acat1, acat2, acat3};
    float* b[kNL] = {gb0, gb1, gb2, gb3};
    for (int l = 0; l < kNL; l++) { sp.zcat[l] = z[l]; sp.acat[l] = a[l]; sp.gbias[l] = b[l]; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int wbytes = P.g.total > P.g.total_t ? P.g.total : P.g.total_t;
    size_t smem = (size_t)wbytes + 8 * kATileBytes + sizeof(LevelC) + 2 * kNL * 64 * sizeof(float) + 64;
    if ((size_t)128 * (2 * P.g.Kp[0] + 1) * 4 > (size_t)8 * kATileBytes) return PSDF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(k_sdf_fused_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); attr_done = true; }
    const int ntiles = div_up(N, kTile);
    k_sdf_fused_backward<<<min(ntiles, sms), kBwdThreads, smem, ST>>>(P, pos, reinterpret_cast<const float2*>(lattice), scale_factor, shift,
                                                                      window, blob, g_sdf, g_grad, g_geom, grad_lattice, sp);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}

}  // extern "C"
