// Shared pieces of the fused colour-network kernels (fused_rgb.cu forward, fused_rgb_bwd.cu backward).
// Operand tiles here are [128 samples x up to 128 columns] bf16 in the K-major core-matrix layout:
//     offset(row, col) = (row / 8) * 2048 + (col / 8) * 128 + (row % 8) * 16 + (col % 8) * 2
// Input column order of layer 0 (= torch.cat in RGB.forward, models.py:395-404), with E = 2L + 4 encoder columns:
//     [0, 2L)          lattice features, level l at columns 2l, 2l+1
//     [2L, 2L+3), 2L+3 concat points (x * scaling), zero pad
//     [E, E+25)        real spherical harmonics of degree 5 of the view direction
//     [E+25, E+28)     normalised sdf gradient
//     [E+28, E+28+G)   geometric feature (G = 32)
#pragma once
#include "fused_common.cuh"
#include "sh.cuh"

namespace psdf_rgb {
using namespace psdf_fused;

constexpr int kRgbThreads = 512;
constexpr int kRgbGroups = kRgbThreads / kTile;      // 4 threads per sample
constexpr int kWSBO = 2048;                          // 8-row-group stride of a 128-column tile
constexpr int kWTileBytes = 128 * 128 * 2;           // one bf16 tile (hi or lo)
constexpr int kGeomDim = 32;
constexpr int kShCols = 25;

struct RgbParams {
    int N, L, T;
    unsigned cap_mask;
    float points_scaling;
    int enc_cols;        // 2L + 4
    int in_dim;          // enc_cols + 25 + 3 + 32
    int acc_sdf_grad = 0; // backward: g_sdf_grad += instead of = (the caller's buffer already holds the compositing / curvature terms)
    MlpGeom g;
};

inline int make_rgb_params(RgbParams& P, int N, int L, int T, int geom_dim, float points_scaling, int h1, int h2, int h3) {
    if (N < 0 || L < 4 || L > kMaxLevels || (L % 4) != 0 || geom_dim != kGeomDim) return -3;
    P.N = N; P.L = L; P.T = T;
    P.cap_mask = t_magic(T);
    P.points_scaling = points_scaling;
    P.enc_cols = 2 * L + 4;
    P.in_dim = P.enc_cols + kShCols + 3 + kGeomDim;
    if (P.in_dim > 128 || h1 > 128 || h2 > 128 || h3 > 128 || (h1 % 16) || (h2 % 16) || (h3 % 16)) return -3;
    int dims[kNL + 1] = {P.in_dim, h1, h2, h3, 3};
    P.g = make_geom_dims(dims);
    return 0;
}

__device__ __forceinline__ void load_level_consts(LevelC* lc, int L, const float* scale, const float* shift, const float* window, int tid,
                                                  int nthreads) {
    for (int i = tid; i < L * 3; i += nthreads) {
        lc->scale[(i / 3) * 4 + (i % 3)] = scale[i];
        lc->shift[(i / 3) * 4 + (i % 3)] = shift ? shift[i] : 0.0f;
    }
    for (int i = tid; i < L; i += nthreads) lc->window[i] = window ? window[i] : 1.0f;
}

// hi + lo block of one layer (adjacent in the blob) -> the weight buffer, one TMA bulk copy (single thread)
__device__ __forceinline__ void load_layer_weights(uint8_t* s_w, const uint8_t* blob, const MlpGeom& g, int l, bool transposed, uint64_t* bar) {
    const uint32_t bytes = (uint32_t)(g.Np[l] * g.Kp[l] * 4);
    const uint8_t* src = transposed ? blob + g.total + g.t_hi[l] : blob + g.w_hi[l];
    umma::mbar_expect_tx(bar, bytes);
    umma::bulk_g2s(s_w, src, bytes, bar);
}

__device__ __forceinline__ void store8w(uint8_t* a_hi, uint8_t* a_lo, int row, int kcore, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) umma::split2_bf16(v[2 * i], v[2 * i + 1], h[i], l[i]);
    const int off = (row >> 3) * kWSBO + kcore * kLBO + (row & 7) * 16;
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// three split products of [128 x Kp] x [Np x Kp]^T into TMEM (single thread); A tiles use the 2048-byte row-group stride
__device__ __forceinline__ void issue_gemm_w(uint32_t tmem_d, const uint8_t* a_hi, const uint8_t* a_lo, const uint8_t* w_hi, const uint8_t* w_lo,
                                             int Kp, int Np) {
    const uint32_t idesc = umma::make_idesc(128, Np, umma::kFmtBF16);
    const uint32_t sbo_w = (Kp / 8) * kLBO;
    const uint64_t dah0 = umma::make_desc(umma::smem_u32(a_hi), kLBO, kWSBO), dal0 = umma::make_desc(umma::smem_u32(a_lo), kLBO, kWSBO);
    const uint64_t dwh0 = umma::make_desc(umma::smem_u32(w_hi), kLBO, sbo_w), dwl0 = umma::make_desc(umma::smem_u32(w_lo), kLBO, sbo_w);
    for (int kk = 0; kk < Kp / 16; kk++) {
        const uint64_t off = (uint64_t)(kk * ((2 * kLBO) >> 4));      // start-address field, 16-byte units
        umma::mma_bf16(tmem_d, dah0 + off, dwh0 + off, idesc, kk > 0 ? 1u : 0u);
        umma::mma_bf16(tmem_d, dah0 + off, dwl0 + off, idesc, 1u);
        umma::mma_bf16(tmem_d, dal0 + off, dwh0 + off, idesc, 1u);
    }
}

// Layer-0 operand tile of one sample row. Thread (row, grp) builds the 8-column cores grp, grp+4, ... of the lattice part and
// the two tail cores T_grp, T_{grp+4} (tail = concat points | SH | normal | geom, 64 columns = 8 cores after column 2L).
__device__ __forceinline__ void build_input_tile(const RgbParams& P, const LevelC* lc, const float2* __restrict__ lattice,
                                                 const float* __restrict__ pos, const float* __restrict__ dirs,
                                                 const float* __restrict__ sdf_grad, const float* __restrict__ geom, int n, bool valid, int row,
                                                 int grp, uint8_t* a_hi, uint8_t* a_lo) {
    float x[3] = {0.f, 0.f, 0.f};
    if (valid) { x[0] = pos[(size_t)n * 3]; x[1] = pos[(size_t)n * 3 + 1]; x[2] = pos[(size_t)n * 3 + 2]; }
    const int level_cores = P.L / 4;
    for (int kc = grp; kc < level_cores; kc += kRgbGroups) {
        float fv[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int l = kc * 4 + q;
            float cf[3], e[4];
#pragma unroll
            for (int i = 0; i < 3; i++) cf[i] = __fmul_rn(__fadd_rn(x[i], lc->shift[l * 4 + i]), lc->scale[l * 4 + i]);
            elevate3(cf, e);
            Simplex3 s;
            locate3(e, s);
            const float2* tab = lattice + (size_t)l * P.T;
            float2 v[4];
#pragma unroll
            for (int r = 0; r < 4; r++) v[r] = __ldg(tab + vindex3(s, r, P.cap_mask, (unsigned)P.T));
            const float w = lc->window[l];
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int r = 0; r < 4; r++) { const float wr = s.bary[r] * w; a0 = fmaf(v[r].x, wr, a0); a1 = fmaf(v[r].y, wr, a1); }
            fv[2 * q] = a0; fv[2 * q + 1] = a1;
        }
        store8w(a_hi, a_lo, row, kc, fv);
    }
    // ---- tail cores
    float t0[8], t1[8];
    const float* gp = geom + (size_t)n * kGeomDim + grp * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) t1[i] = valid ? gp[i] : 0.f;                 // T_{grp+4}: geom[8 grp .. 8 grp + 7]
    float sh[kShCols];
    {
        float dx = 0.f, dy = 0.f, dz = 0.f;
        if (valid) { dx = dirs[(size_t)n * 3]; dy = dirs[(size_t)n * 3 + 1]; dz = dirs[(size_t)n * 3 + 2]; }
        psdf::sh_eval(dx, dy, dz, 5, sh);
        if (!valid) {
#pragma unroll
            for (int i = 0; i < kShCols; i++) sh[i] = 0.f;
        }
    }
    if (grp == 0) {
        t0[0] = x[0] * P.points_scaling; t0[1] = x[1] * P.points_scaling; t0[2] = x[2] * P.points_scaling; t0[3] = 0.f;
        t0[4] = sh[0]; t0[5] = sh[1]; t0[6] = sh[2]; t0[7] = sh[3];
    } else if (grp == 1) {
#pragma unroll
        for (int i = 0; i < 8; i++) t0[i] = sh[4 + i];
    } else if (grp == 2) {
#pragma unroll
        for (int i = 0; i < 8; i++) t0[i] = sh[12 + i];
    } else {
#pragma unroll
        for (int i = 0; i < 5; i++) t0[i] = sh[20 + i];
        float gx = 0.f, gy = 0.f, gz = 0.f;
        if (valid) { gx = sdf_grad[(size_t)n * 3]; gy = sdf_grad[(size_t)n * 3 + 1]; gz = sdf_grad[(size_t)n * 3 + 2]; }
        const float inv = 1.0f / fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);       // F.normalize(eps = 1e-12)
        t0[5] = gx * inv; t0[6] = gy * inv; t0[7] = gz * inv;
    }
    store8w(a_hi, a_lo, row, level_cores + grp, t0);
    store8w(a_hi, a_lo, row, level_cores + 4 + grp, t1);
    // zero padding up to Kp[0] (in_dim is a multiple of 8; a multiple of 16 only for even L/4... pad the odd core)
    const int cores = P.g.Kp[0] / 8;
    for (int kc = level_cores + 8 + grp; kc < cores; kc += kRgbGroups) {
        float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store8w(a_hi, a_lo, row, kc, z8);
    }
}

}  // namespace psdf_rgb
