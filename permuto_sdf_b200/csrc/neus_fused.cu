// Fused NeuS compositing + per-ray losses for training (sm_100a), forward and backward, one warp per ray.
//
// Replaces, in one kernel each way, what the reference runs as ~60 PyTorch element-wise launches plus 8 of its own
// kernels per iteration (volume_rendering_modules.py:129-176 `VolumeRenderingNeus.compute_weights` + `integrate`,
// volume_rendering_funcs.py:55-223, and the losses of train_permuto_sdf.py:349-353,380-383 /
// permuto_sdf_utils.py:43-51):
//   true_cos = dir . g ; iter_cos = -(relu(.5 - .5 c)(1-r) + relu(-c) r)
//   est_prev/next = sdf -/+ iter_cos dt / 2 ; alpha = clip((sig(prev s) - sig(next s) + 1e-5)/(sig(prev s) + 1e-5), 0, 1)
//   T_i = prod_{j<i} (1 - alpha_j + 1e-7) ; w = alpha T ; pred_rgb = sum w rgb ; w_sum = sum w ; bg_T = T_last
//   losses per ray: sum_c |gt_c - pred_c| * hit ; BCE(clip(w_sum, 1e-3, 1-1e-3), mask) ; sum_i (|g_i| - 1)^2
// The backward kernel re-derives alpha / T (saved) and returns d loss / d {sdf, g, rgb, inv_s} for
//   loss = mean_{R,3}(rgb term) + w_eik mean_N(eik term) + w_mask mean_R(bce term).
// The transmittance recurrence runs as the same uniform serial chain as volrender.cu (left-to-right order).
#include "common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf;

namespace {
constexpr int kThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct NeusSample {
    float alpha, q, pc, nc, prev, next, ic_d;   // ic_d = d iter_cos / d true_cos
};
__device__ __forceinline__ NeusSample neus_alpha(float sdf, float gx, float gy, float gz, float dx, float dy, float dz, float dt, float s,
                                                 float r) {
    NeusSample o;
    float tc = dx * gx + dy * gy + dz * gz;
    float a = fmaxf(-tc * 0.5f + 0.5f, 0.0f), b = fmaxf(-tc, 0.0f);
    float ic = -(a * (1.0f - r) + b * r);
    o.ic_d = (tc < 1.0f ? 0.5f * (1.0f - r) : 0.0f) + (tc < 0.0f ? r : 0.0f);
    o.next = sdf + ic * dt * 0.5f;
    o.prev = sdf - ic * dt * 0.5f;
    o.pc = sigmoidf_(o.prev * s);
    o.nc = sigmoidf_(o.next * s);
    o.q = (o.pc - o.nc + 1e-5f) / (o.pc + 1e-5f);
    o.alpha = fminf(fmaxf(o.q, 0.0f), 1.0f);
    return o;
}

__global__ void __launch_bounds__(kThreads)
k_neus_forward(int nr_rays, int max_nr_samples, const int* __restrict__ start_end, bool equal, int fixed_n, const float* __restrict__ sdf,
               const float* __restrict__ grad, const float* __restrict__ rgb, const float* __restrict__ dirs, const float* __restrict__ dt_,
               const float* __restrict__ inv_s_dev, float cos_anneal, const float* __restrict__ cos_anneal_dev, const float* __restrict__ gt_rgb, const float* __restrict__ gt_mask,
               const uint8_t* __restrict__ hit, const float* __restrict__ bg_rgb, float* __restrict__ alpha_out, float* __restrict__ T_out, float* __restrict__ w_out,
               float* __restrict__ pred_rgb, float* __restrict__ w_sum, float* __restrict__ bg_T, float* __restrict__ ray_loss /* [R,3] */) {
    int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (ray >= nr_rays) return;
    RayRange rr = ray_range(ray, start_end, equal, fixed_n);
    bool skip = (rr.end > max_nr_samples) || (rr.n == 0);
    const float s = fminf(fmaxf(inv_s_dev[0], 1e-6f), 1e6f);
    if (cos_anneal_dev) cos_anneal = cos_anneal_dev[0];
    float T = 1.0f, ax = 0.f, ay = 0.f, az = 0.f, ws = 0.f, eik = 0.f;
    if (!skip) {
        for (int base = 0; base < rr.n; base += 32) {
            int i = base + lane;
            float al = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
            if (i < rr.n) {
                int k = rr.start + i;
                float gx = grad[3 * k], gy = grad[3 * k + 1], gz = grad[3 * k + 2];
                NeusSample ns = neus_alpha(sdf[k], gx, gy, gz, dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2], dt_[k], s, cos_anneal);
                al = ns.alpha;
                cr = rgb[3 * k]; cg = rgb[3 * k + 1]; cb = rgb[3 * k + 2];
                float nrm = sqrtf(gx * gx + gy * gy + gz * gz) - 1.0f;
                eik += nrm * nrm;
                alpha_out[k] = al;
            }
            float myT = 1.0f;
            int cnt = min(32, rr.n - base);
            for (int k = 0; k < cnt; k++) {
                float ak = __shfl_sync(kFull, al, k);
                if (lane == k) myT = T;
                float wk = ak * T;
                ax = fmaf(wk, __shfl_sync(kFull, cr, k), ax);
                ay = fmaf(wk, __shfl_sync(kFull, cg, k), ay);
                az = fmaf(wk, __shfl_sync(kFull, cb, k), az);
                ws += wk;
                if (base + k < rr.n - 1) T *= (1.0f - ak + 1e-7f);
            }
            if (i < rr.n) { T_out[rr.start + i] = myT; w_out[rr.start + i] = al * myT; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) eik += __shfl_xor_sync(kFull, eik, o);
    }
    if (lane == 0) {
        if (bg_rgb) { ax = fmaf(T, bg_rgb[3 * ray], ax); ay = fmaf(T, bg_rgb[3 * ray + 1], ay); az = fmaf(T, bg_rgb[3 * ray + 2], az); }
        pred_rgb[3 * ray] = ax; pred_rgb[3 * ray + 1] = ay; pred_rgb[3 * ray + 2] = az;
        w_sum[ray] = ws;
        bg_T[ray] = T;
        float h = hit ? (hit[ray] ? 1.0f : 0.0f) : 1.0f;
        float l1 = (fabsf(gt_rgb[3 * ray] - ax) + fabsf(gt_rgb[3 * ray + 1] - ay) + fabsf(gt_rgb[3 * ray + 2] - az)) * h;
        float bce = 0.0f;
        if (gt_mask) {
            float p = fminf(fmaxf(ws, 1e-3f), 1.0f - 1e-3f), m = gt_mask[ray];
            bce = -(m * fmaxf(logf(p), -100.0f) + (1.0f - m) * fmaxf(logf(1.0f - p), -100.0f));
        }
        ray_loss[3 * ray] = l1; ray_loss[3 * ray + 1] = bce; ray_loss[3 * ray + 2] = eik;
    }
}

// scale_rgb = g_total / (3R), scale_mask = g_total w_mask / R, scale_eik = g_total w_eik / N
__global__ void __launch_bounds__(kThreads)
k_neus_backward(int nr_rays, int max_nr_samples, const int* __restrict__ start_end, bool equal, int fixed_n, const float* __restrict__ sdf,
                const float* __restrict__ grad, const float* __restrict__ rgb, const float* __restrict__ dirs, const float* __restrict__ dt_,
                const float* __restrict__ inv_s_dev, float cos_anneal, const float* __restrict__ cos_anneal_dev, const float* __restrict__ gt_rgb, const float* __restrict__ gt_mask,
                const uint8_t* __restrict__ hit, const float* __restrict__ bg_rgb, const float* __restrict__ alpha_in,
                const float* __restrict__ T_in, const float* __restrict__ pred_rgb, const float* __restrict__ w_sum,
                const float* __restrict__ bg_T, const float* __restrict__ g_total_dev, float scale_rgb,
                float scale_mask, float scale_eik, const int* __restrict__ nr_samples_dev, float* __restrict__ g_sdf, float* __restrict__ g_grad, float* __restrict__ g_rgb,
                float* __restrict__ g_bg_rgb, float* __restrict__ g_inv_s) {
    int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (ray >= nr_rays) return;
    RayRange rr = ray_range(ray, start_end, equal, fixed_n);
    const bool skip = (rr.end > max_nr_samples) || (rr.n == 0);
    if (skip && !bg_rgb) return;
    const float gtot = g_total_dev ? g_total_dev[0] : 1.0f;
    const float s_raw = inv_s_dev[0];
    const float s = fminf(fmaxf(s_raw, 1e-6f), 1e6f);
    const bool s_active = (s_raw >= 1e-6f && s_raw <= 1e6f);
    if (cos_anneal_dev) cos_anneal = cos_anneal_dev[0];
    if (nr_samples_dev) scale_eik = scale_eik / (float)max(nr_samples_dev[0], 1);
    // upstream of the per-ray outputs
    float h = hit ? (hit[ray] ? 1.0f : 0.0f) : 1.0f;
    float gp[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float dlt = pred_rgb[3 * ray + c] - gt_rgb[3 * ray + c];
        gp[c] = gtot * scale_rgb * h * (dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f));
    }
    float gws = 0.0f;
    if (gt_mask) {
        float ws = w_sum[ray], m = gt_mask[ray];
        if (ws >= 1e-3f && ws <= 1.0f - 1e-3f) gws = gtot * scale_mask * (-(m / ws) + (1.0f - m) / (1.0f - ws));
    }
    const float ge = gtot * scale_eik;
    float gbg = 0.0f;   // d loss / d bg_T times bg_T
    if (bg_rgb) {
        float bt = bg_T[ray];
        gbg = (gp[0] * bg_rgb[3 * ray] + gp[1] * bg_rgb[3 * ray + 1] + gp[2] * bg_rgb[3 * ray + 2]) * bt;
        if (lane == 0 && g_bg_rgb) { g_bg_rgb[3 * ray] = gp[0] * bt; g_bg_rgb[3 * ray + 1] = gp[1] * bt; g_bg_rgb[3 * ray + 2] = gp[2] * bt; }
    }
    if (skip) return;
    // reverse sweep: S = sum_{k>i} gw_k w_k
    float S = 0.0f, gs_acc = 0.0f;
    const int nchunks = (rr.n + 31) / 32;
    for (int ch = nchunks - 1; ch >= 0; ch--) {
        int base = ch * 32;
        int i = base + lane;
        float al = 0.f, Ti = 0.f, gw = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
        int k = rr.start + i;
        if (i < rr.n) {
            al = alpha_in[k]; Ti = T_in[k];
            cr = rgb[3 * k]; cg = rgb[3 * k + 1]; cb = rgb[3 * k + 2];
            gw = gp[0] * cr + gp[1] * cg + gp[2] * cb + gws;
        }
        float gwww = gw * al * Ti;            // gw_i w_i
        float mySuffix = 0.0f;
        int cnt = min(32, rr.n - base);
        for (int j = cnt - 1; j >= 0; j--) {
            if (lane == j) mySuffix = S;
            S += __shfl_sync(kFull, gwww, j);
        }
        if (i < rr.n) {
            // d loss / d alpha_i : own weight + all later transmittances (the last sample does not enter any T)
            float b = fmaxf(1.0f - al + 1e-7f, 1e-6f);
            float ga = gw * Ti;
            if (i < rr.n - 1) ga -= (mySuffix + gbg) / b;
            float gx = grad[3 * k], gy = grad[3 * k + 1], gz = grad[3 * k + 2];
            float dx = dirs[3 * k], dy = dirs[3 * k + 1], dz = dirs[3 * k + 2];
            float dt = dt_[k];
            NeusSample ns = neus_alpha(sdf[k], gx, gy, gz, dx, dy, dz, dt, s, cos_anneal);
            float gq = (ns.q >= 0.0f && ns.q <= 1.0f) ? ga : 0.0f;
            float den = ns.pc + 1e-5f;
            float dq_dpc = (ns.nc) / (den * den);         // d/dpc [(pc - nc + e)/(pc + e)] = (nc)/(pc+e)^2
            float dq_dnc = -1.0f / den;
            float gpc = gq * dq_dpc * ns.pc * (1.0f - ns.pc);   // wrt (prev * s)
            float gnc = gq * dq_dnc * ns.nc * (1.0f - ns.nc);   // wrt (next * s)
            float gsdf = (gpc + gnc) * s;
            float gic = (gnc - gpc) * s * dt * 0.5f;
            float gtc = gic * ns.ic_d;
            gs_acc += gpc * ns.prev + gnc * ns.next;
            // eikonal term
            float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
            float ek = nrm > 0.f ? ge * 2.0f * (nrm - 1.0f) / nrm : 0.0f;
            g_sdf[k] = gsdf;
            g_grad[3 * k] = gtc * dx + ek * gx; g_grad[3 * k + 1] = gtc * dy + ek * gy; g_grad[3 * k + 2] = gtc * dz + ek * gz;
            float wi = al * Ti;
            g_rgb[3 * k] = gp[0] * wi; g_rgb[3 * k + 1] = gp[1] * wi; g_rgb[3 * k + 2] = gp[2] * wi;
        }
    }
    if (g_inv_s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gs_acc += __shfl_xor_sync(kFull, gs_acc, o);
        if (lane == 0 && s_active && gs_acc != 0.0f) atomicAdd(g_inv_s, gs_acc);
    }
}
#define ST ((cudaStream_t)stream)
inline int ray_blocks(int nr_rays) { return div_up((long long)nr_rays * 32, kThreads); }
}  // namespace

extern "C" {
int psdf_neus_render_loss_forward(int nr_rays, int max_nr_samples, const int* ray_start_end, int equal, int fixed_n, const float* sdf,
                                  const float* grad, const float* rgb, const float* dirs, const float* dt, const float* inv_s_dev,
                                  float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb, const float* gt_mask, const uint8_t* hit, const float* bg_rgb,
                                  float* alpha, float* transmittance, float* weights, float* pred_rgb, float* weights_sum, float* bg_transmittance,
                                  float* ray_loss, void* stream) {
    if (nr_rays <= 0) return nr_rays == 0 ? PSDF_OK : PSDF_ERR_ARG;
    k_neus_forward<<<ray_blocks(nr_rays), kThreads, 0, ST>>>(nr_rays, max_nr_samples, ray_start_end, equal != 0, fixed_n, sdf, grad, rgb, dirs, dt,
                                                            inv_s_dev, cos_anneal_ratio, cos_anneal_dev, gt_rgb, gt_mask, hit, bg_rgb, alpha, transmittance, weights,
                                                            pred_rgb, weights_sum, bg_transmittance, ray_loss);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
int psdf_neus_render_loss_backward(int nr_rays, int max_nr_samples, const int* ray_start_end, int equal, int fixed_n, const float* sdf,
                                   const float* grad, const float* rgb, const float* dirs, const float* dt, const float* inv_s_dev,
                                   float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb, const float* gt_mask, const uint8_t* hit, const float* bg_rgb,
                                   const float* alpha, const float* transmittance, const float* pred_rgb, const float* weights_sum,
                                   const float* bg_transmittance, const float* g_total_dev,
                                   float scale_rgb, float scale_mask, float scale_eik, const int* nr_samples_dev, float* g_sdf, float* g_grad, float* g_rgb,
                                   float* g_bg_rgb, float* g_inv_s, void* stream) {
    if (nr_rays <= 0) return nr_rays == 0 ? PSDF_OK : PSDF_ERR_ARG;
    k_neus_backward<<<ray_blocks(nr_rays), kThreads, 0, ST>>>(nr_rays, max_nr_samples, ray_start_end, equal != 0, fixed_n, sdf, grad, rgb, dirs, dt,
                                                             inv_s_dev, cos_anneal_ratio, cos_anneal_dev, gt_rgb, gt_mask, hit, bg_rgb, alpha,
                                                             transmittance, pred_rgb, weights_sum, bg_transmittance, g_total_dev, scale_rgb,
                                                             scale_mask, scale_eik, nr_samples_dev, g_sdf, g_grad, g_rgb, g_bg_rgb, g_inv_s);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
}  // extern "C"
