// Small fused pieces around the colour network (sm_100a), each replacing a chain of tiny element-wise PyTorch launches:
//   * Lipschitz weight normalisation of LipshitzMLP (permuto_sdf_py/models/models.py:96-110):
//       W_eff[r,:] = W[r,:] * min(1, softplus(c) / sum_j |W[r,j]|)           forward + backward, one warp per row
//   * per-image colour calibration + sigmoid on packed samples (models.py:395-414, Colorcal :677-741):
//       rgb = sigmoid(x * (1 + weight_delta[img]) + bias[img]),  identity for img == idx_with_fixed_calib
//     forward + backward, one warp per ray (the per-image parameter gradients are reduced per ray, then one atomic per channel)
#include "common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf;

namespace {
constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }   // torch threshold 20

__global__ void __launch_bounds__(kThreads)
k_lipschitz_forward(int rows, int cols, const float* __restrict__ W, const float* __restrict__ c, float* __restrict__ W_eff) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float a = 0.f;
    for (int j = lane; j < cols; j += 32) a += fabsf(W[(size_t)r * cols + j]);
    a = warp_sum(a);
    const float s = fminf(softplus_f(c[0]) / a, 1.0f);
    for (int j = lane; j < cols; j += 32) W_eff[(size_t)r * cols + j] = W[(size_t)r * cols + j] * s;
}
// dW = G s - [ratio <= 1] sp / A^2 sign(W) D,  dc += [ratio <= 1] D / A * sigmoid(c),  D = sum_j G_j W_j
__global__ void __launch_bounds__(kThreads)
k_lipschitz_backward(int rows, int cols, const float* __restrict__ W, const float* __restrict__ c, const float* __restrict__ G,
                     float* __restrict__ gW, float* __restrict__ gc) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (r >= rows) return;
    float a = 0.f, d = 0.f;
    for (int j = lane; j < cols; j += 32) {
        const float w = W[(size_t)r * cols + j];
        a += fabsf(w);
        d += G[(size_t)r * cols + j] * w;
    }
    a = warp_sum(a);
    d = warp_sum(d);
    const float cv = c[0], sp = softplus_f(cv);
    const float ratio = sp / a;
    const bool active = ratio <= 1.0f;               // torch.clamp(max=1) passes the gradient where ratio <= 1
    const float s = fminf(ratio, 1.0f);
    const float k = active ? sp / (a * a) * d : 0.0f;
    for (int j = lane; j < cols; j += 32) {
        const float w = W[(size_t)r * cols + j];
        const float sg = w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f);
        gW[(size_t)r * cols + j] = G[(size_t)r * cols + j] * s - k * sg;
    }
    if (lane == 0 && active && gc) {
        const float sig = cv > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-cv));
        atomicAdd(gc, d / a * sig);
    }
}

__global__ void __launch_bounds__(kThreads)
k_calib_sigmoid_forward(int nr_rays, int max_nr_samples, const int* __restrict__ start_end, bool equal, int fixed_n,
                        const float* __restrict__ x, const int* __restrict__ img_idx, const float* __restrict__ weight_delta,
                        const float* __restrict__ bias, int fixed_img, float* __restrict__ out) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= nr_rays) return;
    RayRange rr = ray_range(ray, start_end, equal, fixed_n);
    if (rr.end > max_nr_samples || rr.n <= 0) return;
    float w[3] = {1.f, 1.f, 1.f}, b[3] = {0.f, 0.f, 0.f};
    if (img_idx) {
        const int im = img_idx[ray];
        if (im != fixed_img) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) { w[ch] = 1.0f + weight_delta[3 * im + ch]; b[ch] = bias[3 * im + ch]; }
        }
    }
    for (int e = lane; e < rr.n * 3; e += 32) {
        const int ch = e % 3;
        const size_t i = (size_t)rr.start * 3 + e;
        const float v = x[i] * w[ch] + b[ch];
        out[i] = 1.0f / (1.0f + expf(-v));
    }
}
__global__ void __launch_bounds__(kThreads)
k_calib_sigmoid_backward(int nr_rays, int max_nr_samples, const int* __restrict__ start_end, bool equal, int fixed_n,
                         const float* __restrict__ x, const float* __restrict__ out, const float* __restrict__ g_out,
                         const int* __restrict__ img_idx, const float* __restrict__ weight_delta, int fixed_img, float* __restrict__ g_x,
                         float* __restrict__ g_weight_delta, float* __restrict__ g_bias) {
    const int ray = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ray >= nr_rays) return;
    RayRange rr = ray_range(ray, start_end, equal, fixed_n);
    if (rr.end > max_nr_samples || rr.n <= 0) return;
    float w[3] = {1.f, 1.f, 1.f};
    int im = -1;
    if (img_idx) {
        im = img_idx[ray];
        if (im != fixed_img) {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) w[ch] = 1.0f + weight_delta[3 * im + ch];
        } else im = -1;
    }
    float gw[3] = {0.f, 0.f, 0.f}, gb[3] = {0.f, 0.f, 0.f};
    // 96 = lcm(32, 3): within a 96-element chunk the channel of (q, lane) does not depend on the chunk
    for (int e0 = 0; e0 < rr.n * 3; e0 += 96) {
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int e = e0 + q * 32 + lane;
            if (e < rr.n * 3) {
                const int c = (q * 32 + lane) % 3;
                const size_t i = (size_t)rr.start * 3 + e;
                const float o = out[i];
                const float gv = g_out[i] * o * (1.0f - o);
                g_x[i] = gv * w[c];
                gw[c] += gv * x[i];
                gb[c] += gv;
            }
        }
    }
    if (im >= 0 && g_weight_delta) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float sw = warp_sum(gw[c]), sb = warp_sum(gb[c]);
            if (lane == 0) { atomicAdd(g_weight_delta + 3 * im + c, sw); atomicAdd(g_bias + 3 * im + c, sb); }
        }
    }
}
#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {
int psdf_lipschitz_normalize(int rows, int cols, const float* W, const float* c, float* W_eff, void* stream) {
    if (rows <= 0 || cols <= 0) return PSDF_ERR_ARG;
    k_lipschitz_forward<<<div_up((long long)rows * 32, kThreads), kThreads, 0, ST>>>(rows, cols, W, c, W_eff);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
int psdf_lipschitz_normalize_backward(int rows, int cols, const float* W, const float* c, const float* grad_W_eff, float* grad_W,
                                      float* grad_c, void* stream) {
    if (rows <= 0 || cols <= 0) return PSDF_ERR_ARG;
    k_lipschitz_backward<<<div_up((long long)rows * 32, kThreads), kThreads, 0, ST>>>(rows, cols, W, c, grad_W_eff, grad_W, grad_c);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
int psdf_calib_sigmoid_forward(int nr_rays, int max_nr_samples, const int* ray_start_end, int equal, int fixed_n, const float* x,
                               const int* img_idx, const float* weight_delta, const float* bias, int fixed_img, float* out,
                               void* stream) {
    if (nr_rays <= 0) return nr_rays == 0 ? PSDF_OK : PSDF_ERR_ARG;
    k_calib_sigmoid_forward<<<div_up((long long)nr_rays * 32, kThreads), kThreads, 0, ST>>>(nr_rays, max_nr_samples, ray_start_end, equal != 0,
                                                                                          fixed_n, x, img_idx, weight_delta, bias, fixed_img, out);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
int psdf_calib_sigmoid_backward(int nr_rays, int max_nr_samples, const int* ray_start_end, int equal, int fixed_n, const float* x,
                                const float* out, const float* grad_out, const int* img_idx, const float* weight_delta, int fixed_img,
                                float* grad_x, float* grad_weight_delta, float* grad_bias, void* stream) {
    if (nr_rays <= 0) return nr_rays == 0 ? PSDF_OK : PSDF_ERR_ARG;
    k_calib_sigmoid_backward<<<div_up((long long)nr_rays * 32, kThreads), kThreads, 0, ST>>>(
        nr_rays, max_nr_samples, ray_start_end, equal != 0, fixed_n, x, out, grad_out, img_idx, weight_delta, fixed_img, grad_x,
        grad_weight_delta, grad_bias);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
}  // extern "C"
