// Dense fused AdamW sweep for the hash tables and MLP parameters (sm_100a).
// One pass over {param, grad, exp_avg, exp_avg_sq}: 28 B/parameter (read 16, write 12), optional gradient
// zeroing folded in (+4 B) so that no separate zero_grad memset is needed. Math = torch.optim.AdamW /
// apex FusedAdam as used by the reference (train_permuto_sdf.py:293-304: betas (0.9, 0.99), eps 1e-15,
// decoupled weight decay, no amsgrad):
//   p  <- p (1 - lr wd);  m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2
//   p  <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound streaming kernel: 128-bit loads/stores, grid = multiple of the SM count, grid-stride loop.
#include "common.cuh"
#include "../../include/psdf_b200.h"

using namespace psdf;

namespace {
__global__ void __launch_bounds__(256)
k_adamw(long long n, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float lr,
        float beta1, float beta2, float eps, float weight_decay, float bias_c1, float bias_c2_sqrt, const int* __restrict__ step_dev,
        const float* __restrict__ hyper_dev, float grad_scale, int zero_grad) {
    const long long n4 = n >> 2;
    if (hyper_dev) { lr = hyper_dev[0]; weight_decay = hyper_dev[1]; }   // schedule values kept in device memory (CUDA-graph replay)
    if (step_dev) {   // step count kept in device memory (CUDA-graph replay): bias corrections computed here
        const float t = (float)step_dev[0];
        bias_c1 = 1.0f - powf(beta1, t);
        bias_c2_sqrt = sqrtf(1.0f - powf(beta2, t));
    }
    const float decay = 1.0f - lr * weight_decay;
    const float step_size = lr / bias_c1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two independent float4 quadruples per thread and iteration: 8 x 16-byte loads in flight before the first use
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
        const long long i1 = i0 + stride;
        const bool two = i1 < n4;
        float4 P[2], G[2], M[2], V[2];
        P[0] = reinterpret_cast<float4*>(p)[i0]; G[0] = reinterpret_cast<float4*>(g)[i0];
        M[0] = reinterpret_cast<float4*>(m)[i0]; V[0] = reinterpret_cast<float4*>(v)[i0];
        if (two) {
            P[1] = reinterpret_cast<float4*>(p)[i1]; G[1] = reinterpret_cast<float4*>(g)[i1];
            M[1] = reinterpret_cast<float4*>(m)[i1]; V[1] = reinterpret_cast<float4*>(v)[i1];
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (q == 1 && !two) break;
            float* pp = &P[q].x; float* gg = &G[q].x; float* mm = &M[q].x; float* vv = &V[q].x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float gr = gg[k] * grad_scale;
                float pk = pp[k] * decay;
                mm[k] = beta1 * mm[k] + (1.0f - beta1) * gr;
                vv[k] = beta2 * vv[k] + (1.0f - beta2) * gr * gr;
                float denom = sqrtf(vv[k]) / bias_c2_sqrt + eps;
                pp[k] = pk - step_size * (mm[k] / denom);
            }
            const long long i = q ? i1 : i0;
            reinterpret_cast<float4*>(p)[i] = P[q];
            reinterpret_cast<float4*>(m)[i] = M[q];
            reinterpret_cast<float4*>(v)[i] = V[q];
            if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // tail
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        long long i = (n4 << 2) + threadIdx.x;
        float gr = g[i] * grad_scale;
        float pk = p[i] * decay;
        float mk = beta1 * m[i] + (1.0f - beta1) * gr;
        float vk = beta2 * v[i] + (1.0f - beta2) * gr * gr;
        m[i] = mk; v[i] = vk;
        p[i] = pk - step_size * (mk / (sqrtf(vk) / bias_c2_sqrt + eps));
        if (zero_grad) g[i] = 0.f;
    }
}
#define ST ((cudaStream_t)stream)
}  // namespace

extern "C" {
// step >= 1 is the (already incremented) step count, read from step_dev[0] instead when that is not NULL; pointers 16-byte aligned
int psdf_adamw_step(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, const int* step_dev, const float* hyper_dev, float grad_scale, int zero_grad,
                    void* stream) {
    if (n <= 0) return n == 0 ? PSDF_OK : PSDF_ERR_ARG;
    if (step < 1 && !step_dev) return PSDF_ERR_ARG;
    if (step < 1) step = 1;
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0) return PSDF_ERR_ARG;
    double c1 = 1.0 - pow((double)beta1, (double)step);
    double c2 = 1.0 - pow((double)beta2, (double)step);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long n4 = n >> 2;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > sms * 8) blocks = sms * 8;
    k_adamw<<<blocks, 256, 0, ST>>>(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, (float)c1, (float)sqrt(c2),
                                   step_dev, hyper_dev, grad_scale, zero_grad);
    PSDF_CHECK_LAUNCH();
    return PSDF_OK;
}
}  // extern "C"
