// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// Host shim that launches the UNMODIFIED reference CUDA kernels (the header-only
// /root/reference/kernels/permuto_sdf/*GPU.cuh files, included from where they lie) behind a
// plain C ABI so the tests can compare our kernels against the reference's own device code on
// the same inputs, and so tests/golden/make_ref_golden.py can produce reference-made fixtures.
//
// Built by oracle/ref_shim/Makefile into oracle/_ref/libpsdf_ref_gpu.so (git-ignored, travels to the
// GPU box with gpurun). No reference source is copied into this repository: the headers are only
// #included at build time. The reference's own host launchers (src/*.cu) cannot be built here
// (EasyPBR / loguru / Eigen / DataLoaders are absent), so this file re-creates just their launch
// glue: grid = ceil(n/256) blocks of 256 threads on the given stream (src/OccupancyGrid.cu:88-117 etc.),
// output initial values as in the reference (zeros/ones/empty), and the pcg32 passed by value.
#include <torch/torch.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "permuto_sdf/OccupancyGridGPU.cuh"
#include "permuto_sdf/RaySamplerGPU.cuh"
#include "permuto_sdf/RaySamplesPackedGPU.cuh"
#include "permuto_sdf/SphereGPU.cuh"
#include "permuto_sdf/VolumeRenderingGPU.cuh"
#include "permuto_sdf/PermutoSDFGPU.cuh"

namespace {
template <typename T>
torch::PackedTensorAccessor32<T, 1, torch::RestrictPtrTraits> A1(const T* p, int64_t n) {
    int64_t s[1] = {n}, st[1] = {1};
    return torch::PackedTensorAccessor32<T, 1, torch::RestrictPtrTraits>(const_cast<T*>(p), s, st);
}
template <typename T>
torch::PackedTensorAccessor32<T, 2, torch::RestrictPtrTraits> A2(const T* p, int64_t n, int64_t c) {
    int64_t s[2] = {n, c}, st[2] = {c, 1};
    return torch::PackedTensorAccessor32<T, 2, torch::RestrictPtrTraits>(const_cast<T*>(p), s, st);
}
template <typename T>
torch::PackedTensorAccessor32<T, 3, torch::RestrictPtrTraits> A3(const T* p, int64_t a, int64_t b, int64_t c) {
    int64_t s[3] = {a, b, c}, st[3] = {b * c, c, 1};
    return torch::PackedTensorAccessor32<T, 3, torch::RestrictPtrTraits>(const_cast<T*>(p), s, st);
}
template <typename T>
torch::PackedTensorAccessor32<T, 4, torch::RestrictPtrTraits> A4(const T* p, int64_t a, int64_t b, int64_t c, int64_t d) {
    int64_t s[4] = {a, b, c, d}, st[4] = {b * c * d, c * d, d, 1};
    return torch::PackedTensorAccessor32<T, 4, torch::RestrictPtrTraits>(const_cast<T*>(p), s, st);
}
inline unsigned nb(int n) { return (unsigned)((n + 255) / 256); }
inline pcg32 mk_rng(uint64_t state, uint64_t inc) { pcg32 r; r.state = state; r.inc = inc; return r; }
#define ST ((cudaStream_t)stream)
#define RET() return (int)cudaGetLastError()
}  // namespace

extern "C" {

// ----- Sphere (src/Sphere.cu:42-109) -----
int ref_sphere_ray_intersection(int n, float radius, const float* center_dev, const float* o, const float* d, float* pe,
                                float* te, float* px, float* tx, bool* hit, void* stream) {
    ray_intersection_gpu<<<nb(n), 256, 0, ST>>>(n, radius, A1(center_dev, 3), A2(o, n, 3), A2(d, n, 3), A2(pe, n, 3),
                                                 A2(te, n, 1), A2(px, n, 3), A2(tx, n, 1), A2(hit, n, 1));
    RET();
}
int ref_sphere_rand_points_inside(int n, float radius, const float* center_dev, const float* phi, const float* ct,
                                  const float* u, float* pts, void* stream) {
    rand_points_inside_gpu<<<nb(n), 256, 0, ST>>>(n, radius, A1(center_dev, 3), A1(phi, n), A1(ct, n), A1(u, n), A2(pts, n, 3));
    RET();
}

// ----- OccupancyGrid (src/OccupancyGrid.cu) -----
int ref_occ_compute_grid_points(int nr_voxels, int V, float extent, const float* trans_dev, uint64_t rs, uint64_t ri,
                                bool randomize, float* out, void* stream) {
    OccupancyGridGPU::compute_grid_points_gpu<<<nb(nr_voxels), 256, 0, ST>>>(nr_voxels, V, extent, A1(trans_dev, 3),
                                                                            mk_rng(rs, ri), randomize, A2(out, nr_voxels, 3));
    RET();
}
int ref_occ_compute_random_sample_of_grid_points(int n, int V, float extent, const float* trans_dev, const int* idx,
                                                 uint64_t rs, uint64_t ri, bool randomize, float* out, void* stream) {
    OccupancyGridGPU::compute_random_sample_of_grid_points_gpu<<<nb(n), 256, 0, ST>>>(
        n, V, extent, A1(trans_dev, 3), A1(idx, n), mk_rng(rs, ri), randomize, A2(out, n, 3));
    RET();
}
int ref_occ_update_with_density(int nr_voxels, const float* density, int V, float decay, float thresh, float* values,
                                bool* occ, void* stream) {
    OccupancyGridGPU::update_with_density_gpu<<<nb(nr_voxels), 256, 0, ST>>>(nr_voxels, A2(density, nr_voxels, 1), V, decay,
                                                                            thresh, A1(values, nr_voxels), A1(occ, nr_voxels));
    RET();
}
int ref_occ_update_with_density_random_sample(int n, const float* density, int V, const int* idx, float decay,
                                              float thresh, float* values, bool* occ, void* stream) {
    int64_t nv = (int64_t)V * V * V;
    OccupancyGridGPU::update_with_density_random_sample_gpu<<<nb(n), 256, 0, ST>>>(n, A2(density, n, 1), V, A1(idx, n), decay,
                                                                                  thresh, A1(values, nv), A1(occ, nv));
    RET();
}
int ref_occ_update_with_sdf(int nr_voxels, const float* sdf, float extent, int V, float inv_s, float max_eik, float thresh,
                            float* values, bool* occ, void* stream) {
    OccupancyGridGPU::update_with_sdf_gpu<<<nb(nr_voxels), 256, 0, ST>>>(nr_voxels, A2(sdf, nr_voxels, 1), extent, V, inv_s,
                                                                        max_eik, thresh, A1(values, nr_voxels),
                                                                        A1(occ, nr_voxels));
    RET();
}
int ref_occ_update_with_sdf_random_sample(int n, const float* sdf, float extent, int V, const int* idx,
                                          const float* inv_s_dev, float thresh, float* values, bool* occ, void* stream) {
    int64_t nv = (int64_t)V * V * V;
    OccupancyGridGPU::update_with_sdf_random_sample_gpu<<<nb(n), 256, 0, ST>>>(n, A2(sdf, n, 1), extent, V, A1(idx, n),
                                                                              A1(inv_s_dev, 1), thresh, A1(values, nv),
                                                                              A1(occ, nv));
    RET();
}
int ref_occ_compute_samples_in_occupied_regions(int nr_rays, int V, float extent, const float* trans_dev, const float* o,
                                                const float* d, const float* te, const float* tx, const bool* occ,
                                                float min_dist, int max_per_ray, int max_nr_samples, uint64_t rs,
                                                uint64_t ri, bool jitter, float* s_pos, float* s_dirs, float* s_z,
                                                float* s_dt, float* fixed_dt, int* start_end, int* cur, void* stream) {
    int64_t nv = (int64_t)V * V * V, M = max_nr_samples;
    OccupancyGridGPU::compute_samples_in_occupied_regions_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, V, extent, A1(trans_dev, 3), A2(o, nr_rays, 3), A2(d, nr_rays, 3), A2(te, nr_rays, 1), A2(tx, nr_rays, 1),
        A1(occ, nv), min_dist, max_per_ray, max_nr_samples, mk_rng(rs, ri), jitter, A2(s_pos, M, 3), A2(s_dirs, M, 3),
        A2(s_z, M, 1), A2(s_dt, M, 1), A2(fixed_dt, nr_rays, 1), A2(start_end, nr_rays, 2), A1(cur, 1));
    RET();
}
int ref_occ_compute_first_sample_start_of_occupied_regions(int nr_rays, int V, float extent, const float* trans_dev,
                                                           const float* o, const float* d, const float* te,
                                                           const float* tx, const bool* occ, int max_nr_samples,
                                                           float* s_pos, float* s_dirs, float* s_z, float* s_dt,
                                                           float* fixed_dt, int* start_end, int* cur, void* stream) {
    int64_t nv = (int64_t)V * V * V, M = max_nr_samples;
    OccupancyGridGPU::compute_first_sample_start_of_occupied_regions_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, V, extent, A1(trans_dev, 3), A2(o, nr_rays, 3), A2(d, nr_rays, 3), A2(te, nr_rays, 1), A2(tx, nr_rays, 1),
        A1(occ, nv), max_nr_samples, A2(s_pos, M, 3), A2(s_dirs, M, 3), A2(s_z, M, 1), A2(s_dt, M, 1),
        A2(fixed_dt, nr_rays, 1), A2(start_end, nr_rays, 2), A1(cur, 1));
    RET();
}
int ref_occ_advance_sample_to_next_occupied_voxel(int n, int V, float extent, const float* trans_dev, const float* dirs,
                                                  float* pos_io, const bool* occ, bool* within, void* stream) {
    int64_t nv = (int64_t)V * V * V;
    OccupancyGridGPU::advance_sample_to_next_occupied_voxel_gpu<<<nb(n), 256, 0, ST>>>(
        n, V, extent, A1(trans_dev, 3), A2(dirs, n, 3), A2(pos_io, n, 3), A1(occ, nv), A2(pos_io, n, 3), A2(within, n, 1));
    RET();
}
int ref_occ_check_occupancy(int n, int V, float extent, const float* trans_dev, const bool* occ, const float* pts,
                            bool* out, void* stream) {
    int64_t nv = (int64_t)V * V * V;
    OccupancyGridGPU::check_occupancy_gpu<<<nb(n), 256, 0, ST>>>(n, V, extent, A1(trans_dev, 3), A1(occ, nv), A2(pts, n, 3),
                                                                A2(out, n, 1));
    RET();
}

// ----- RaySamplesPacked (src/RaySamplesPacked.cu:57-146) -----
int ref_packed_compact(int nr_rays, int M_in, int M_out, const float* pos, const float* pos4, const float* dirs,
                       const float* z, const float* dt, const float* sdf, const float* fixed_dt, const int* start_end,
                       float* o_pos, float* o_pos4, float* o_dirs, float* o_z, float* o_dt, float* o_sdf,
                       float* o_fixed_dt, int* o_start_end, int* o_cur, void* stream) {
    RaySamplesPackedGPU::compact_to_valid_samples_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, A2(pos, M_in, 3), A2(pos4, M_in, 4), A2(dirs, M_in, 3), A2(z, M_in, 1), A2(dt, M_in, 1), A2(sdf, M_in, 1),
        A2(fixed_dt, nr_rays, 1), A2(start_end, nr_rays, 2), A2(o_pos, M_out, 3), A2(o_pos4, M_out, 4), A2(o_dirs, M_out, 3),
        A2(o_z, M_out, 1), A2(o_dt, M_out, 1), A2(o_sdf, M_out, 1), A2(o_fixed_dt, nr_rays, 1), A2(o_start_end, nr_rays, 2),
        A1(o_cur, 1));
    RET();
}
int ref_packed_per_sample_ray_idx(int nr_rays, int nr_samples, const int* start_end, int* out, void* stream) {
    RaySamplesPackedGPU::compute_per_sample_ray_idx_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, nr_samples,
                                                                                  A2(start_end, nr_rays, 2),
                                                                                  A1(out, nr_samples));
    RET();
}

// ----- RaySampler (src/RaySampler.cu:37-152) -----
int ref_sampler_bg(int nr_rays, int n_per_ray, const float* o, const float* d, const float* tx, float radius,
                   const float* center_dev, uint64_t rs, uint64_t ri, bool randomize, bool contract, float* s3, float* s4,
                   float* s_dirs, float* s_z, float* s_dt, float* fixed_dt, int* start_end, void* stream) {
    RaySamplerGPU::compute_samples_bg_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, n_per_ray, A2(o, nr_rays, 3), A2(d, nr_rays, 3), A2(tx, nr_rays, 1), radius, A1(center_dev, 3),
        mk_rng(rs, ri), randomize, contract, A3(s3, nr_rays, n_per_ray, 3), A3(s4, nr_rays, n_per_ray, 4),
        A3(s_dirs, nr_rays, n_per_ray, 3), A2(s_z, nr_rays, n_per_ray), A2(s_dt, nr_rays, n_per_ray),
        A2(fixed_dt, nr_rays, 1), A2(start_end, nr_rays, 2));
    RET();
}
int ref_sampler_fg(int nr_rays, const float* o, const float* d, const float* te, const float* tx, float radius,
                   const float* center_dev, float min_dist, int max_per_ray, int max_nr_samples, uint64_t rs, uint64_t ri,
                   bool jitter, float* s_pos, float* s_dirs, float* s_z, float* s_dt, float* fixed_dt, int* start_end,
                   int* cur, void* stream) {
    int64_t M = max_nr_samples;
    RaySamplerGPU::compute_samples_fg_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, A2(o, nr_rays, 3), A2(d, nr_rays, 3), A2(te, nr_rays, 1), A2(tx, nr_rays, 1), radius, A1(center_dev, 3),
        min_dist, max_per_ray, max_nr_samples, mk_rng(rs, ri), jitter, A2(s_pos, M, 3), A2(s_dirs, M, 3), A2(s_z, M, 1),
        A2(s_dt, M, 1), A2(fixed_dt, nr_rays, 1), A2(start_end, nr_rays, 2), A1(cur, 1));
    RET();
}

// ----- VolumeRendering (src/VolumeRendering.cu) ; N = samples in the packed container -----
#define RSP int nr_rays, int N, int max_nr_samples, const int* start_end, bool equal, int fixed_n
#define SE A2(start_end, nr_rays, 2)
int ref_vr_cumprod_alpha2transmittance(RSP, const float* alpha, float* T, float* bgT, void* stream) {
    VolumeRenderingGPU::cumprod_alpha2transmittance_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, equal, fixed_n,
                                                                                    A2(alpha, N, 1), A2(T, N, 1),
                                                                                    A2(bgT, nr_rays, 1));
    RET();
}
int ref_vr_integrate_with_weights(RSP, const float* vals, const float* w, float* out, void* stream) {
    VolumeRenderingGPU::integrate_with_weights_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, equal, fixed_n,
                                                                               A2(vals, N, 3), A2(w, N, 1),
                                                                               A2(out, nr_rays, 3));
    RET();
}
int ref_vr_sdf2alpha(RSP, const float* fixed_dt, const float* dt, const float* sdf, float inv_s, bool dynamic_inv_s,
                     float inv_s_mult, float* alpha, void* stream) {
    VolumeRenderingGPU::sdf2alpha_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, A2(fixed_dt, nr_rays, 1),
                                                                  A2(dt, N, 1), equal, fixed_n, A2(sdf, N, 1), inv_s,
                                                                  dynamic_inv_s, inv_s_mult, A2(alpha, N, 1));
    RET();
}
int ref_vr_sum_over_each_ray(RSP, int val_dim, const float* vals, float* sum_ray, float* sum_sample, void* stream) {
#define SUMK(D)                                                                                                            \
    VolumeRenderingGPU::sum_over_each_ray_gpu<D><<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, equal, fixed_n,   \
                                                                             A2(vals, N, D), A2(sum_ray, nr_rays, D),      \
                                                                             A2(sum_sample, N, D))
    if (val_dim == 1) SUMK(1); else if (val_dim == 2) SUMK(2); else if (val_dim == 3) SUMK(3); else if (val_dim == 32) SUMK(32);
    else return -1;
    RET();
}
int ref_vr_cumsum_over_each_ray(RSP, const float* vals, bool inverse, float* out, void* stream) {
    VolumeRenderingGPU::cumsum_over_each_ray_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, equal, fixed_n,
                                                                             A2(vals, N, 1), inverse, A2(out, N, 1));
    RET();
}
int ref_vr_compute_cdf(RSP, const float* w, float* cdf, void* stream) {
    VolumeRenderingGPU::compute_cdf_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, max_nr_samples, SE, equal, fixed_n, A2(w, N, 1),
                                                                    A2(cdf, N, 1));
    RET();
}
int ref_vr_importance_sample(RSP, const float* o, const float* d, const float* fixed_dt, const float* z, const float* cdf,
                             int nr_imp, uint64_t rs, uint64_t ri, bool jitter, float* o_pos, float* o_dirs, float* o_z,
                             int* o_start_end, void* stream) {
    int64_t Mi = (int64_t)nr_rays * nr_imp;
    VolumeRenderingGPU::importance_sample_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, A2(o, nr_rays, 3), A2(d, nr_rays, 3), max_nr_samples, SE, A2(fixed_dt, nr_rays, 1), equal, fixed_n,
        A2(z, N, 1), A2(cdf, N, 1), nr_imp, mk_rng(rs, ri), jitter, A2(o_pos, Mi, 3), A2(o_dirs, Mi, 3), A2(o_z, Mi, 1),
        A2(o_start_end, nr_rays, 2));
    RET();
}
int ref_vr_combine_uniform_samples_with_imp(RSP, const float* o, const float* d, const float* tx, const float* u_fixed_dt,
                                            const float* u_z, const float* u_sdf, bool u_has_sdf, int imp_n,
                                            const float* i_z, const float* i_sdf, bool i_has_sdf, int c_max, float* c_pos,
                                            float* c_dirs, float* c_z, float* c_dt, float* c_sdf, float* c_fixed_dt,
                                            int* c_start_end, int* c_cur, void* stream) {
    int64_t Mi = (int64_t)nr_rays * imp_n, Mc = c_max;
    VolumeRenderingGPU::combine_uniform_samples_with_imp_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, A2(o, nr_rays, 3), A2(d, nr_rays, 3), A2(tx, nr_rays, 1), max_nr_samples, SE, A2(u_fixed_dt, nr_rays, 1),
        equal, fixed_n, A2(u_z, N, 1), A2(u_sdf, N, 1), u_has_sdf, (int)Mi, A2(start_end, nr_rays, 2),
        A2(u_fixed_dt, nr_rays, 1), true, imp_n, A2(i_z, Mi, 1), A2(i_sdf, Mi, 1), i_has_sdf, c_max, A2(c_pos, Mc, 3),
        A2(c_dirs, Mc, 3), A2(c_z, Mc, 1), A2(c_dt, Mc, 1), A2(c_sdf, Mc, 1), A2(c_fixed_dt, nr_rays, 1),
        A2(c_start_end, nr_rays, 2), A1(c_cur, 1));
    RET();
}
int ref_vr_cumprod_alpha2transmittance_backward(RSP, const float* gT, const float* gbg, const float* alpha, const float* T,
                                                const float* bgT, const float* cumsumLV, float* g_alpha, void* stream) {
    VolumeRenderingGPU::cumprod_alpha2transmittance_backward_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, max_nr_samples, SE, equal, fixed_n, A2(gT, N, 1), A2(gbg, nr_rays, 1), A2(alpha, N, 1), A2(T, N, 1),
        A2(bgT, nr_rays, 1), A2(cumsumLV, N, 1), A2(g_alpha, N, 1));
    RET();
}
int ref_vr_integrate_with_weights_backward(RSP, const float* g_pred, const float* vals, const float* w, const float* pred,
                                           float* g_vals, float* g_w, void* stream) {
    VolumeRenderingGPU::integrate_with_weights_backward_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, max_nr_samples, SE, equal, fixed_n, A2(g_pred, nr_rays, 3), A2(vals, N, 3), A2(w, N, 1), A2(pred, nr_rays, 3),
        A2(g_vals, N, 3), A2(g_w, N, 1));
    RET();
}
int ref_vr_sum_over_each_ray_backward(RSP, int val_dim, const float* g_ray, const float* g_sample, const float* vals,
                                      float* g_vals, void* stream) {
#define SUMB(D)                                                                                                          \
    VolumeRenderingGPU::sum_over_each_ray_backward_gpu<D><<<nb(nr_rays), 256, 0, ST>>>(                                    \
        nr_rays, max_nr_samples, SE, equal, fixed_n, A2(g_ray, nr_rays, D), A2(g_sample, N, D), A2(vals, N, D), A2(g_vals, N, D))
    if (val_dim == 1) SUMB(1); else if (val_dim == 2) SUMB(2); else if (val_dim == 3) SUMB(3); else return -1;
    RET();
}
int ref_vr_compute_dt(RSP, bool use_t_exit, const float* tx, const float* z, float* dt, void* stream) {
    VolumeRenderingGPU::compute_dt_gpu<<<nb(nr_rays), 256, 0, ST>>>(nr_rays, use_t_exit, A2(tx, nr_rays, 1), max_nr_samples,
                                                                   A2(z, N, 1), SE, equal, fixed_n, A2(dt, N, 1));
    RET();
}
int ref_vr_volume_render_nerf(RSP, const float* tx, const float* rgb, const float* radiance, const float* z,
                              const float* dt, float* pred_rgb, float* pred_depth, float* bgT, float* w, void* stream) {
    VolumeRenderingGPU::volume_render_nerf<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, false, A2(tx, nr_rays, 1), A2(rgb, N, 3), A2(radiance, N, 1), max_nr_samples, A2(z, N, 1), A2(dt, N, 1), SE,
        equal, fixed_n, A2(pred_rgb, nr_rays, 3), A2(pred_depth, nr_rays, 1), A2(bgT, nr_rays, 1), A2(w, N, 1));
    RET();
}
int ref_vr_volume_render_nerf_backward(RSP, const float* g_rgb_ray, const float* g_bgT, const float* g_w,
                                       const float* pred_rgb, const float* tx, const float* bgT, const float* rgb,
                                       const float* radiance, const float* dt, float* g_rgb, float* g_rad, void* stream) {
    VolumeRenderingGPU::volume_render_nerf_backward<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, false, A2(g_rgb_ray, nr_rays, 3), A2(g_bgT, nr_rays, 1), A2(g_w, N, 1), A2(pred_rgb, nr_rays, 3),
        A2(tx, nr_rays, 1), A2(bgT, nr_rays, 1), A2(rgb, N, 3), A2(radiance, N, 1), max_nr_samples, A2(dt, N, 1), SE, equal,
        fixed_n, A2(g_rgb, N, 3), A2(g_rad, N, 1));
    RET();
}

// ----- PermutoSDF statics (src/PermutoSDF.cu:67-112,167-204) -----
int ref_spherical_harmonics(int n, int degree, const float* dirs, float* out, void* stream) {
    spherical_harmonics_gpu<<<nb(n), 256, 0, ST>>>(n, degree, A2(dirs, n, 3), A2(out, n, degree * degree));
    RET();
}
int ref_random_rays_from_reel(int nr_rays, int nr_images, int H, int W, const float* rgb_reel, const float* mask_reel,
                              const float* K, const float* tf, const int* pix, const int* img, bool has_mask, float* o,
                              float* d, float* gt_rgb, float* gt_mask, void* stream) {
    random_rays_from_reel_gpu<<<nb(nr_rays), 256, 0, ST>>>(
        nr_rays, nr_images, H, W, A4(rgb_reel, nr_images, 3, H, W), A4(mask_reel, nr_images, 1, H, W), A3(K, nr_images, 3, 3),
        A3(tf, nr_images, 4, 4), A1(pix, nr_rays), A1(img, nr_rays), has_mask, A2(o, nr_rays, 3), A2(d, nr_rays, 3),
        A2(gt_rgb, nr_rays, 3), A2(gt_mask, nr_rays, 1));
    RET();
}

}  // extern "C"
