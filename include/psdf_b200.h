/*
 * psdf_b200.h -- C ABI of libpsdf_b200.so, the sm_100a (B200) implementation of PermutoSDF's per-ray
 * training / sphere-tracing hot path.
 *
 * Boundary: the reference exposes this path to Python as a pybind11 module `permuto_sdf`
 * (src/PyBridge.cxx:30-169) plus the external package `permutohedral_encoding`
 * (permuto_sdf_py/models/models.py:20,149,186). Both are re-created in Python on top of THIS header
 * (permuto_sdf_b200/permuto_sdf, permuto_sdf_b200/permutohedral_encoding); every entry point below names
 * the reference method it replaces. INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers (contiguous, fp32 / int32 / uint8 bool) unless the name ends in
 *     `[3]` (small host arrays passed by the caller); no allocation happens inside the library;
 *   - `stream` is a cudaStream_t; kernels are asynchronous on it; the library never synchronises;
 *   - return value: 0 = ok, <0 = error (PSDF_ERR_*); no global state: RNG state is passed explicitly
 *     as the (state, inc) pair of a pcg32 generator (kernels/permuto_sdf/pcg32.h:45-171). rng_inc == 0 (never a
 *     valid pcg32 stream) marks rng_state as a DEVICE pointer to uint64_t[2] = {state, inc}: the generator then lives
 *     in device memory, is advanced with psdf_rng_advance_dev, and the call can be replayed from a CUDA graph;
 *   - packed ray containers follow RaySamplesPacked (include/permuto_sdf/RaySamplesPacked.cuh:7-46):
 *     ray_start_end is int32 [R,2] = [start,end); `equal`/`fixed_n` is the rays_have_equal_nr_of_samples
 *     fast path; rays with end > max_nr_samples or zero samples are skipped
 *     (kernels/permuto_sdf/VolumeRenderingGPU.cuh:30-60,103).
 */
#ifndef PSDF_B200_H
#define PSDF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSDF_ABI_VERSION 1
#define PSDF_ERR_ARG_ -1
#define PSDF_ERR_LAUNCH_ -2
#define PSDF_ERR_UNSUPPORTED_ -3

int psdf_abi_version(void);
/* 1 when the library was built for sm_100a and a CUDA device is present, else 0 (never a CPU fallback) */
int psdf_device_ok(void);

/* ---------------------------------------------------------------- Sphere (include/permuto_sdf/Sphere.cuh:12-24) */
/* Sphere::ray_intersection, src/Sphere.cu:42-79 */
int psdf_sphere_ray_intersection(int nr_rays, float radius, const float center[3], const float* origins, const float* dirs,
                                 float* pts_entry, float* t_entry, float* pts_exit, float* t_exit, uint8_t* hit, void* stream);
/* Sphere::rand_points_inside, src/Sphere.cu:82-109 (random inputs phi,costheta,u are drawn by the caller) */
int psdf_sphere_rand_points_inside(int n, float radius, const float* phi, const float* costheta, const float* u, float* points,
                                   void* stream);
/* Sphere::check_point_inside_primitive, src/Sphere.cu:111-119 */
int psdf_sphere_check_point_inside(int n, float radius, const float center[3], const float* points, uint8_t* out, void* stream);

/* ---------------------------------------------------------------- OccupancyGrid (include/permuto_sdf/OccupancyGrid.cuh:19-64) */
/* compute_grid_points (idx==NULL, n==V^3) / compute_random_sample_of_grid_points (idx = chosen voxels),
 * src/OccupancyGrid.cu:88-117,179-208 */
int psdf_occ_compute_grid_points(int n, int V, float extent, const float trans[3], const int* idx, uint64_t rng_state,
                                 uint64_t rng_inc, int randomize, float* out, void* stream);
/* update_with_density (idx==NULL) / update_with_density_random_sample, src/OccupancyGrid.cu:368-420 */
int psdf_occ_update_with_density(int n, const float* density, const int* idx, float decay, float thresh, float* values,
                                 uint8_t* occ, void* stream);
/* update_with_sdf (idx==NULL, scalar inv_s, variant 0) / update_with_sdf_random_sample (inv_s_dev = 1-element
 * device tensor, variant 1), src/OccupancyGrid.cu:422-475 */
int psdf_occ_update_with_sdf(int n, const float* sdf, const int* idx, float extent, int V, float inv_s, const float* inv_s_dev,
                             float thresh, int random_sample_variant, float* values, uint8_t* occ, void* stream);
/* check_occupancy, src/OccupancyGrid.cu:339-365 */
int psdf_occ_check_occupancy(int n, int V, float extent, const float trans[3], const uint8_t* occ, const float* points,
                             uint8_t* out, void* stream);
/* compute_samples_in_occupied_regions, src/OccupancyGrid.cu:212-257.
 * slot_mode 1: ray i owns slots [i*max_per_ray, ...) (deterministic, needs nr_rays*max_per_ray <= max_nr_samples);
 * slot_mode 0: slots from a global atomic counter like the reference. cur_nr_samples must be zeroed by the caller. */
int psdf_occ_compute_samples_in_occupied_regions(int nr_rays, int V, float extent, const float trans[3], const float* origins,
                                                 const float* dirs, const float* t_entry, const float* t_exit,
                                                 const uint8_t* occ, float min_dist, int max_per_ray, int max_nr_samples,
                                                 uint64_t rng_state, uint64_t rng_inc, int jitter, int slot_mode,
                                                 float* s_pos, float* s_dirs, float* s_z, float* s_dt, float* ray_fixed_dt,
                                                 int* ray_start_end, int* cur_nr_samples, void* stream);
/* compute_first_sample_start_of_occupied_regions, src/OccupancyGrid.cu:259-300 */
int psdf_occ_compute_first_sample_start(int nr_rays, int V, float extent, const float trans[3], const float* origins,
                                        const float* dirs, const float* t_entry, const float* t_exit, const uint8_t* occ,
                                        int max_nr_samples, int slot_mode, float* s_pos, float* s_dirs, float* s_z, float* s_dt,
                                        float* ray_fixed_dt, int* ray_start_end, int* cur_nr_samples, void* stream);
/* advance_sample_to_next_occupied_voxel, src/OccupancyGrid.cu:302-337 (pos updated in place, like the reference) */
int psdf_occ_advance_sample_to_next_occupied_voxel(int n, int V, float extent, const float trans[3], const float* dirs,
                                                   float* pos_io, const uint8_t* occ, uint8_t* within, void* stream);

/* ---------------------------------------------------------------- RaySampler (include/permuto_sdf/RaySampler.cuh:21-22) */
int psdf_sampler_fg(int nr_rays, const float* origins, const float* dirs, const float* t_entry, const float* t_exit,
                    float min_dist, int max_per_ray, int max_nr_samples, uint64_t rng_state, uint64_t rng_inc, int jitter,
                    int slot_mode, float* s_pos, float* s_dirs, float* s_z, float* s_dt, float* ray_fixed_dt,
                    int* ray_start_end, int* cur_nr_samples, void* stream);
int psdf_sampler_bg(int nr_rays, int n_per_ray, const float* origins, const float* dirs, const float* t_exit,
                    float sphere_radius, const float sphere_center[3], uint64_t rng_state, uint64_t rng_inc, int randomize,
                    int contract, float* s3, float* s4, float* s_dirs, float* s_z, float* s_dt, float* ray_fixed_dt,
                    int* ray_start_end, void* stream);

/* device-resident generator: state <- state advanced by `delta` draws (the reference advances its static generators by 2^32
 * after every jittered call, e.g. src/OccupancyGrid.cu:254) */
int psdf_rng_advance_dev(uint64_t* rng_dev, long long delta, void* stream);

/* ---------------------------------------------------------------- RaySamplesPacked (src/RaySamplesPacked.cu:44-146) */
long long psdf_packed_compact_workspace_bytes(int nr_rays);
/* compute_exact_nr_samples without the host sync: total_dev <- sum(end-start) */
int psdf_packed_count_samples(int nr_rays, const int* ray_start_end, int* total_dev, void* stream);
/* compact_to_valid_samples = scan (fills workspace; workspace[nr_rays + nblocks] is the total) + copy */
int psdf_packed_compact_scan(int nr_rays, const int* ray_start_end, int* workspace, void* stream);
int psdf_packed_compact_copy(int nr_rays, const float* pos, const float* pos4, const float* dirs, const float* z,
                             const float* dt, const float* sdf, const float* fixed_dt, const int* ray_start_end,
                             const int* workspace, float* o_pos, float* o_pos4, float* o_dirs, float* o_z, float* o_dt,
                             float* o_sdf, float* o_fixed_dt, int* o_start_end, void* stream);
/* compute_per_sample_ray_idx */
int psdf_packed_per_sample_ray_idx(int nr_rays, int nr_samples, const int* ray_start_end, int* out, void* stream);

/* ---------------------------------------------------------------- VolumeRendering (include/permuto_sdf/VolumeRendering.cuh:22-38) */
#define PSDF_RSP int nr_rays, int max_nr_samples, const int* ray_start_end, int equal, int fixed_n
int psdf_vr_cumprod_alpha2transmittance(PSDF_RSP, const float* one_minus_alpha, float* transmittance, float* bg_transmittance, void* stream);
int psdf_vr_integrate_with_weights(PSDF_RSP, const float* vals, const float* weights, float* out, void* stream);
int psdf_vr_sdf2alpha(PSDF_RSP, const float* ray_fixed_dt, const float* samples_dt, const float* sdf, float inv_s, int dynamic_inv_s,
                      float inv_s_multiplier, float* alpha, void* stream);
int psdf_vr_sum_over_each_ray(PSDF_RSP, int val_dim, const float* vals, float* sum_ray, float* sum_sample, void* stream);
int psdf_vr_cumsum_over_each_ray(PSDF_RSP, const float* vals, int inverse, float* out, void* stream);
int psdf_vr_compute_cdf(PSDF_RSP, const float* weights, float* cdf, void* stream);
int psdf_vr_importance_sample(PSDF_RSP, const float* origins, const float* dirs, const float* ray_fixed_dt, const float* samples_z,
                              const float* cdf, int nr_imp, uint64_t rng_state, uint64_t rng_inc, int jitter, float* o_pos,
                              float* o_dirs, float* o_z, void* stream);
long long psdf_vr_combine_workspace_bytes(int nr_rays);
/* one round of importance_sampling_sdf_model (permuto_sdf_py/utils/sdf_utils.py): sdf2alpha -> clip -> cumprod(1 - alpha + 1e-7) ->
 * weights -> per-ray normalisation (clamp 1e-6) -> cdf -> importance_sample, one launch, bit-identical to the separate calls.
 * cdf [N] is scratch + output (the per-ray exclusive cdf). */
int psdf_vr_importance_round(PSDF_RSP, const float* origins, const float* dirs, const float* ray_fixed_dt, const float* samples_dt,
                             const float* samples_z, const float* sdf, float inv_s, int dynamic_inv_s, float inv_s_multiplier, int nr_imp,
                             uint64_t rng_state, uint64_t rng_inc, int jitter, float* cdf, float* o_pos, float* o_dirs, float* o_z,
                             void* stream);
int psdf_vr_combine_uniform_samples_with_imp(PSDF_RSP, const float* origins, const float* dirs, const float* t_exit,
                                             const float* u_fixed_dt, const float* u_z, const float* u_sdf, int u_has_sdf,
                                             int imp_n, const float* i_z, const float* i_sdf, int i_has_sdf, int c_max,
                                             int* workspace, float* c_pos, float* c_dirs, float* c_z, float* c_dt, float* c_sdf,
                                             float* c_fixed_dt, int* c_start_end, void* stream);
int psdf_vr_cumprod_alpha2transmittance_backward(PSDF_RSP, const float* grad_bg_transmittance, const float* alpha,
                                                 const float* bg_transmittance, const float* cumsumLV, float* grad_alpha,
                                                 void* stream);
int psdf_vr_integrate_with_weights_backward(PSDF_RSP, const float* grad_pred, const float* vals, const float* weights,
                                            int reference_bug, float* grad_vals, float* grad_weights, void* stream);
int psdf_vr_sum_over_each_ray_backward(PSDF_RSP, int val_dim, const float* grad_sum_ray, const float* grad_sum_sample,
                                       float* grad_vals, void* stream);
int psdf_vr_compute_dt(PSDF_RSP, int use_t_exit, const float* t_exit, const float* samples_z, float* dt, void* stream);
int psdf_vr_volume_render_nerf(PSDF_RSP, const float* rgb, const float* radiance, const float* samples_z, const float* samples_dt,
                               float* pred_rgb, float* pred_depth, float* bg_transmittance, float* weight_per_sample, void* stream);
int psdf_vr_volume_render_nerf_backward(PSDF_RSP, const float* grad_pred_rgb, const float* grad_bg_transmittance,
                                        const float* pred_rgb, const float* bg_transmittance, const float* rgb,
                                        const float* radiance, const float* samples_dt, float* grad_rgb, float* grad_radiance,
                                        void* stream);

/* Fused NeuS compositing + per-ray losses (training): replaces VolumeRenderingNeus.compute_weights + integrate
 * (permuto_sdf_py/volume_rendering/volume_rendering_modules.py:129-176) and the rgb / mask / eikonal losses
 * (permuto_sdf_py/train_permuto_sdf.py:349-383, permuto_sdf_py/utils/permuto_sdf_utils.py:43-51) in one launch each way.
 * sdf [N], grad/rgb/dirs [N,3], dt [N], inv_s_dev [1] (device scalar, clipped to [1e-6,1e6] inside), gt_rgb [R,3],
 * gt_mask [R] or NULL, hit [R] u8 or NULL, bg_rgb [R,3] or NULL (pred += bg_T * bg_rgb).
 * ray_loss [R,3] = {sum_c |gt-pred| * hit, BCE(clip(w_sum,1e-3,1-1e-3), mask), sum_i (|grad_i|-1)^2}. */
int psdf_neus_render_loss_forward(PSDF_RSP, const float* sdf, const float* grad, const float* rgb, const float* dirs, const float* dt,
                                  const float* inv_s_dev, float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb,
                                  const float* gt_mask, const uint8_t* hit, const float* bg_rgb, float* alpha, float* transmittance, float* weights,
                                  float* pred_rgb, float* weights_sum, float* bg_transmittance, float* ray_loss, void* stream);
/* d loss / d {sdf, grad, rgb, bg_rgb, inv_s} for loss = g_total * (scale_rgb * sum rgb-term + scale_mask * sum bce + scale_eik * sum eik);
 * g_total_dev [1] device scalar or NULL (=1); g_bg_rgb / g_inv_s may be NULL; g_inv_s is accumulated (+=).
 * cos_anneal_dev [1] (both calls) overrides cos_anneal_ratio when not NULL; nr_samples_dev [1] int, when not NULL, divides
 * scale_eik on the device (static-capacity containers under CUDA-graph replay, where the sample count is not known on the host). */
int psdf_neus_render_loss_backward(PSDF_RSP, const float* sdf, const float* grad, const float* rgb, const float* dirs, const float* dt,
                                   const float* inv_s_dev, float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb,
                                   const float* gt_mask, const uint8_t* hit, const float* bg_rgb, const float* alpha, const float* transmittance,
                                   const float* pred_rgb, const float* weights_sum, const float* bg_transmittance,
                                   const float* g_total_dev, float scale_rgb, float scale_mask, float scale_eik,
                                   const int* nr_samples_dev, float* g_sdf,
                                   float* g_grad, float* g_rgb, float* g_bg_rgb, float* g_inv_s, void* stream);

/* ---------------------------------------------------------------- colour network helpers
 * LipshitzMLP weight normalisation (permuto_sdf_py/models/models.py:96-110): W_eff[r,:] = W[r,:] * min(1, softplus(c) / sum|W[r,:]|);
 * W [rows,cols], c [1] device scalar. Backward: grad_W (=), grad_c (+=, may be NULL). */
int psdf_lipschitz_normalize(int rows, int cols, const float* W, const float* c, float* W_eff, void* stream);
int psdf_lipschitz_normalize_backward(int rows, int cols, const float* W, const float* c, const float* grad_W_eff, float* grad_W,
                                      float* grad_c, void* stream);
/* Colorcal.calib_RGB_samples_packed + sigmoid (models.py:395-414, 677-741) on packed samples: out = sigmoid(x (1 + weight_delta[img]) +
 * bias[img]), identity calibration for img == fixed_img or img_idx == NULL (img_idx [R] per ray). Rows outside every ray are not
 * written. Backward: grad_x (=, rows inside rays), grad_weight_delta / grad_bias [nr_imgs,3] (+=, may be NULL). */
int psdf_calib_sigmoid_forward(PSDF_RSP, const float* x, const int* img_idx, const float* weight_delta, const float* bias, int fixed_img,
                               float* out, void* stream);
int psdf_calib_sigmoid_backward(PSDF_RSP, const float* x, const float* out, const float* grad_out, const int* img_idx,
                                const float* weight_delta, int fixed_img, float* grad_x, float* grad_weight_delta, float* grad_bias,
                                void* stream);

/* Curvature loss of SDF.get_sdf_and_curvature_1d_precomputed_gradient_normal_based (models.py:261-294): shifted sample positions
 * points + eps * cross(normalize(sdf_grad), normalize(rand_dirs)); curvature [n] = acos(clamp(n . n_shifted)) / pi (rows >=
 * nr_valid_dev[0] give 0; nr_valid_dev may be NULL) with loss_sum [1] += their sum; backward of  g_loss * scale * mean-or-sum:
 * the sum is divided by max(nr_valid, 1) when nr_valid_dev is given. */
int psdf_curvature_shift_points(int n, const float* points, const float* sdf_grad, const float* rand_dirs, float eps, float* out,
                                void* stream);
int psdf_curvature_loss_forward(int n, const float* sdf_grad, const float* sdf_grad_shifted, const int* nr_valid_dev, float* curvature,
                                float* loss_sum, void* stream);
int psdf_curvature_loss_backward(int n, const float* sdf_grad, const float* sdf_grad_shifted, const int* nr_valid_dev,
                                 const float* g_loss_dev, float scale, float* grad_sdf_grad, float* grad_sdf_grad_shifted, void* stream);

/* ---------------------------------------------------------------- PermutoSDF statics (include/permuto_sdf/PermutoSDF.cuh:46-55) */
int psdf_spherical_harmonics(int n, int degree, const float* dirs, float* out, void* stream);
int psdf_random_rays_from_reel(int nr_rays, int nr_images, int H, int W, const float* rgb_reel, const float* mask_reel,
                               const float* K, const float* tf_world_cam, const int* pixel_indices, const int* img_indices,
                               int has_mask, float* origins, float* dirs, float* gt_rgb, float* gt_mask, void* stream);

/* ---------------------------------------------------------------- permutohedral_encoding (external package; call sites
 * permuto_sdf_py/models/models.py:149,186; semantics SURVEY.md Appendix B)
 * pos [N,D] fp32, lattice [L,T,F] fp32, scale_factor [L,D], shift [L,D], window [L];
 * out [N,(L+E)*F] with E=ceil(D/F) concat columns when concat_points. D in {3,4}, F == 2. */
int psdf_enc_forward(int N, int D, int L, int F, int T, const float* pos, const float* lattice, const float* scale_factor,
                     const float* shift, const float* window, int concat_points, float points_scaling, float* out, void* stream);
/* grad wrt lattice (+=, caller zeroes) and/or positions (=) from grad_out [N,(L+E)*F]; either output may be NULL */
int psdf_enc_backward(int N, int D, int L, int F, int T, const float* pos, const float* lattice, const float* scale_factor,
                      const float* shift, const float* window, int concat_points, float points_scaling, const float* grad_out,
                      float* grad_lattice, float* grad_pos, void* stream);
/* double backward from the positions gradient: given gg_pos [N,D] (grad of grad_pos) produce
 * grad_lattice (+=) and grad_grad_out [N,(L+E)*F] (=); either may be NULL */
int psdf_enc_double_backward(int N, int D, int L, int F, int T, const float* pos, const float* lattice, const float* scale_factor,
                             const float* shift, const float* window, int concat_points, float points_scaling,
                             const float* gg_pos, const float* grad_out, float* grad_lattice, float* grad_grad_out, void* stream);

/* ---------------------------------------------------------------- fused encoding + SDF MLP (tcgen05)
 * Replaces SDF.forward / SDF.get_sdf_and_gradient of permuto_sdf_py/models/models.py:176-259 for evaluations that
 * need no parameter gradients: importance sampling (sdf_utils.py:388,401), occupancy refresh
 * (train_permuto_sdf.py:388-391), sphere tracing (sdf_utils.py:166,199-204).
 * MLP = Linear(in,h) GELU Linear(h,h) GELU Linear(h,h) GELU Linear(h,out), weights in torch.nn.Linear layout.
 * psdf_sdf_mlp_pack converts them once per weight update into the tensor-core operand blob. */
long long psdf_sdf_mlp_blob_bytes(int in_dim, int hidden, int out_dim);
int psdf_sdf_mlp_pack(int in_dim, int hidden, int out_dim, const float* W0, const float* b0, const float* W1, const float* b1,
                      const float* W2, const float* b2, const float* W3, const float* b3, uint8_t* blob, void* stream);
/* pos [N,3] -> sdf [N], grad [N,3] (NULL: value only), geom [N,out_dim-1] (NULL: skip). D = 3, F = 2, L % 4 == 0,
 * concat_points on (in_dim = 2L + 4). */
int psdf_sdf_fused_forward(int N, int L, int T, const float* pos, const float* lattice, const float* scale_factor, const float* shift,
                           const float* window, float points_scaling, int hidden, int out_dim, const uint8_t* blob, float* sdf,
                           float* grad, float* geom, void* stream);
/* Fused sphere tracing (sphere_trace, permuto_sdf_py/utils/sdf_utils.py:120-218): nr_iters iterations of
 *   sdf = SDF(p); p += dir * sdf * sdf_multiplier; converged |= |sdf| < tresh | left the occupied region / bounding sphere
 * with advance_sample_to_next_occupied_voxel applied after every step when an occupancy grid is given (occupancy != NULL; else
 * the bounding-sphere test). Persistent CTAs of 128 ray slots pull rays from a global queue and refill finished slots;
 * pos [N,3] start points, dirs [N,3]; pos_out [N,3], converged [N] (may be NULL); queue_counter: device int32 [8], ZERO on entry:
 * [0] scratch (ray queue); statistics of the launch: [1] network evaluations, [2] rounds summed over the CTAs, [3] rounds that ran the
 * network, [4] rounds of the longest CTA. Bit-identical per ray to the masked Python loop on the per-op kernels. */
int psdf_sdf_sphere_trace(int N, int L, int T, const float* pos, const float* dirs, const float* lattice, const float* scale_factor,
                          const float* shift, const float* window, float points_scaling, int hidden, int out_dim, const uint8_t* blob,
                          int nr_iters, float sdf_multiplier, float sdf_converged_tresh, const uint8_t* occupancy, int V, float extent,
                          const float trans[3], float sphere_radius, const float sphere_center[3], float* pos_out, uint8_t* converged,
                          int* queue_counter, void* stream);
/* Training backward of psdf_sdf_fused_forward (double backward of encoding + MLP on the tensor cores, ONE kernel, weight gradients
 * formed on chip):
 * upstream gradients g_sdf [N], g_grad [N,3], g_geom [N,out_dim-1] (any may be NULL) -> grad_lattice (+=), weight gradients
 * gW_l [N_l, K_l] (+=) and bias gradients gb_l [N_l] (+=). workspace: psdf_sdf_fused_backward_workspace_bytes(N) bytes of
 * scratch (the encoder operand tile of every 128-sample tile, written and re-read by the same CTA). Replaces loss.backward() through
 * SDF.get_sdf_and_gradient
 * (models.py:199-259). */
long long psdf_sdf_fused_backward_workspace_bytes(int N);
int psdf_sdf_fused_backward(int N, int L, int T, const float* pos, const float* lattice, const float* scale_factor, const float* shift,
                            const float* window, float points_scaling, int hidden, int out_dim, const uint8_t* blob, const float* g_sdf,
                            const float* g_grad, const float* g_geom, float* grad_lattice, uint8_t* workspace, float* gW0, float* gW1,
                            float* gW2, float* gW3, float* gb0, float* gb1, float* gb2, float* gb3, void* stream);
/* ---------------------------------------------------------------- fused colour network (RGB.forward, permuto_sdf_py/models/models.py:309-420)
 * encoding(points) | SH_5(dirs) | normalize(sdf_grad) | geom [N,32] -> 4-layer GELU MLP (in -> h1 -> h2 -> h3 -> 3) on the tensor
 * cores; out [N,3] is the linear output (colour calibration + sigmoid: psdf_calib_sigmoid_*). Weights are passed as a packed
 * operand blob (psdf_rgb_mlp_pack of the already Lipschitz-normalised matrices W_l [N_l,K_l] + biases); L % 4 == 0,
 * in_dim = 2L + 64 <= 128, h* <= 128 and % 16 == 0. */
long long psdf_rgb_mlp_blob_bytes(int in_dim, int h1, int h2, int h3, int out_dim);
int psdf_rgb_mlp_pack(int in_dim, int h1, int h2, int h3, int out_dim, const float* W0, const float* b0, const float* W1, const float* b1,
                      const float* W2, const float* b2, const float* W3, const float* b3, uint8_t* blob, void* stream);
int psdf_rgb_fused_forward(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                           const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                           int h1, int h2, int h3, const uint8_t* blob, float* out, void* stream);

/* backward of psdf_rgb_fused_forward: g_out [N,3] = d loss / d out -> grad_lattice (+=), g_sdf_grad [N,3] (=), g_geom [N,32] (=),
 * weight gradients gW_l [N_l,K_l] (+=, wrt the normalised matrices) and bias gradients gb_l (+=). workspace:
 * psdf_rgb_fused_backward_workspace_bytes(N) bytes of scratch. */
/* psdf_sdf_fused_backward over up to three independent sample sets in ONE launch (N_i = 0: unused set); every set has its own
 * positions and upstream gradients, the accumulated parameter gradients are shared. One launch of 1024 tiles loses 1 % to the last
 * partial wave on 148 SMs where two launches of 512 tiles lose 13 % each. */
long long psdf_sdf_fused_backward_multi_workspace_bytes(int N0, int N1, int N2);
int psdf_sdf_fused_backward_multi(int L, int T, const float* lattice, const float* scale_factor, const float* shift, const float* window,
                                  float points_scaling, int hidden, int out_dim, const uint8_t* blob, int N0, const float* pos0,
                                  const float* g_sdf0, const float* g_grad0, const float* g_geom0, int N1, const float* pos1,
                                  const float* g_sdf1, const float* g_grad1, const float* g_geom1, int N2, const float* pos2,
                                  const float* g_sdf2, const float* g_grad2, const float* g_geom2, float* grad_lattice, uint8_t* workspace,
                                  float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1, float* gb2, float* gb3, void* stream);
long long psdf_rgb_fused_backward_workspace_bytes(int N);
int psdf_rgb_fused_backward(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                            const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                            int h1, int h2, int h3, const uint8_t* blob, const float* g_out, float* grad_lattice, float* g_sdf_grad,
                            float* g_geom, uint8_t* workspace, float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1,
                            float* gb2, float* gb3, void* stream);

/* ---------------------------------------------------------------- direct (autograd-free) training iteration
 * The reference's iteration (permuto_sdf_py/train_permuto_sdf.py:311-422) spends ~70 one-block PyTorch launches between its big kernels
 * (schedule ramps, Lipschitz normalisation, calibration, loss reductions, zero-fills, gradient adds). These entry points fold them into
 * the neighbouring kernels; permuto_sdf_b200/train.py Trainer._iteration_direct strings them together. */
/* kernel variant of the value + tangent forward (A/B measurements): 1 = two independent 64-sample groups per CTA (default),
 * 0 = lock-step 128-sample tiles; any other value only queries. Returns the variant in use. Same results to the last bit. */
int psdf_sdf_forward_variant(int variant);
/* two independent sample sets in one psdf_sdf_fused_forward launch (N1 = 0: unused); grad / geom may be NULL per set */
int psdf_sdf_fused_forward_multi(int L, int T, const float* lattice, const float* scale_factor, const float* shift, const float* window,
                                 float points_scaling, int hidden, int out_dim, const uint8_t* blob, int N0, const float* pos0, float* sdf0,
                                 float* grad0, float* geom0, int N1, const float* pos1, float* sdf1, float* grad1, float* geom1, void* stream);
/* psdf_sdf_mlp_pack + the end-of-iteration counters in the same launch: step_dev [1] (device int, AdamW step count) += 1 and
 * it_dev [1] (device float, iteration number) += 1; either may be NULL */
int psdf_sdf_mlp_pack_advance(int in_dim, int hidden, int out_dim, const float* W0, const float* b0, const float* W1, const float* b1,
                              const float* W2, const float* b2, const float* W3, const float* b3, uint8_t* blob, int* step_dev, float* it_dev,
                              void* stream);
/* psdf_rgb_fused_backward with g_sdf_grad accumulated (+=) instead of written */
int psdf_rgb_fused_backward_acc(int N, int L, int T, const float* pos, const float* dirs, const float* sdf_grad, const float* geom, int geom_dim,
                                const float* lattice, const float* scale_factor, const float* shift, const float* window, float points_scaling,
                                int h1, int h2, int h3, const uint8_t* blob, const float* g_out, float* grad_lattice, float* g_sdf_grad,
                                float* g_geom, uint8_t* workspace, float* gW0, float* gW1, float* gW2, float* gW3, float* gb0, float* gb1,
                                float* gb2, float* gb3, void* stream);
/* psdf_neus_render_loss_forward / _backward with the colour head (Colorcal.calib_RGB_samples_packed + sigmoid, models.py:395-414,677-741)
 * and the curvature term (models.py:283-294) folded in: x_raw [N,3] is the colour network's linear output; the backward returns
 * g_sdf [n_rows], g_grad [n_rows,3] (compositing + eikonal + curvature), g_x [n_rows,3], g_grad_shifted [n_rows,3] (curvature term wrt
 * the gradients at the shifted points; grad_shifted NULL: no curvature term), accumulates the calibration gradients (+=, may be NULL)
 * and zero-fills rows [nr_valid_dev[0], n_rows) (nr_valid_dev NULL: no tail). curv_scale weighs the curvature MEAN over
 * nr_samples_dev[0] (or max_nr_samples) rows; scale_eik is divided by the same count. */
int psdf_neus_head_loss_forward(PSDF_RSP, const float* sdf, const float* grad, const float* x_raw, const float* dirs, const float* dt,
                                const float* inv_s_dev, float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb,
                                const float* gt_mask, const uint8_t* hit, const int* img_idx, const float* weight_delta, const float* bias,
                                int fixed_img, float* alpha, float* transmittance, float* weights, float* pred_rgb, float* weights_sum,
                                float* bg_transmittance, float* ray_loss, void* stream);
int psdf_neus_head_loss_backward(PSDF_RSP, const float* sdf, const float* grad, const float* x_raw, const float* dirs, const float* dt,
                                 const float* inv_s_dev, float cos_anneal_ratio, const float* cos_anneal_dev, const float* gt_rgb,
                                 const float* gt_mask, const uint8_t* hit, const int* img_idx, const float* weight_delta, const float* bias,
                                 int fixed_img, const float* alpha, const float* transmittance, const float* pred_rgb, const float* weights_sum,
                                 const float* bg_transmittance, float scale_rgb, float scale_mask, float scale_eik, const int* nr_samples_dev,
                                 const float* grad_shifted, float curv_scale, const float* curv_scale_dev, int n_rows, const int* nr_valid_dev,
                                 float* g_sdf, float* g_grad, float* g_x, float* g_grad_shifted, float* g_weight_delta, float* g_bias,
                                 void* stream);
/* schedule ramps from the device-resident iteration (map_range_val, permuto_sdf_py/utils/common_utils.py:156-160): out[i] = out0 +
 * (out1 - out0) / (in1 - in0) * (clamp(it, in0, in1) - in0); kinds[i] = 1 / 2: exp(10 * ramp) unclipped / clipped to [1e-6, 1e6]
 * (inv_s of SingleVarianceNetwork with a forced variance). params [n,4] = {in0, in1, out0, out1} and kinds [n] are HOST arrays, n <= 8;
 * it_dev [1] device float or NULL (then it_host). */
int psdf_iter_scalars(int n, const float* params, const int* kinds, const float* it_dev, float it_host, float* out, void* stream);
/* psdf_lipschitz_normalize of the four colour-MLP matrices + psdf_rgb_mlp_pack of the result in one launch */
int psdf_lipschitz_pack4(int in_dim, int h1, int h2, int h3, int out_dim, const float* W0, const float* b0, const float* c0, const float* W1,
                         const float* b1, const float* c1, const float* W2, const float* b2, const float* c2, const float* W3, const float* b3,
                         const float* c3, uint8_t* blob, void* stream);
/* psdf_lipschitz_normalize_backward of the four layers in one launch: G_l = d loss / d W_eff_l is consumed and reset to zero,
 * grad_W_l (+=), grad_c_l (+=); lip_weight != 0 adds d (lip_weight * prod_l softplus(c_l)) / d c_l (LipshitzMLP.lipshitz_bound_full) */
int psdf_lipschitz_backward4(int in_dim, int h1, int h2, int h3, int out_dim, const float* W0, const float* c0, float* G0, float* gW0, float* gc0,
                             const float* W1, const float* c1, float* G1, float* gW1, float* gc1, const float* W2, const float* c2, float* G2,
                             float* gW2, float* gc2, const float* W3, const float* c3, float* G3, float* gW3, float* gc3, float lip_weight,
                             void* stream);
/* every scalar loss term of the iteration + the weighted total + the off-surface gradient seed in one launch (see csrc/iter_glue.cu);
 * acc [8] scratch: zero before the first call, left zero by every call; terms [12] = {sum L1, sum BCE, sum eikonal, mean curvature,
 * mean off-surface, Lipschitz bound, valid rows, divisor of the per-sample means, loss_rgb, loss_eikonal, unused, unused} */
int psdf_loss_terms(int N, const int* nr_valid_dev, const int* nr_mean_dev, const float* grad, const float* grad_shifted, int R,
                    const float* ray_loss, int n_off, const float* sdf_off, float* g_off, float c_rgb, float c_mask, float w_eik, float w_curv,
                    const float* w_curv_dev, float w_off, float w_lip, const float* lip_c0, const float* lip_c1, const float* lip_c2,
                    const float* lip_c3, float* acc, float* loss, float* terms, void* stream);
/* Sphere.rand_points_inside (src/Sphere.cu) from ONE uniform draw u01 [3,n]: rows = phi / (2 pi), (cos(theta) + 1) / 2, u */
int psdf_sphere_rand_points_inside_u01(int n, float radius, const float* u01, float* out, void* stream);
/* psdf_adamw_step over up to 8 parameter groups in one launch: device-resident step count, per-group hyper_dev [2] = {lr, weight_decay};
 * gradients scaled by grad_scale and reset to zero. n_l % 4 == 0, 16-byte aligned pointers. n and the five tables of device addresses
 * (one entry per group) are HOST arrays. The step used is step_dev[0] + step_offset (step_offset = 1: the counter is advanced afterwards
 * by psdf_sdf_mlp_pack_advance). leave_room != 0: few resident blocks with more loads in flight each, so that kernels of another
 * stream (the next iteration's occupancy sampling) can be co-resident with the sweep. */
int psdf_adamw_multi_step(int n_groups, const long long* n, const uint64_t* param, const uint64_t* grad, const uint64_t* exp_avg,
                          const uint64_t* exp_avg_sq, const uint64_t* hyper_dev, float beta1, float beta2, float eps, const int* step_dev,
                          int step_offset, float grad_scale, int leave_room, void* stream);

/* ---------------------------------------------------------------- dense fused AdamW (torch.optim.AdamW / apex FusedAdam math,
 * train_permuto_sdf.py:293-304); step >= 1 is the incremented step count; grad is multiplied by grad_scale and, when
 * zero_grad != 0, reset to zero in the same pass. Pointers 16-byte aligned. step_dev [1] (device int32), when not NULL,
 * replaces `step` so that the call can be replayed from a CUDA graph; hyper_dev [2] (device float: lr, weight_decay), when not
 * NULL, replaces `lr` / `weight_decay` the same way (LR schedulers and the wd = 1.0 switch of train_permuto_sdf.py:400-403 edit
 * param_groups between replays of one captured graph). */
int psdf_adamw_step(long long n, float* param, float* grad, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, const int* step_dev, const float* hyper_dev, float grad_scale, int zero_grad,
                    void* stream);

/* Data-parallel step without NCCL: gradient reduction + AdamW + parameter broadcast in ONE kernel over NVLink peer memory. This rank
 * owns the elements [shard_lo, shard_lo + n) of the group (both multiples of 4): it sums that shard of ALL ranks' gradient buffers
 * (peer loads), updates its local moments and writes the new parameters into ALL ranks' parameter buffers (peer stores).
 * grad_ptrs / param_ptrs: HOST arrays of `world` <= 8 device pointers (rank r's buffer of this group, peer-mapped here, e.g.
 * torch.distributed._symmetric_memory); mc_grad_ptr / mc_param_ptr: the NVSwitch MULTICAST mapping of the same two buffers (group base),
 * or 0: when given, the reduction is one multimem.ld_reduce (summed in the switch) and the broadcast one multimem.st per 16 bytes;
 * exp_avg / exp_avg_sq: local, full group size. world in {1, 2, 4, 8}. The caller puts a cross-rank barrier before
 * (gradients final) and after (parameters landed) the call and zeroes its own gradients afterwards. Replaces all-reduce + AdamW of
 * the reference's (non-existent) data-parallel loop: SURVEY.md 8(e). */
int psdf_adamw_dp_step(long long n, long long shard_lo, int world, int rank, const uint64_t* grad_ptrs, const uint64_t* param_ptrs,
                       uint64_t mc_grad_ptr, uint64_t mc_param_ptr, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, const int* step_dev, const float* hyper_dev, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
