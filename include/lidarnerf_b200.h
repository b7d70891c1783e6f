/*
 * lidarnerf_b200.h - C ABI of the B200-native LiDAR-NeRF volume-rendering hot path.
 *
 * One entry point per function the reference binds through pybind11 for this path
 * (boundary B1).  Every entry point takes plain device pointers and
 * sizes - no torch types - plus the CUDA stream to launch on, is asynchronous (no host
 * sync), allocates nothing and returns an int status:
 *      0                       success
 *      > 0                     a cudaError_t raised by the launch
 *      LNB_ERR_* (negative)    argument / configuration rejected before any launch
 * All outputs are pre-allocated by the caller and written in place, exactly like the
 * reference's `_backend.*` functions.  `lnb_strerror` turns a status into text.
 *
 * Citations are relative to the reference checkout (tangtaogo/lidar-nerf).
 */
#ifndef LIDARNERF_B200_H_
#define LIDARNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *lnb_stream_t; /* a cudaStream_t; NULL = legacy default stream */

#define LNB_OK 0
#define LNB_ERR_INVALID_ARGUMENT (-1) /* NULL pointer, zero size where not allowed, bad enum */
#define LNB_ERR_UNSUPPORTED (-2)      /* shape / dtype outside what this build implements     */
#define LNB_ERR_NO_DEVICE (-3)        /* no sm_100 device / CUDA runtime unavailable           */
#define LNB_ERR_WORKSPACE (-4)        /* caller-provided workspace too small                    */

/* element types for the gridencoder tables (reference dispatches on embeddings.scalar_type()) */
#define LNB_F32 0
#define LNB_F16 1

/* memory layout of the per-level feature tensor */
#define LNB_LAYOUT_LBC 0 /* [L, B, C]   - what the reference kernels read/write (gridencoder.cu:117,287) */
#define LNB_LAYOUT_BLC 1 /* [B, L * C]  - what the MLP consumes; saves the torch permute (grid.py:87,104)   */

const char *lnb_strerror(int status);
/* library version (major*10000 + minor*100 + patch) and the arch it was compiled for ("sm_100a") */
int lnb_version(void);
const char *lnb_arch(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t lnb_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * raymarching  (replaces lidarnerf/raymarching/src/raymarching.h:6-96, bindings.cpp:5-21)
 * All floating tensors are fp32 (the reference wrappers cast to float32:
 * raymarching.py:17,53,138,173,294,369,465).
 * ---------------------------------------------------------------------------------------- */

/* raymarching.cu:159-177  rays_o/rays_d [N,3], aabb [6] -> nears/fars [N] */
int lnb_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N,
                           float min_near, float *nears, float *fars, lnb_stream_t stream);

/* raymarching.cu:219-233  -> coords [N,2] in [-1,1] */
int lnb_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N,
                     float *coords, lnb_stream_t stream);

/* raymarching.cu:249-253 / 274-280  coords [N,3] int32 <-> indices [N] int32 */
int lnb_morton3D(const int32_t *coords, uint32_t N, int32_t *indices, lnb_stream_t stream);
int lnb_morton3D_invert(const int32_t *indices, uint32_t N, int32_t *coords, lnb_stream_t stream);

/* raymarching.cu:308-320  grid [N*8] fp32 -> bitfield [N] u8, bit i = grid[8n+i] > thresh */
int lnb_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield,
                 lnb_stream_t stream);
/* packbits with the threshold min(*mean_density_dev, density_thresh_cap) read on the device (the occupancy-grid rule
 * density_thresh = min(mean_density, density_thresh) of SURVEY.md Appendix A without a host round trip) */
int lnb_packbits_dev(const float *grid, uint32_t N, const float *mean_density_dev, float density_thresh_cap,
                     uint8_t *bitfield, lnb_stream_t stream);

/* raymarching.cu:536-568.  counter[0] += samples, counter[1] += rays (device atomics);
 * rays [N,3] = (ray id, sample offset, sample count) in arrival order; samples of a ray whose
 * span would exceed M are not written (raymarching.cu:456-457). */
int lnb_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                         float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                         uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                         float *deltas, int32_t *rays, int32_t *counter, const float *noises,
                         lnb_stream_t stream);

/* Extension: additionally writes ray_ids [M] (nullable) = rays_o/rays_d row each sample belongs to, and accepts
 * dirs == NULL when ray_ids is given (all samples of a ray share its direction; the fused field kernels below look
 * the direction terms up per ray instead of reading 12 B per sample).  It also zeroes the rows of xyzs / dirs /
 * deltas / ray_ids between the produced total counter[0] and the next multiple of 128 (capped at M): the padding of
 * the last 128-row tile that the per-sample kernels process (what lnb_zero_sample_tail_ex does as a separate launch). */
int lnb_march_rays_train_ex(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                            float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                            uint32_t M, const float *nears, const float *fars, float *xyzs, float *dirs,
                            float *deltas, int32_t *rays, int32_t *counter, const float *noises,
                            int32_t *ray_ids, lnb_stream_t stream);

/* raymarching.cu:657-678.  rgbs [M,3]; outputs indexed by rays[n,0]. */
int lnb_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas,
                                     const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                     float *weights_sum, float *depth, float *image,
                                     lnb_stream_t stream);

/* raymarching.cu:774-802.  grad_sigmas/grad_rgbs must be zero-filled by the caller
 * (raymarching.py:338-339); there is no depth gradient (raymarching.py:329-330). */
int lnb_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image,
                                      const float *sigmas, const float *rgbs, const float *deltas,
                                      const int32_t *rays, const float *weights_sum,
                                      const float *image, uint32_t M, uint32_t N, float T_thresh,
                                      float *grad_sigmas, float *grad_rgbs, lnb_stream_t stream);

/* Extension used by this repo's run_cuda glue: `channels` in {1,2,3,4}
 * (the LiDAR head emits 2: ray-drop, intensity); optional depth gradient (`grad_depth` and
 * `depth` may both be NULL = reference behaviour).  With channels == 3 and grad_depth == NULL
 * these are the two functions above. */
int lnb_composite_rays_train_forward_ex(const float *sigmas, const float *rgbs, const float *deltas,
                                        const int32_t *rays, uint32_t M, uint32_t N, float T_thresh,
                                        uint32_t channels, float *weights_sum, float *depth,
                                        float *image, lnb_stream_t stream);
int lnb_composite_rays_train_backward_ex(const float *grad_weights_sum, const float *grad_depth,
                                         const float *grad_image, const float *sigmas,
                                         const float *rgbs, const float *deltas, const int32_t *rays,
                                         const float *weights_sum, const float *depth,
                                         const float *image, uint32_t M, uint32_t N, float T_thresh,
                                         uint32_t channels, float *grad_sigmas, float *grad_rgbs,
                                         lnb_stream_t stream);

/* raymarching.cu:930-964 (inference march; xyzs/dirs/deltas zero-filled by the caller) */
int lnb_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t *rays_alive, const float *rays_t,
                   const float *rays_o, const float *rays_d, float bound, float dt_gamma,
                   uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t *grid,
                   const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                   const float *noises, lnb_stream_t stream);

/* raymarching.cu:1055-1077 (inference composite, in place; rays_alive[n] = -1 on termination) */
int lnb_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t *rays_alive,
                       float *rays_t, const float *sigmas, const float *rgbs, const float *deltas,
                       float *weights_sum, float *depth, float *image, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * gridencoder  (replaces lidarnerf/gridencoder/src/gridencoder.h:12-55, bindings.cpp:5-11)
 * inputs are always fp32 in [0,1] (gridencoder.cu:96,629); `dtype` is the type of embeddings,
 * outputs, dy_dx, grad, grad_embeddings and grad_inputs (LNB_F32 / LNB_F16).
 * D in {2,3}; C in {1,2,4,8}.  S = log2(per_level_scale); H = base resolution.
 * ---------------------------------------------------------------------------------------- */

/* gridencoder.cu:594-637.  dy_dx [B, L*D*C] may be NULL.  layout selects the outputs layout. */
int lnb_grid_encode_forward(const float *inputs, const void *embeddings, const int32_t *offsets,
                            void *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                            uint32_t H, void *dy_dx, uint32_t gridtype, int align_corners,
                            uint32_t interp, int dtype, int layout, lnb_stream_t stream);

/* gridencoder.cu:639-693.  grad_embeddings (and grad_inputs when dy_dx != NULL) must be
 * zero-filled by the caller (grid.py:106-109); gradients are accumulated with atomics. */
int lnb_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings,
                             const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                             uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx,
                             void *grad_inputs, uint32_t gridtype, int align_corners,
                             uint32_t interp, int dtype, int layout, lnb_stream_t stream);
/* _gridencoder.grad_total_variation (gridencoder/src/bindings.cpp:10, gridencoder.cu:695-911): adds the gradient of a
 * total-variation penalty over the cells that contain `inputs` [B,D] (in [0,1], SAME dtype as the table - the
 * reference reinterprets the inputs as scalar_t) to `grad` (table layout, table dtype).  No caller in the reference. */
int lnb_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int32_t *offsets,
                             float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             uint32_t gridtype, int align_corners, int dtype, lnb_stream_t stream);

/* Extensions used by the fused training step:
 *  - `in_bound` > 0: inputs are world coordinates in [-bound, bound], mapped to [0,1] inside the kernel exactly
 *    as GridEncoder.forward does, (x + bound) / (2 bound) (grid.py:213); 0 = inputs already in [0,1];
 *  - `accumulate_f32` != 0 (backward, dtype == LNB_F16): the incoming gradient is fp16 but grad_embeddings is an
 *    fp32 table accumulated with fp32 atomics (no fp16 rounding of small updates); dy_dx must be NULL;
 *  - `n_active` (device pointer, nullable; [B, L*C] layout only): only the first round_up(*n_active, 128) rows are
 *    processed - the sample counter written by lnb_march_rays_train, read on the device (no host sync). */
int lnb_grid_encode_forward_ex(const float *inputs, const void *embeddings, const int32_t *offsets,
                               void *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                               uint32_t H, void *dy_dx, uint32_t gridtype, int align_corners,
                               uint32_t interp, int dtype, int layout, float in_bound,
                               const int32_t *n_active, lnb_stream_t stream);
int lnb_grid_encode_backward_ex(const void *grad, const float *inputs, const void *embeddings,
                                const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, const void *dy_dx,
                                void *grad_inputs, uint32_t gridtype, int align_corners,
                                uint32_t interp, int dtype, int layout, float in_bound,
                                int accumulate_f32, const int32_t *n_active, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * freqencoder  (replaces lidarnerf/freqencoder/src/freqencoder.h, bindings.cpp:5-9)   fp32
 * ---------------------------------------------------------------------------------------- */
int lnb_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                            float *outputs, lnb_stream_t stream);
int lnb_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D,
                             uint32_t deg, uint32_t C, float *grad_inputs, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * shencoder  (replaces lidarnerf/shencoder/src/shencoder.h, bindings.cpp:5-8)   fp32, D == 3,
 * C = degree in [1,8]; outputs [B, C*C]; dy_dx [B, 3*C*C] may be NULL.
 * ---------------------------------------------------------------------------------------- */
int lnb_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C,
                          float *dy_dx, lnb_stream_t stream);
int lnb_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                           const float *dy_dx, float *grad_inputs, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * ffmlp  (replaces lidarnerf/ffmlp/src/ffmlp.h:7-47, bindings.cpp:5-10)
 * fp16 storage, bias-free:  h0 = act(x W_in^T), h_k = act(h_{k-1} W_k^T), y = h_last W_out^T.
 * weights = [hidden*in | (num_layers-1)*hidden*hidden | out*hidden] row-major fp16
 * (ffmlp.cu:861-864).  B must be a multiple of 128 (ffmlp.py:254-262 pads).  This build
 * implements hidden_dim == 64, input_dim % 16 == 0 (<= 128), output_dim == 16, activation ReLU
 * (0) and output activation None (6); anything else returns LNB_ERR_UNSUPPORTED.  (The
 * reference's hidden_dim 16 / 32 are served one level up, lidar-nerf_b200/backend.py: the host
 * pads the weights to 64 hidden units with zeros - exact zeros in every accumulation - and calls
 * these entry points with hidden_dim = 64.)
 * Accumulation is fp32 in tensor memory (a superset of the reference's fp16 accumulators).
 * forward_buffer / backward_buffer: [num_layers, B, hidden] fp16 (post-activation / d(pre-act)).
 * ---------------------------------------------------------------------------------------- */
int lnb_ffmlp_forward(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                      uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                      uint32_t activation, uint32_t output_activation, void *forward_buffer,
                      void *outputs, lnb_stream_t stream);
int lnb_ffmlp_inference(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                        uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                        uint32_t activation, uint32_t output_activation, void *inference_buffer,
                        void *outputs, lnb_stream_t stream);
/* grad_weights (fp16, zero-filled by the caller, ffmlp.py:124) receives the weight gradient;
 * grad_inputs [B,in] is written when calc_grad_inputs != 0.  `workspace` must hold
 * lnb_ffmlp_backward_workspace_bytes(...) bytes of device memory (fp32 weight-gradient
 * accumulators); it is zeroed and consumed inside the call. */
size_t lnb_ffmlp_backward_workspace_bytes(uint32_t input_dim, uint32_t output_dim,
                                          uint32_t hidden_dim, uint32_t num_layers);
int lnb_ffmlp_backward(const void *grad, const void *inputs, const void *weights,
                       const void *forward_buffer, uint32_t B, uint32_t input_dim,
                       uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                       uint32_t activation, uint32_t output_activation, int calc_grad_inputs,
                       void *backward_buffer, void *grad_inputs, void *grad_weights,
                       void *workspace, size_t workspace_bytes, lnb_stream_t stream);
/* Same backward, but the weight gradient is ADDED (fp32 atomics) into `grad_weights_f32` (flat weight layout,
 * not zeroed, not converted) - the form the fused training step feeds straight to lnb_adam_step. */
int lnb_ffmlp_backward_accumulate(const void *grad, const void *inputs, const void *weights,
                                  const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                  uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                  uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                  float *grad_weights_f32, const int32_t *n_active, lnb_stream_t stream);
/* forward with the device-side row count (see lnb_grid_encode_forward_ex); forward_buffer == NULL = inference */
int lnb_ffmlp_forward_ex(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                         uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                         uint32_t output_activation, void *forward_buffer, void *outputs, const int32_t *n_active,
                         lnb_stream_t stream);
/* ffmlp.cu:1030-1049 create/destroy split-K side streams.  This implementation accumulates the
 * weight gradient in tensor memory inside the backward kernel and needs no side streams; the
 * two symbols are kept so the reference's FFMLP.__init__ (ffmlp.py:230) binds unchanged. */
int lnb_allocate_splitk(size_t size);
int lnb_free_splitk(void);

/* ------------------------------------------------------------------------------------------
 * Training-step helpers that the reference leaves to torch.
 * ---------------------------------------------------------------------------------------- */

/* Fused Adam over a flat fp32 parameter vector with fp32 gradients:
 *   g = grad * grad_scale;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
 *   p -= lr * (m / bc1) / (sqrt(v / bc2) + eps)          (torch.optim.Adam semantics)
 * If `params_half` != NULL the updated parameters are also written as fp16 (the table the
 * encoders read); if `zero_grad` != 0 the gradient is cleared in the same pass. */
int lnb_adam_step(float *params, float *grad, float *exp_avg, float *exp_avg_sq, void *params_half,
                  size_t n, float lr, float beta1, float beta2, float eps, float bias_correction1,
                  float bias_correction2, float grad_scale, int zero_grad, lnb_stream_t stream);
/* The same update with the step-dependent scalars read from device memory, so the launch can live inside a CUDA graph
 * (by-value arguments of a captured launch are frozen; the bias corrections and a scheduled learning rate change every
 * step).  hyper_dev [5] floats is written by lnb_adam_set_hyper (a one-thread kernel on the same stream, launched
 * before each replay); enable == 0 turns the captured update into a no-op (nothing pending). */
int lnb_adam_set_hyper(float *hyper_dev, float lr, float bias_correction1, float bias_correction2, float grad_scale,
                       int enable, lnb_stream_t stream);
/* Data-parallel exchange fused with the optimiser over NVLink peer memory (SURVEY.md 8e "better variant"): this rank
 * sums shard [shard_lo, shard_lo + shard_n) of the flat fp32 gradient straight out of the `world` gradient buffers
 * grad_ptrs[q] (peer-mapped device pointers, rank order), runs Adam on its shard (params/exp_avg/exp_avg_sq hold only
 * the shard) and stores the updated fp16 parameters into every rank's shadow half_ptrs[q].  grad_ptrs / half_ptrs are
 * HOST arrays of `world` <= 16 pointers.  The caller provides the cross-rank barriers: every rank's gradient complete
 * before the launch, every rank's launch complete before the shadows are read. */
int lnb_dp_adam_exchange(const void *const *grad_ptrs, void *const *half_ptrs, uint32_t world, float *params_shard,
                         float *exp_avg_shard, float *exp_avg_sq_shard, size_t shard_lo, size_t shard_n, float lr,
                         float beta1, float beta2, float eps, float bias_correction1, float bias_correction2,
                         float grad_scale, lnb_stream_t stream);
/* The same exchange through NVSwitch multicast / in-switch reduction (NVLS): `grad_multicast` / `half_multicast` are the
 * MULTICAST addresses of the flat fp32 gradient / fp16 shadow (torch symmetric memory `multicast_ptr`); the gradient sum
 * of this rank's shard comes back from one multimem.ld_reduce per 16 bytes, the updated parameters reach every rank
 * with one multimem.st.  shard_lo and shard_n must be multiples of 8. */
int lnb_dp_adam_exchange_mc(const void *grad_multicast, void *half_multicast, float *params_shard, float *exp_avg_shard,
                            float *exp_avg_sq_shard, size_t shard_lo, size_t shard_n, float lr, float beta1, float beta2,
                            float eps, float bias_correction1, float bias_correction2, float grad_scale,
                            lnb_stream_t stream);
int lnb_adam_step_dev(float *params, float *grad, float *exp_avg, float *exp_avg_sq, void *params_half, size_t n,
                      float beta1, float beta2, float eps, const float *hyper_dev, int zero_grad, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Glue of the LiDAR field step - the elementwise torch code between the reference's extension calls
 * (nerf/network.py:162-237, nerf/utils.py:707-734, dataset/base_dataset.py:85-100), one pass each.
 * ---------------------------------------------------------------------------------------- */

/* The per-sample kernels below take `n_active` (nullable device pointer to the march's sample counter): rows at or
 * beyond round_up(*n_active, 128) are skipped.
 * zero xyzs/dirs/deltas rows [counter[0], round_up(counter[0], 128)) on the device: the padding inside the last
 * partially filled tile (raymarching.py:235-237 zero-fills whole buffers on the host side every call) */
int lnb_zero_sample_tail(float *xyzs, float *dirs, float *deltas, const int32_t *counter, uint32_t M,
                         lnb_stream_t stream);
/* same with nullable dirs and an optional ray_ids [M] tail (set to ray 0) */
int lnb_zero_sample_tail_ex(float *xyzs, float *dirs, float *deltas, int32_t *ray_ids, const int32_t *counter,
                            uint32_t M, lnb_stream_t stream);
/* sigma_out [M,16] fp16 (density head output), dirs [M,3] ->
 *   sigma [M] fp32 = exp(h0) * density_scale ; head_in [M,in_pad] fp16 = [freq_enc(dir,degree) | geo_feat(15) | 0] */
int lnb_field_head_input(const void *sigma_out, const float *dirs, uint32_t M, uint32_t degree, uint32_t in_pad,
                         float density_scale, float *sigma, void *head_in, const int32_t *n_active,
                         lnb_stream_t stream);
/* head_out [M,16] fp16 -> rgb [M,2] fp32 = sigmoid(h[0:2])  (ray-drop, intensity) */
int lnb_field_head_rgb(const void *head_out, uint32_t M, float *rgb, const int32_t *n_active, lnb_stream_t stream);
/* LiDAR loss (nerf/utils.py:726-734, mean over rays) and its gradient w.r.t. (weights_sum, depth, image[N,2]);
 * gt [N,3] = (ray-drop, intensity, depth); t0 [N] (nullable) = march start added as t0*weights_sum to depth;
 * loss_out[0] += loss; gradients are multiplied by loss_scale. */
int lnb_lidar_loss(const float *weights_sum, const float *depth, const float *image, const float *gt,
                   const float *t0, uint32_t N, float alpha_d, float alpha_r, float alpha_i, float loss_scale,
                   float *g_weights_sum, float *g_depth, float *g_image, float *loss_out, lnb_stream_t stream);
/* One kernel for lnb_composite_rays_train_forward_ex (2 channels) -> lnb_lidar_loss -> lnb_composite_rays_train_
 * backward_ex with the depth gradient: the warp that composites a ray keeps its (weights_sum, depth, image) in
 * registers, evaluates the loss terms of that ray and walks the ray's samples again for the backward pass.
 * Replaces raymarching.cu:578-802 + nerf/utils.py:726-734 in the fused training step.  Every sample of every
 * marched ray receives a gradient (zero behind the early stop) and the rows between counter[0] and the next
 * multiple of 128 are zeroed, so grad_sigmas / grad_rgbs need no zero fill.  The march start of each ray,
 * near + clamp(near * dt_gamma, dt_min, dt_max) * noise (raymarching.cu:369-375), is recomputed from
 * (nears, noises, dt_gamma, max_steps, C, H) and written to t0 [N] (nullable).  loss_out[0] += loss.
 * live_idx [M] / n_live [1] (both nullable): the rows that can carry a gradient - each ray's samples up to and
 * including the one that drove T below T_thresh - are appended to live_idx in arrival order and counted in n_live
 * (the caller zeroes n_live); the *_rows backward entry points below walk that list instead of all marched rows
 * (25-30 % of the marched samples of a LiDAR scene sit behind the first surface and have an exactly-zero gradient). */
int lnb_lidar_composite_step(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                              const float *gt, const float *nears, const float *noises, float dt_gamma,
                              uint32_t max_steps, uint32_t C, uint32_t H, const int32_t *counter, uint32_t M,
                              uint32_t N, float T_thresh, float alpha_d, float alpha_r, float alpha_i,
                              float loss_scale, float *weights_sum, float *depth, float *image, float *t0,
                              float *grad_sigmas, float *grad_rgbs, float *loss_out, int32_t *live_idx,
                              int32_t *n_live, lnb_stream_t stream);
/* lnb_lidar_loss + the patch depth-gradient term of the KITTI-360 configurations (nerf/utils.py:748-876 with
 * grad_loss = True, sobel_grad = False, depth_grad_loss = l1): rays arrive as patches [N / (px*py), px, py] of the range
 * image (base_dataset.py:52-74, change_patch_size_lidar = [2, 8]) and the loss adds
 *   alpha_grad * mean_{pairs (i,j),(i,j+1)} | |P - P'| mask - (G - G') mask |,  P = D m * inv_scale, G = d_gt m * inv_scale,
 *   mask = m * [ |G - G'| < grad_clip ]  (the reference's 0.01), inv_scale = 1 / opt.scale.
 * patch_x = patch_y = 1 or alpha_grad = 0 reduce it to lnb_lidar_loss. */
int lnb_lidar_loss_ex(const float *weights_sum, const float *depth, const float *image, const float *gt,
                      const float *t0, uint32_t N, float alpha_d, float alpha_r, float alpha_i, float loss_scale,
                      uint32_t patch_x, uint32_t patch_y, float alpha_grad, float inv_scale, float grad_clip,
                      float *g_weights_sum, float *g_depth, float *g_image, float *loss_out, lnb_stream_t stream);
/* lnb_lidar_composite_step split at the loss, for losses that couple neighbouring rays (lnb_lidar_loss_ex):
 *   lnb_lidar_composite_forward  -> weights_sum / depth / image / t0 of every ray;
 *   lnb_lidar_composite_backward <- per-ray gradients g_weights_sum / g_depth / g_image (+ the forward results): sample
 *   gradients for every marched row (zero behind the early stop) and the compact live-row list, as the one-pass kernel. */
int lnb_lidar_composite_forward(const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                const float *gt, const float *nears, const float *noises, float dt_gamma,
                                uint32_t max_steps, uint32_t C, uint32_t H, uint32_t M, uint32_t N, float T_thresh,
                                float *weights_sum, float *depth, float *image, float *t0, lnb_stream_t stream);
int lnb_lidar_composite_backward(const float *g_weights_sum, const float *g_depth, const float *g_image,
                                 const float *sigmas, const float *rgbs, const float *deltas, const int32_t *rays,
                                 const float *gt, const float *nears, const float *noises, float dt_gamma,
                                 uint32_t max_steps, uint32_t C, uint32_t H, const int32_t *counter, uint32_t M,
                                 uint32_t N, float T_thresh, const float *weights_sum, const float *depth,
                                 const float *image, float *grad_sigmas, float *grad_rgbs, int32_t *live_idx,
                                 int32_t *n_live, lnb_stream_t stream);
int lnb_field_head_out_grad(const float *g_rgb, const float *rgb, uint32_t M, void *g_head_out,
                            const int32_t *n_active, lnb_stream_t stream);
int lnb_field_sigma_out_grad(const float *g_sigma, const void *sigma_out, const void *g_head_in, uint32_t M,
                             uint32_t in_pad, uint32_t degree, float density_scale, void *g_sigma_out,
                             const int32_t *n_active, lnb_stream_t stream);
/* get_lidar_rays (dataset/base_dataset.py:85-100): pose [3x4 or 4x4 row-major], inds [N] flat pixel ids */
int lnb_lidar_rays(const float *pose, const int32_t *inds, uint32_t N, uint32_t H, uint32_t W, float fov_up,
                   float fov, float *rays_o, float *rays_d, lnb_stream_t stream);
/* One training batch from a frame resident on the device - the per-step collate of kitti360_dataset.py:123-159:
 * lnb_lidar_rays + gather of the sampled pixels' ground-truth rows from image [H*W, 3] (ray-drop, intensity, depth) */
int lnb_lidar_batch(const float *pose, const int32_t *inds, const float *image, uint32_t N, uint32_t H, uint32_t W,
                    float fov_up, float fov, float *rays_o, float *rays_d, float *gt, lnb_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused LiDAR field network (density MLP -> LiDAR head) - same arithmetic as
 *   lnb_ffmlp_forward_ex(sigma) -> lnb_field_head_input -> lnb_ffmlp_forward_ex(head) -> lnb_field_head_rgb
 * (nerf/network.py:162-237 wiring over ffmlp.cu:460-576 / freqencoder.cu:34-61), using the fact that every sample
 * of a ray shares its direction: W_in[:, :nfreq] . freq_enc(dir) is evaluated once per ray (lnb_field_ray_terms)
 * and enters the head's first layer as a per-ray bias; only the 15 geo features go through the tensor cores.
 * Layer shapes as lnb_ffmlp_*: hidden = 64, outputs padded to 16, ReLU, no biases.
 * ---------------------------------------------------------------------------------------- */
/* Direction encoding of the head input.  Every `degree` argument of the lnb_field_* entry points takes either a
 * frequency degree (freqencoder.cu:34-61: 3 + 6*degree columns; the LiDAR head of network.py:83 uses 12) or
 * LNB_DIR_SH(deg): real spherical harmonics of that degree (shencoder.cu:31-277: deg*deg columns; the direction
 * encoder of network.py:64 / network_tcnn.py:74-80, degree 4), evaluated once per ray like the frequency terms. */
#define LNB_DIR_SH(deg) (0x100u | (uint32_t)(deg))
/* LNB_OK when the fused kernels implement this configuration (else use the unfused chain) */
int lnb_field_supported(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                        uint32_t degree, uint32_t hidden);
/* rays_d [N,3], w_head = flat fp16 head weights ->
 *   ray_enc [N,in_pad] fp16 = [freq_enc(dir) | 0] ; ray_bias [N,64] fp32 = W_in[:, :nfreq] . ray_enc[n, :nfreq] */
int lnb_field_ray_terms(const float *rays_d, const void *w_head, uint32_t N, uint32_t degree, uint32_t in_pad,
                        void *ray_enc, float *ray_bias, lnb_stream_t stream);
/* enc [M,enc_dim] fp16 -> sigma [M] fp32 = exp(h0)*density_scale, rgb [M,2] fp32 = sigmoid(head[0:2]); keeps for the
 * backward pass: fb_sigma [sigma_layers,M,64], sig_out [M,16], fb_head [head_layers,M,64] (all fp16). */
int lnb_field_forward(const void *enc, const void *w_sigma, const void *w_head, const int32_t *ray_ids,
                      const float *ray_bias, uint32_t M, uint32_t enc_dim, uint32_t sigma_layers,
                      uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                      float density_scale, void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                      const int32_t *n_active, lnb_stream_t stream);
/* L2 residency of the hash table for the persistent forward kernel: declares [table, table + bytes) as the table later
 * lnb_field_fused_forward*() launches read; reserves persisting L2 for it and makes those launches carry an access-policy
 * window (table lines persisting, everything else streaming).  bytes = 0 switches it off.  Process-wide setting.
 * (No counterpart in the reference: its gather, gridencoder.cu:95-199, leaves residency to the hardware.) */
int lnb_field_set_l2_window(const void *table, size_t bytes);

/* ------------------------------------------------------------------------------------------
 * ONE persistent kernel per ray packet: lnb_grid_encode_forward_ex + lnb_field_forward in a single launch
 * (north star of this library; replaces gridencoder.cu:95-199 + ffmlp.cu:460-576 x 2 + the torch glue of
 * network.py:162-237 between them).  One CTA per SM, warp-specialised: 16 gather warps (warp <-> level) write the
 * fp16 features straight into the shared-memory operand tile of the first MLP layer while the tensor core (one MMA
 * warp, tcgen05 + TMEM) and two 128-thread epilogue groups work on earlier tiles.  The MLP weights arrive through the
 * TMA engine (cp.async.bulk on an mbarrier) from a pre-laid-out image: call lnb_field_pack_weights once per parameter
 * update.  Results are bit-identical to the two-kernel path; `enc` is still written (the backward pass reads it).
 * Supported: C = 2 fp16 hash/tiled grids with L*C in {16,32,48,64}, D = 3, linear interpolation, align_corners = 0;
 * MLP shapes as lnb_field_supported.  lnb_field_fused_weight_bytes returns 0 for unsupported shapes.
 * ---------------------------------------------------------------------------------------- */
size_t lnb_field_fused_weight_bytes(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                                    uint32_t degree, uint32_t hidden);
/* w_sigma / w_head: flat fp16 weights (ffmlp.cu:861-864 layout) -> image [lnb_field_fused_weight_bytes], which must
 * have been zero-filled once when it was allocated */
int lnb_field_pack_weights(const void *w_sigma, const void *w_head, uint32_t enc_dim, uint32_t sigma_layers,
                           uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden, void *image,
                           lnb_stream_t stream);
/* xyzs [M,3] fp32 in [-in_bound, in_bound] (in_bound = 0: already in [0,1]); table/offsets/L/C/S/H as
 * lnb_grid_encode_forward; outputs as lnb_grid_encode_forward_ex (enc, [M, L*C] layout) + lnb_field_forward */
int lnb_field_fused_forward(const float *xyzs, const void *table, const int32_t *offsets, uint32_t L, uint32_t C, float S,
                            uint32_t H, float in_bound, const void *weight_image, const int32_t *ray_ids,
                            const float *ray_bias, uint32_t M, uint32_t sigma_layers, uint32_t head_in_pad,
                            uint32_t head_layers, uint32_t degree, uint32_t hidden, float density_scale, void *enc,
                            void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                            const int32_t *n_active, lnb_stream_t stream);
/* LiDAR-head backward = lnb_field_head_out_grad + lnb_ffmlp_backward_accumulate(head) + lnb_field_sigma_out_grad:
 * g_rgb/rgb [M,2], g_sigma [M] -> g_sig_out [M,16] fp16 (gradient w.r.t. the density MLP's output row) and
 * grad_w_head_f32 (+=, flat fp32, head weight layout). */
int lnb_field_head_backward(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                            const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                            uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                            float density_scale, void *g_sig_out, float *grad_w_head_f32, const int32_t *n_active,
                            lnb_stream_t stream);
/* Row-compacted backward entry points (the fused step's backward pass on live samples only, see
 * lnb_lidar_composite_step): row r of the kernel reads its per-row INPUTS (saved activations, layer inputs, sig_out,
 * g_rgb, rgb, g_sigma, ray_ids, sample coordinates) at row row_idx[r]; the gradients handed from one backward kernel
 * to the next (g_sig_out, grad_inputs, grid `grad`) are in compact order; n_rows[0] (device) is the exact row count. */
int lnb_field_head_backward_rows(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                                 const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                                 uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree,
                                 uint32_t hidden, float density_scale, void *g_sig_out, float *grad_w_head_f32,
                                 const int32_t *row_idx, const int32_t *n_rows, lnb_stream_t stream);
int lnb_ffmlp_backward_accumulate_rows(const void *grad, const void *inputs, const void *weights,
                                       const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                       uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                       uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                       float *grad_weights_f32, const int32_t *row_idx, const int32_t *n_rows,
                                       lnb_stream_t stream);
/* [B, L*C] gradient layout only */
int lnb_grid_encode_backward_rows(const void *grad, const float *inputs, const void *embeddings,
                                  const int32_t *offsets, void *grad_embeddings, uint32_t B, uint32_t D,
                                  uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype, int align_corners,
                                  uint32_t interp, int dtype, float in_bound, int accumulate_f32,
                                  const int32_t *row_idx, const int32_t *n_rows, lnb_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * Evaluation-side callers of the hot path: range image <-> point cloud
 * (lidarnerf/convert.py) and the Chamfer nearest-neighbour search (extern/chamfer3D/chamfer3D.cu).
 * ---------------------------------------------------------------------------------------- */

/* chamfer_3D.forward (extern/chamfer3D/chamfer_cuda.cpp:26-33, chamfer3D.cu:135-166): xyz1 [B,N,3], xyz2 [B,M,3] ->
 * dist1 [B,N] / idx1 [B,N] = squared distance to / index of the nearest xyz2 point (first minimum), and the same
 * for xyz2 against xyz1. */
int lnb_chamfer_forward(const float *xyz1, const float *xyz2, uint32_t B, uint32_t N, uint32_t M, float *dist1,
                        float *dist2, int32_t *idx1, int32_t *idx2, lnb_stream_t stream);
/* chamfer_3D.backward (chamfer_cuda.cpp:35-45, chamfer3D.cu:167-236): ADDS into grad_xyz1 / grad_xyz2 (caller zeroes) */
int lnb_chamfer_backward(const float *xyz1, const float *xyz2, float *grad_xyz1, float *grad_xyz2,
                         const float *grad_dist1, const float *grad_dist2, const int32_t *idx1, const int32_t *idx2,
                         uint32_t B, uint32_t N, uint32_t M, lnb_stream_t stream);
/* lidar_to_pano_with_intensities (lidarnerf/convert.py:99-160): points [N, point_stride >= 3] (x, y, z[, intensity])
 * in the sensor frame -> pano [H,W] (range of the closest point of each pixel, 0 = empty) and intensities [H,W]
 * (nullable; 0 when point_stride == 3).  workspace: lnb_lidar_to_pano_workspace_bytes(H, W) bytes. */
size_t lnb_lidar_to_pano_workspace_bytes(uint32_t H, uint32_t W);
int lnb_lidar_to_pano(const float *points, uint32_t point_stride, uint32_t N, uint32_t H, uint32_t W, float fov_up,
                      float fov, float max_depth, float *pano, float *intensities, void *workspace,
                      lnb_stream_t stream);
/* pano_to_lidar_with_intensities (lidarnerf/convert.py:194-235): every non-empty pixel -> (x, y, z, intensity) along
 * its beam, in row-major pixel order; points_out [H*W,4] (16-byte aligned), count_out[0] = number of points.
 * workspace: lnb_pano_to_lidar_workspace_bytes(H, W) bytes. */
size_t lnb_pano_to_lidar_workspace_bytes(uint32_t H, uint32_t W);
int lnb_pano_to_lidar(const float *pano, const float *intensities, uint32_t H, uint32_t W, float fov_up, float fov,
                      float *points_out, int32_t *count_out, void *workspace, lnb_stream_t stream);


/* ------------------------------------------------------------------------------------------
 * bf16 builds of the tensor-core MLP entry points (BASELINE config 5: "bf16 MLP on tensor cores"): identical signatures
 * and semantics, but weights, activations, saved activations, sig_out / ray_enc rows and activation gradients are
 * bfloat16 (tcgen05 kind::f16 with bf16 operand formats, fp32 accumulation).  The hash table stays fp16 (the persistent
 * forward kernel converts the interpolated features) and the input gradient the density MLP hands to the hash-grid
 * scatter stays fp16.  Same translation units compiled with -DLNB_BF16 (lidar-nerf_b200/build.py).
 * ---------------------------------------------------------------------------------------- */
int lnb_field_supported_bf16(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                        uint32_t degree, uint32_t hidden);
int lnb_field_ray_terms_bf16(const float *rays_d, const void *w_head, uint32_t N, uint32_t degree, uint32_t in_pad,
                        void *ray_enc, float *ray_bias, lnb_stream_t stream);
int lnb_field_forward_bf16(const void *enc, const void *w_sigma, const void *w_head, const int32_t *ray_ids,
                      const float *ray_bias, uint32_t M, uint32_t enc_dim, uint32_t sigma_layers,
                      uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                      float density_scale, void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                      const int32_t *n_active, lnb_stream_t stream);
int lnb_field_head_backward_bf16(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                            const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                            uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden,
                            float density_scale, void *g_sig_out, float *grad_w_head_f32, const int32_t *n_active,
                            lnb_stream_t stream);
int lnb_field_head_backward_rows_bf16(const float *g_rgb, const float *rgb, const float *g_sigma, const void *sig_out,
                                 const int32_t *ray_ids, const void *ray_enc, const void *w_head, const void *fb_head,
                                 uint32_t M, uint32_t head_in_pad, uint32_t head_layers, uint32_t degree,
                                 uint32_t hidden, float density_scale, void *g_sig_out, float *grad_w_head_f32,
                                 const int32_t *row_idx, const int32_t *n_rows, lnb_stream_t stream);
size_t lnb_field_fused_weight_bytes_bf16(uint32_t enc_dim, uint32_t sigma_layers, uint32_t head_in_pad, uint32_t head_layers,
                                    uint32_t degree, uint32_t hidden);
int lnb_field_pack_weights_bf16(const void *w_sigma, const void *w_head, uint32_t enc_dim, uint32_t sigma_layers,
                           uint32_t head_in_pad, uint32_t head_layers, uint32_t degree, uint32_t hidden, void *image,
                           lnb_stream_t stream);
int lnb_field_fused_forward_bf16(const float *xyzs, const void *table, const int32_t *offsets, uint32_t L, uint32_t C, float S,
                            uint32_t H, float in_bound, const void *weight_image, const int32_t *ray_ids,
                            const float *ray_bias, uint32_t M, uint32_t sigma_layers, uint32_t head_in_pad,
                            uint32_t head_layers, uint32_t degree, uint32_t hidden, float density_scale, void *enc,
                            void *fb_sigma, void *sig_out, float *sigma, void *fb_head, float *rgb,
                            const int32_t *n_active, lnb_stream_t stream);
int lnb_ffmlp_forward_ex_bf16(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                         uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                         uint32_t output_activation, void *forward_buffer, void *outputs, const int32_t *n_active,
                         lnb_stream_t stream);
int lnb_ffmlp_inference_bf16(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim,
                        uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                        uint32_t activation, uint32_t output_activation, void *inference_buffer,
                        void *outputs, lnb_stream_t stream);
int lnb_ffmlp_backward_accumulate_bf16(const void *grad, const void *inputs, const void *weights,
                                  const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                  uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                  uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                  float *grad_weights_f32, const int32_t *n_active, lnb_stream_t stream);
int lnb_ffmlp_backward_accumulate_rows_bf16(const void *grad, const void *inputs, const void *weights,
                                       const void *forward_buffer, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                                       uint32_t hidden_dim, uint32_t num_layers, uint32_t activation,
                                       uint32_t output_activation, int calc_grad_inputs, void *grad_inputs,
                                       float *grad_weights_f32, const int32_t *row_idx, const int32_t *n_rows,
                                       lnb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LIDARNERF_B200_H_ */
