/*
 * seal3d_b200.h -- C ABI of libseal3d_b200.so: the sm_100a implementation of Seal-3D's hot path.
 *
 * Every entry point below replaces one function of the reference's pybind surface (the file:line
 * of the declaration it replaces is cited) or adds a fused / training entry that the reference
 * spreads over Python.  Conventions:
 *   - plain C: device pointers + sizes, no torch / C++ types.  All array pointers are DEVICE
 *     pointers unless the parameter name starts with h_ (small host-side constant blocks).
 *   - the CALLER allocates every output (like the reference, SURVEY.md 8b "Ownership").
 *   - the last argument is the cudaStream_t to launch on (as void*); calls are asynchronous.
 *   - return value: 0 ok; > 0 a cudaError_t from the launch; < 0 an argument error
 *     (S3D_EINVAL = -22, S3D_ENOTSUP = -95: shape not built).  The reference throws
 *     std::runtime_error / TORCH_CHECK in the same situations; the pybind shims translate
 *     non-zero codes back into RuntimeError.
 *   - float tensors are float32 unless a `dtype` argument says otherwise (0 = f32, 1 = f16).
 */
#ifndef SEAL3D_B200_H
#define SEAL3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ raymarching ----------- */
/* raymarching/src/raymarching.h:7   near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars) */
int s3d_near_far_from_aabb(const float *rays_o, const float *rays_d, const float *aabb, uint32_t N, float min_near,
                           float *nears, float *fars, void *stream);
/* raymarching.h:8   sph_from_ray(rays_o, rays_d, radius, N, coords[N,2]) */
int s3d_sph_from_ray(const float *rays_o, const float *rays_d, float radius, uint32_t N, float *coords, void *stream);
/* raymarching.h:9   morton3D(coords int32[N,3], N, indices int32[N]) */
int s3d_morton3D(const int *coords, uint32_t N, int *indices, void *stream);
/* raymarching.h:10  morton3D_invert(indices, N, coords) */
int s3d_morton3D_invert(const int *indices, uint32_t N, int *coords, void *stream);
/* raymarching.h:11  packbits(grid float[C*H^3], N = C*H^3/8, density_thresh, bitfield uint8[N]) */
int s3d_packbits(const float *grid, uint32_t N, float density_thresh, uint8_t *bitfield, void *stream);
/* raymarching.h:13  march_rays_train(...).  xyzs/dirs/deltas must be zero-filled by the caller (rows past the
 * last sample stay zero); rays int32[N,3] = (ray id, first sample, sample count) in ray-major, deterministic
 * order; counter int32[2] += (total samples, N); noises float[N] or NULL (= 0). */
int s3d_march_rays_train(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound, float dt_gamma,
                         uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float *nears,
                         const float *fars, float *xyzs, float *dirs, float *deltas, int *rays, int *counter,
                         const float *noises, void *stream);
/* the same op in two phases, for an exactly-sized sample buffer (count+scan, then write) */
int s3d_march_rays_train_count(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                               const float *nears, const float *fars, int *rays, int *counter, const float *noises,
                               void *stream);
int s3d_march_rays_train_write(const float *rays_o, const float *rays_d, const uint8_t *grid, float bound,
                               float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                               const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                               const int *rays, const float *noises, void *stream);
/* raymarching.h:14  composite_rays_train_forward */
int s3d_composite_rays_train_forward(const float *sigmas, const float *rgbs, const float *deltas, const int *rays,
                                     uint32_t M, uint32_t N, float T_thresh, float *weights_sum, float *depth,
                                     float *image, void *stream);
/* raymarching.h:15  composite_rays_train_backward (grad_sigmas / grad_rgbs zero-filled by the caller) */
int s3d_composite_rays_train_backward(const float *grad_weights_sum, const float *grad_image, const float *sigmas,
                                      const float *rgbs, const float *deltas, const int *rays,
                                      const float *weights_sum, const float *image, uint32_t M, uint32_t N,
                                      float T_thresh, float *grad_sigmas, float *grad_rgbs, void *stream);
/* raymarching.h:17  march_rays (inference; outputs zero-filled by the caller, fixed stride n_step) */
int s3d_march_rays(uint32_t n_alive, uint32_t n_step, const int *rays_alive, const float *rays_t, const float *rays_o,
                   const float *rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H,
                   const uint8_t *grid, const float *nears, const float *fars, float *xyzs, float *dirs, float *deltas,
                   const float *noises, void *stream);
/* raymarching.h:18  composite_rays (in place; finished rays get rays_alive = -1) */
int s3d_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int *rays_alive, float *rays_t,
                       const float *sigmas, const float *rgbs, const float *deltas, float *weights_sum, float *depth,
                       float *image, void *stream);

/* ------------------------------------------------------------------ gridencoder ----------- */
/* gridencoder/src/gridencoder.h:12  grid_encode_forward.  inputs float[B,D] in [0,1]; embeddings [sO,C];
 * offsets int32[L+1]; outputs [L,B,C]; dy_dx [B,L,D,C] or NULL; S = log2(per_level_scale); gridtype 0 hash /
 * 1 tiled; interp 0 linear / 1 smoothstep.  D in 2..4, C in {1,2,4,8}. */
int s3d_grid_encode_forward(const float *inputs, const void *embeddings, const int *offsets, void *outputs, uint32_t B,
                            uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, void *dy_dx, uint32_t gridtype,
                            int align_corners, uint32_t interp, int dtype, void *stream);
/* gridencoder.h:13  grid_encode_backward.  grad [L,B,C]; grad_embeddings is accumulated into (caller zero-fills);
 * dy_dx / grad_inputs [B,D] optional. */
int s3d_grid_encode_backward(const void *grad, const float *inputs, const void *embeddings, const int *offsets,
                             void *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                             const void *dy_dx, void *grad_inputs, uint32_t gridtype, int align_corners, uint32_t interp,
                             int dtype, void *stream);
/* gridencoder.h:15  grad_total_variation (float32 tables only) */
int s3d_grad_total_variation(const void *inputs, const void *embeddings, void *grad, const int *offsets, float weight,
                             uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                             int align_corners, int dtype, void *stream);

/* diagnostic: scales[l] = exp2f(l*S)*H - 1 as evaluated on the device (gridencoder.cu:138), float[L] */
int s3d_grid_level_scales(uint32_t L, float S, uint32_t H, float *scales, void *stream);

/* ------------------------------------------------------------------ shencoder / freqencoder */
/* shencoder/src/shencoder.h:9   sh_encode_forward(inputs[B,3], outputs[B,C*C], B, D=3, C=degree<=8, dy_dx[B,3*C*C]|NULL) */
int s3d_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t D, uint32_t C, float *dy_dx,
                          void *stream);
/* shencoder.h:10  sh_encode_backward: grad_inputs[B,3] += sum_ch grad * dy_dx */
int s3d_sh_encode_backward(const float *grad, const float *inputs, uint32_t B, uint32_t D, uint32_t C,
                           const float *dy_dx, float *grad_inputs, void *stream);
/* freqencoder/src/freqencoder.h:7   freq_encode_forward(inputs[B,D], B, D, deg, C = D + 2*D*deg, outputs[B,C]) */
int s3d_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float *outputs,
                            void *stream);
/* freqencoder.h:10  freq_encode_backward(grad[B,C], outputs[B,C], ..., grad_inputs[B,D]) */
int s3d_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                             float *grad_inputs, void *stream);

/* ------------------------------------------------------------------ ffmlp (all tensors float16) */
/* ffmlp/src/ffmlp.h:8   ffmlp_forward: inputs[B,in], weights flat [hidden*in | hidden*hidden*(nl-1) | out*hidden],
 * forward_buffer[nl,B,hidden], outputs[B,out].  hidden in {16,32,64}, in % 16 == 0 (<= 256), out <= 16, nl in 2..5. */
int s3d_ffmlp_forward(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                      uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                      void *forward_buffer, void *outputs, void *stream);
/* ffmlp.h:9   ffmlp_inference (no activations are written; inference_buffer unused) */
int s3d_ffmlp_inference(const void *inputs, const void *weights, uint32_t B, uint32_t input_dim, uint32_t output_dim,
                        uint32_t hidden_dim, uint32_t num_layers, uint32_t activation, uint32_t output_activation,
                        void *inference_buffer, void *outputs, void *stream);
/* ffmlp.h:11  ffmlp_backward: grad[B,out], backward_buffer[nl,B,hidden], grad_inputs[B,in] (if calc_grad_inputs),
 * grad_weights flat (overwritten) */
int s3d_ffmlp_backward(const void *grad, const void *inputs, const void *weights, const void *forward_buffer, uint32_t B,
                       uint32_t input_dim, uint32_t output_dim, uint32_t hidden_dim, uint32_t num_layers,
                       uint32_t activation, uint32_t output_activation, int calc_grad_inputs, void *backward_buffer,
                       void *grad_inputs, void *grad_weights, void *stream);
/* ffmlp.h:13-14  allocate_splitk / free_splitk: kept for interface parity, nothing to allocate */
int s3d_allocate_splitk(size_t size);
int s3d_free_splitk(void);

/* ------------------------------------------------------------------ Seal proxy mapping ---- */
/* SealNeRF/seal_utils.py:237-279 SealBBoxMapper.map_to_origin (+ :132-153 map_mask, :630-685 points_in_mesh):
 * h_transform[16] inverse 4x4 row-major, h_rotation[9] inverse 3x3, h_scale[3] = 1/scale, h_center[3];
 * bounds device [nb,2,3]; tris device [F,3,3]; h_test_dir[3]|NULL; h_src_bound[6]+h_map_source[3]|NULL;
 * out_points/out_dirs are full copies with the masked rows replaced; mask uint8[P]. */
int s3d_seal_bbox_map_to_origin(const float *points, const float *dirs, uint32_t P, const float *h_transform,
                                const float *h_rotation, const float *h_scale, const float *h_center,
                                const float *bounds, uint32_t nb, const float *tris, uint32_t F, const float *h_test_dir,
                                const float *h_src_bound, const float *h_map_source, float *out_points, float *out_dirs,
                                uint8_t *mask, void *stream);
/* seal_utils.py:48-57,739-769 SealMapper.map_color (hsv shift and/or rgb replacement with V re-lighting around the
 * batch mean of V), in place on rows with mask != 0; d_stats = device float[2] scratch */
int s3d_seal_map_color(float *rgbs, const uint8_t *mask, uint32_t M, const float *h_hsv_mod, const float *h_rgb_target,
                       float light_offset, float *d_stats, void *stream);
/* SURVEY 8f-4.  SealBrushMapper.map_to_origin (seal_utils.py:408-453; map_mask :132-153 with the brush's own test
 * direction): inside samples move by -normal_expand, plus |att - d| / att * normal_expand where the projected sample is
 * closer than `attenuation_distance` to the nearest border point.  mode 0 = 'linear', 1 = 'dry'.  h_* host float[3],
 * d_* device arrays (bounds [nb,2,3], tris [F,3,3], border points [K,3]). */
int s3d_seal_brush_map_to_origin(const float *points, uint32_t P, const float *d_bounds, uint32_t nb, const float *d_tris, uint32_t F,
                                 const float *h_test_dir, const float *h_normal_expand, const float *h_center,
                                 const float *d_border_points, uint32_t K, float attenuation_distance, int mode,
                                 float *out_points, uint8_t *mask, void *stream);
/* SealAnchorMapper.map_to_origin (seal_utils.py:514-570): cone / plane-side test around the anchor, pull along v_h,
 * per-axis scale about the anchor.  Like the reference, the map-region mask only gates an early exit: once any sample
 * is inside, every sample takes the cone test and `mask` is the cone mask.  d_flag: device int scratch. */
int s3d_seal_anchor_map_to_origin(const float *points, uint32_t P, const float *d_bounds, uint32_t nb, const float *d_tris, uint32_t F,
                                  const float *h_test_dir, const float *h_v_anchor, const float *h_v_offset, const float *h_v_h,
                                  float len_h, float radius, const float *h_scale, int *d_flag, float *out_points, uint8_t *mask,
                                  void *stream);
/* texture branch of SealMapper.map_color (seal_utils.py:58-79): target colour = image[pixel of the sample projected on
 * the image plane], V re-lit around the batch mean like modify_rgb, blended with image_mask[pixel]. */
int s3d_seal_map_color_image(float *rgbs, const float *points, const uint8_t *mask, uint32_t M, const float *d_image,
                             const float *d_alpha, uint32_t H, uint32_t W, const float *h_norm, const float *h_o, const float *h_w,
                             const float *h_h, float light_offset, float *d_stats, void *stream);
/* SealNeRF/renderer.py:21-66 init_mapper + hack_bitfield: cells [h_cell_lo, h_cell_hi) -> bitfield bytes = 255 */
int s3d_seal_force_fill_bitfield(uint8_t *bitfield, const int *h_cell_lo, const int *h_cell_hi, uint32_t H,
                                 uint32_t cascade_index, void *stream);

/* ------------------------------------------------------------------ distillation step glue - */
/* SealNeRF/trainer.py:456-469: loss[0] += mean|ds| + mean|dc|; grads of the means (either may be NULL) */
int s3d_pretrain_loss(const float *sigma_s, const float *rgb_s, const float *sigma_t, const float *rgb_t, uint32_t M,
                      float *loss, float *grad_sigma, float *grad_rgb, void *stream);
/* nerf/utils.py:484-489,530: student image = comp + (1-ws)*bg; loss[0] += MSE, loss[1] += L1 depth (if depth_t);
 * grad_image = dL/dcomp, grad_ws = dL/dws */
int s3d_finetune_loss(const float *comp_s, const float *ws_s, const float *depth_s, const float *image_t,
                      const float *depth_t, uint32_t N, float bg_color, float *loss, float *grad_image, float *grad_ws,
                      void *stream);
/* main_SealNeRF.py:283-284 torch.optim.Adam semantics in one pass over a flat arena; writes the fp16 shadow
 * (or NULL) and zeroes the gradient when asked.  grad_dtype 0 = f32, 1 = f16. */
int s3d_adam_step(float *params, void *grads, float *exp_avg, float *exp_avg_sq, void *shadow_f16, uint64_t n, float lr,
                  float beta1, float beta2, float eps, uint32_t step, float grad_scale, int zero_grad, int grad_dtype,
                  const float *scaler_state, void *stream);
/* torch.cuda.amp.GradScaler as the reference trainer drives it (nerf/utils.py:361,857-859: scale(loss).backward(),
 * step(optimizer), update()), kept ON THE DEVICE so a step never waits for the host:
 *   scaler_state = float[16], two 8-float blocks (torch.optim.Adam counts steps per parameter, so the hash tables and the
 *   MLP arena -- frozen during pretraining, SealNeRF/trainer.py:484-488 -- each get their own count):
 *     block 0 (tables / a flat arena): [0] loss scale  [1] growth tracker  [2] found_inf  [3] optimizer steps applied
 *                                      [4] 1/(1-beta1^t)  [5] 1/sqrt(1-beta2^t)   (t = [3], refreshed by _check)
 *     block 1 = scaler_state + 8 (MLP): [8] scale, [10] found_inf (copies), [11] steps, [12] [13] bias corrections
 * s3d_grad_scaler_check scans the (all-reduced) gradient arena for non-finite values, sets found_inf and, when the
 * step will be applied, advances [3] (and [11] when advance_mlp != 0) and the bias corrections.  The Adam entry points,
 * given scaler_state (block 0) or scaler_state + 8 (block 1), divide the
 * gradient by [0], use [4],[5] instead of the host `step`, and when found_inf is set only clear the gradient (the
 * skipped step of GradScaler.step).  s3d_grad_scaler_update = GradScaler.update(): backoff on overflow, growth after
 * growth_interval clean steps, found_inf reset.  scaler_state NULL = static scale through grad_scale (as before). */
int s3d_grad_scaler_check(const float *grads, uint64_t n, float *scaler_state, float beta1, float beta2, int advance_mlp, void *stream);
int s3d_grad_scaler_update(float *scaler_state, float growth_factor, float backoff_factor, uint32_t growth_interval, void *stream);
/* torch_ema.ExponentialMovingAverage.update (nerf/utils.py:356-357,882-883): shadow -= (1 - decay) * (shadow - param) */
int s3d_ema_update(float *shadow, const float *params, uint64_t n, float decay, void *stream);
int s3d_cast_f32_to_f16(const float *src, void *dst, uint64_t n, void *stream);
/* nerf/renderer.py:445-538 update_extra_state as a chain of launches without a host read (csrc/density.cu):
 *   s3d_density_pick_cells   :487-499 the partial update's cell list for one cascade: cells_out[0, n_uniform) = morton3D of
 *                            uniformly drawn coords, [n_uniform, n_uniform + n_occ) = draws (with repetition) from the occupied
 *                            cells {density_grid_cas > 0}, compacted on the device in ascending order (replaces torch.nonzero);
 *                            n_occupied_out (device uint32, optional) receives the size of that set
 *   s3d_density_cells_to_xyz :470-479 / :501-509 cell -> centre in the cascade + uniform jitter of half a cell, one rounding per
 *                            torch op; cell_morton NULL = cells 0..n-1 (the full sweep of the first 16 refreshes)
 *   s3d_density_scatter      :483 / :513 tmp_grid[cell] = sigma * density_scale; a cell drawn more than once keeps its largest
 *                            value (the reference's index_put keeps an arbitrary one); cell_morton NULL = identity
 *   s3d_density_grid_update  :521-524 + :528 grid = max(grid * decay, tmp) where both >= 0; stats_out[0] = mean(clamp(grid, 0)),
 *                            stats_out[1] = min(mean, density_thresh), reduced in a fixed order (bit-identical across ranks)
 *   s3d_packbits_dev_thresh  :529-530 packbits with the threshold read from device memory (stats_out + 1)
 *   s3d_mean_count           :533-536 int(sum(step_counter[:total_step, 0]) / total_step) -> device int (-1 when total_step = 0)
 * Draws are a counter-based hash of (seed, index): every data-parallel rank derives the same cells and jitter from the
 * step-derived seed.  s3d_density_grid_ema is the older single-launch EMA with an atomically accumulated sum. */
int s3d_density_pick_cells(const float *density_grid_cas, uint32_t H, uint32_t n_uniform, uint32_t n_occ, uint32_t seed, int *cells_out,
                           uint32_t *n_occupied_out, void *stream);
int s3d_density_cells_to_xyz(const int *cell_morton, uint32_t n, uint32_t H, float bound_cas, uint32_t seed, float *xyz,
                             void *stream);
int s3d_density_scatter(const int *cell_morton, const float *sigma, uint32_t n, float density_scale, float *tmp_grid,
                        void *stream);
int s3d_density_grid_update(float *grid, const float *tmp_grid, uint32_t n, float decay, float density_thresh, float *stats_out,
                            void *stream);
int s3d_packbits_dev_thresh(const float *grid, uint32_t N, const float *density_thresh_dev, uint8_t *bitfield, void *stream);
int s3d_mean_count(const int *step_counter, uint32_t total_step, int *mean_count_out, void *stream);
int s3d_density_grid_ema(float *grid, const float *tmp_grid, uint32_t n, float decay, float *sum_out, void *stream);
/* Teacher + student NGP field forward on one sample buffer in ONE kernel (csrc/field.cu k_ngp_pair_fwd): hash-grid gather from
 * the paired fp16 table (s3d_ngp_pair_tables) written straight into tensor memory + both tcgen05 MLP chains; the two halves of
 * nerf/network.py:99-128 for SealNeRF's teacher (on the proxy-mapped samples) and student.  xyz_teacher / mask / dirs_teacher
 * describe the moved samples (NULL: none / dirs); feats_student [M,64] fp16 is what s3d_ngp_mlp_backward recomputes from. */
int s3d_ngp_pair_forward(const float *xyz, const float *xyz_teacher, const uint8_t *mask, const float *dirs, const float *dirs_teacher,
                         uint32_t M, float bound, const void *table8, const int *offsets, uint32_t L, float S, uint32_t H,
                         const void *t_s0, const void *t_s1, const void *t_c0, const void *t_c1, const void *t_c2, const void *s_s0,
                         const void *s_s1, const void *s_c0, const void *s_c1, const void *s_c2, float density_scale_teacher,
                         float density_scale_student, float *sigma_t, float *rgb_t, float *sigma_s, float *rgb_s, void *feats_student,
                         void *stream);
/* One bias-free nn.Linear on tensor cores (the wide FFMLP kernels with no hidden layer): y [B,out] = x [B,in] . W^T, W [out,in]
 * row-major, fp16 in / out, fp32 accumulation; in a multiple of 16 (<= 256), out <= 256.  TensoRF's basis_mat
 * (tensoRF/network.py:42,155).  Backward: grad_x [B,in] (NULL = not wanted), grad_w [out,in] overwritten. */
int s3d_linear_forward(const void *x, const void *w, uint32_t B, uint32_t in_dim, uint32_t out_dim, void *y, void *stream);
int s3d_linear_backward(const void *grad_y, const void *x, const void *w, uint32_t B, uint32_t in_dim, uint32_t out_dim, void *grad_x,
                        void *grad_w, void *stream);
/* TensoRF colour head glue (tensoRF/network.py:170-172): h = cat([freq(color_feat), freq(dirs)]) zero padded to K columns, fp16,
 * in one launch (feat fp16 [B, ld_feat] with Fd valid columns -- the zero-padded output rows of s3d_linear_forward --, dirs fp32
 * [B,3], both encoders `frequency` with multires = deg); backward: grad_feat fp16 [B, ld_feat] from grad_h and the stored h. */
int s3d_tensorf_head_encode(const void *feat, uint32_t Fd, uint32_t ld_feat, const float *dirs, uint32_t B, uint32_t deg, uint32_t K, void *h,
                            void *stream);
int s3d_tensorf_head_encode_backward(const void *grad_h, const void *h, uint32_t Fd, uint32_t ld_feat, uint32_t B, uint32_t deg, uint32_t K,
                                     void *grad_feat, void *stream);
/* development / test switch of the train marcher: 1 (default) walks a ray only inside the widened bounding box of the occupied
 * cells (single cascade, dt_gamma = 0; the step lattice before the box is jumped in closed form), 0 walks it from its near
 * point like the reference.  The samples are identical either way (tests/test_gpu_parity.py compares them at full size). */
int s3d_march_set_clip(int enable);
/* The per-ray part of a training step in one launch (csrc/raymarching.cu k_distill_rays): composite_rays_train_forward of the
 * teacher's sigma / rgb on the student's samples (sig_t, rgb_t; + background -> targets) or targets given as image_t [N,3] /
 * depth_t [N] (NULL ok), composite_rays_train_forward of the student, the loss of nerf/utils.py:484-489,530 (loss[0] += MSE
 * term, loss[1] += L1 depth term) and composite_rays_train_backward with the loss gradient times scale (*scale_dev if given).
 * grad_sigmas [M], grad_rgbs [M,3] zero-initialised by the caller. */
int s3d_distill_rays(const float *sig_t, const float *rgb_t, const float *image_t, const float *depth_t, const float *sig_s,
                     const float *rgb_s, const float *deltas, const int *rays, uint32_t M, uint32_t N, float T_thresh, float bg_color,
                     float scale, const float *scale_dev, float *loss, float *grad_sigmas, float *grad_rgbs, void *stream);
/* nerf/utils.py:53-140 get_rays: poses device [B,4,4] cam2world, pixel ids inds device int64 [inds_rows, N] (row * W + col,
 * inds_rows = 1 shares them across views like the reference's expand, = B per view) or NULL for all H*W pixels in order;
 * rays_o / rays_d [B,N,3] (unit directions) */
int s3d_get_rays(const float *poses, uint32_t B, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W, const long long *inds,
                 uint32_t inds_rows, uint32_t N, float *rays_o, float *rays_d, void *stream);
/* nerf/renderer.py:379-443 mark_untrained_grid(poses, intrinsic): density_grid[c, cell] = -1 for every cell whose centre no camera
 * sees (in front, inside the frustum widened by one cell).  poses device [B,4,4] cam2world (B <= 4000), kx = cx/fx, ky = cy/fy,
 * count_out optional int32 [C, H^3] (morton order) = number of cameras per cell */
int s3d_mark_untrained_grid(float *density_grid, const float *poses, uint32_t B, float kx, float ky, uint32_t C, uint32_t H, float bound,
                            int *count_out, void *stream);

/* ------------------------------------------------------------------ fused NGP field (nerf/network.py:99-128) */
/* The higher-level seam of SURVEY.md 8b: samples -> (sigma, rgb) and back, four kernels, one 128-byte fp16 row per
 * sample between them.  table4 = both hash tables interleaved, fp16 [N] x {s0,s1,c0,c1}; grad4 = fp32, same layout. */
int s3d_ngp_interleave_tables(const float *table_sigma, const float *table_color, void *table4, uint64_t n_entries, void *stream);
/* table entry i = {s0,s1,c0,c1} fp16 at table + i * table_stride (8: table4; 16: one half of a paired table8, pointer pre-offset) */
int s3d_ngp_encode(const float *xyz, uint32_t M, float bound, const void *table, uint32_t table_stride, const int *offsets, uint32_t L,
                   float S, uint32_t H, void *feats, int sigma_only, void *stream);
/* teacher + student on the same samples: table8 [N] x {teacher entry | student entry} (16 B); xyz_teacher / mask (uint8) = the
 * proxy-mapped positions and which samples moved (NULL: none) */
int s3d_ngp_pair_tables(const void *teacher_table4, const void *student_table4, void *table8, uint64_t n_entries, void *stream);
int s3d_ngp_encode_pair(const float *xyz, const float *xyz_teacher, const uint8_t *mask, uint32_t M, float bound, const void *table8,
                        const int *offsets, uint32_t L, float S, uint32_t H, void *feats_teacher, void *feats_student, void *stream);
/* weights: fp16 row-major nn.Linear matrices sigma_net.0 [64,32], sigma_net.1 [16,64], color_net.0 [64,63], .1 [64,64], .2 [3,64];
 * sigma = density_scale * exp(h[0]); geo [M,15] optional; rgb = sigmoid(...) */
int s3d_ngp_mlp_forward(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                        const void *w_c1, const void *w_c2, float density_scale, float *sigma, float *rgb, float *geo, int sigma_only,
                        void *stream);
/* dfeats [M,64] fp16 = out_scale * dL/dfeats; gw_* fp32 weight gradients (nn.Linear shapes), accumulated into */
int s3d_ngp_mlp_backward(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                         const void *w_c1, const void *w_c2, float density_scale, const float *g_sigma, const float *g_rgb, void *dfeats,
                         float out_scale, float *gw_s0, float *gw_s1, float *gw_c0, float *gw_c1, float *gw_c2, int train_mlp, void *stream);
int s3d_ngp_scatter(const float *xyz, const void *dfeats, uint32_t M, float bound, float *grad4, const int *offsets, uint32_t L, float S,
                    uint32_t H, float grad_scale, void *stream);
/* the scatter restricted to levels [level_begin, level_end), multiples of 4 (level_end may be L): data-parallel runs launch the
 * levels in chunks and all-reduce each finished slice of grad4 under the next chunk */
int s3d_ngp_scatter_levels(const float *xyz, const void *dfeats, uint32_t M, float bound, float *grad4, const int *offsets, uint32_t L,
                           float S, uint32_t H, float grad_scale, uint32_t level_begin, uint32_t level_end, void *stream);
/* Deterministic mode (no counterpart in the reference, whose gridencoder.cu:246-337 backward is float atomics): gradients are
 * accumulated as 64-bit fixed point (2^-32) with integer reductions, so the result does not depend on the order the reductions
 * arrive in and two runs of one step are bit-identical.  fixed4 [N_entries*4] / gw_* (fixed-point, the fp32 matrices' element
 * layout) are accumulated into; *nonfinite (device word) is raised by an inf / NaN / out-of-range contribution.
 * s3d_fixed_to_float: grad[i] += fixed[i] * 2^-32, fixed[i] = 0; a raised *nonfinite becomes a NaN in grad[0] and is cleared. */
int s3d_ngp_scatter_fixed(const float *xyz, const void *dfeats, uint32_t M, float bound, long long *fixed4, const int *offsets, uint32_t L,
                          float S, uint32_t H, float grad_scale, uint32_t *nonfinite, void *stream);
int s3d_ngp_mlp_backward_fixed(const void *feats, const float *dirs, uint32_t M, const void *w_s0, const void *w_s1, const void *w_c0,
                               const void *w_c1, const void *w_c2, float density_scale, const float *g_sigma, const float *g_rgb,
                               void *dfeats, float out_scale, long long *gw_s0, long long *gw_s1, long long *gw_c0, long long *gw_c1,
                               long long *gw_c2, int train_mlp, uint32_t *nonfinite, void *stream);
int s3d_fixed_to_float(long long *fixed, float *grad, size_t n, uint32_t *nonfinite, void *stream);
/* measurement aid: count[0] += the number of global reductions s3d_ngp_scatter issues for these samples, count[1] += those on
 * hashed levels (pseudo-random addresses); device, 64-bit each */
int s3d_ngp_scatter_count(const float *xyz, uint32_t M, float bound, const int *offsets, uint32_t L, float S, uint32_t H,
                          unsigned long long *count, void *stream);
int s3d_ngp_adam_tables(float *table_sigma, float *table_color, float *grad4, float *exp_avg4, float *exp_avg_sq4, void *shadow,
                        uint32_t shadow_stride, uint64_t n_entries, float lr, float beta1, float beta2, float eps, uint32_t step,
                        float grad_scale, const float *scaler_state, void *stream);
/* Data parallel over NVLink peer memory (new functionality: the reference's DDP path, nerf/utils.py:330-332, 940-954, is dead code):
 * this rank's shard [entry_begin, entry_end) of the tables is reduced over the `world` gradient arenas in array order, stepped
 * with Adam and written to every rank's fp32 tables and fp16 shadows in one launch.  *_peers are HOST arrays of `world` device
 * pointers (peer-mapped, e.g. from a symmetric-memory rendezvous), the same order on every rank; the caller brackets the launch
 * with cross-rank barriers and clears its own arena afterwards.  s3d_peer_sum: out = sum of the ranks' vectors, array order. */
int s3d_ngp_peer_adam_tables(const void *const *grad4_peers, void *const *sigma_peers, void *const *color_peers, void *const *shadow_peers,
                             uint32_t world, uint32_t rank, float *exp_avg4, float *exp_avg_sq4, uint32_t shadow_stride, uint64_t entry_begin,
                             uint64_t entry_end, float lr, float beta1, float beta2, float eps, uint32_t step, float grad_scale, void *stream);
int s3d_peer_sum(const void *const *src_peers, uint32_t world, float *out, uint64_t n, void *stream);

/* ------------------------------------------------------------------ TensoRF VM field */
/* The reference has no extension here: tensoRF/network.py:99-151 issues twelve F.grid_sample(bilinear, zeros,
 * align_corners=True) calls per batch; these entries are the kernel form of that seam (get_sigma_feat / get_color_feat).
 * Factor images are CHANNEL-LAST float32: plane i [H_i, W_i, R] (the reference's [1,R,H,W] parameter in torch channels_last
 * memory format), line i [D_i, R].  h_dims = HOST int[9] = {H_i, W_i, D_i} for i = 0..2; plane i is indexed by
 * (x[mat_ids[i][0]] -> W, x[mat_ids[i][1]] -> H), line i by x[vec_ids[i]] (network.py:37-38).  xyz [M,3] are world
 * coordinates normalised with aabb (device float[6], network.py:158) or, if aabb is NULL, already in [-1,1].
 * R % 4 == 0; reduce = 1 (sum over channels and planes, network.py:118-121) additionally needs R/4 a power of two <= 32. */
int s3d_vm_forward(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                   const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, int reduce,
                   float *out /* [M] or [M,3R] */, void *stream);
/* autograd of the above (grid_sampler_2d_backward w.r.t. the input images): g_* accumulate, caller zero-fills */
int s3d_vm_backward(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                    const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, int reduce,
                    const float *grad, float *g_mat0, float *g_mat1, float *g_mat2, float *g_vec0, float *g_vec1, float *g_vec2,
                    void *stream);
/* the unreduced lookup with fp16 output [M, 3R] / fp16 output gradient: the colour features of an fp16 (autocast) step, rounded
 * once on the way out like the fp16 cast the reference's autocast basis_mat applies (tensoRF/network.py:155) */
int s3d_vm_forward_f16(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                       const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, void *out, void *stream);
int s3d_vm_backward_f16(const float *xyz, uint32_t M, const float *aabb, const float *mat0, const float *mat1, const float *mat2,
                        const float *vec0, const float *vec1, const float *vec2, const int *h_dims, uint32_t R, const void *grad,
                        float *g_mat0, float *g_mat1, float *g_mat2, float *g_vec0, float *g_vec1, float *g_vec2, void *stream);
/* tensoRF/network.py:263-270 upsample_params: F.interpolate(bilinear, align_corners=True) of one channel-last image */
int s3d_vm_resize(const float *src, uint32_t H, uint32_t W, float *dst, uint32_t H2, uint32_t W2, uint32_t R, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEAL3D_B200_H */
