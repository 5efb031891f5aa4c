/*
 * nfs_b200.h — C ABI of the B200 (sm_100a) NeRF render hot path.
 *
 * This is the drop-in boundary below the reference's Python import surface
 * (src/models/*.py + src/utils/ray_utils.py of ANKITSANJYAL/nerf-few-shot-limitations).
 * The reference has no FFI layer of its own: its "operator API" is a set of
 * Python functions / nn.Modules that run chains of ATen ops.  Each entry point
 * below replaces one such chain and cites it (file:line, relative to the
 * reference root).  The Python wrappers in nerf-few-shot-limitations_b200/
 * bind these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers + sizes; all pointers are DEVICE pointers unless named
 *     h_* ; all float tensors are fp32, row-major, contiguous;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - every function returns 0 on success, a positive cudaError_t value when
 *     the CUDA runtime reported an error, or a negative NFS_E_* code for an
 *     argument error.  nfs_last_error_string() describes the last failure of
 *     the calling thread.  Nothing here synchronises the device.
 *   - no CPU fallback exists: without a CUDA device the functions return the
 *     runtime's error.
 */
#ifndef NFS_B200_H_
#define NFS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFS_B200_ABI_VERSION 7

/* negative return codes (argument errors) */
#define NFS_E_BADARG   (-1)  /* null pointer / non-positive size            */
#define NFS_E_TOOLARGE (-2)  /* size above what the kernel supports         */
#define NFS_E_ALIGN    (-3)  /* pointer not aligned as the kernel requires  */
#define NFS_E_UNSUPPORTED (-4)

int         nfs_abi_version(void);
const char *nfs_last_error_string(void);
/* Number of kernels this library has launched since load (per process);
 * bench.py reports the delta over its timed region as gpu_launches. */
uint64_t    nfs_launch_count(void);
/* Developer bisection switches for the fused MLP kernel (skip epilogue / weight reloads / MMAs);
 * results are wrong while non-zero.  0 = production behaviour. */
void        nfs_set_debug_flags(int32_t flags);
/* Developer timeline of the fused MLP kernel: buf = device buffer of 18*1024 uint64 (NULL = off); CTA 0 appends
 * (clock64 << 16 | event << 12 | layer << 4 | tile) per warp, entry 0 of each warp's 1024 = count. */
void        nfs_set_debug_trace(void *buf);

/* ------------------------------------------------------------------------- *
 * K1 — alpha compositing (volume rendering)
 *   replaces VolumeRenderer.forward        src/models/nerf_mlp.py:165-215
 *        and volume_render_radiance        src/models/volume_renderer.py:4-43
 *
 *   dists_i = (z_{i+1}-z_i) (last = 1e10) * ||rays_d||      nerf_mlp.py:181-185
 *   alpha_i = 1 - exp(-relu(density_i [+ noise_i*noise_std]) * dists_i)   :188-193
 *   T_i     = prod_{k<i} ((1-alpha_k) + 1e-10)              :196-199
 *   w_i     = alpha_i * T_i                                 :202
 *   rgb = sum w_i rgb_i ; depth = sum w_i z_i               :205-208
 *   white_bkgd: rgb += 1 - sum w_i                          :211-213
 *
 * Layouts (packed == 0, the VolumeRenderer form):
 *   rgb (N,S,3)  density (N,S)  z_vals (N,S)  rays_d (N,3)  noise (N,S)|NULL
 * Layout (packed != 0, the volume_render_radiance form):
 *   rgb = rgb_sigma (N,S,4) [R,G,B,sigma]; density is ignored (pass NULL).
 * Outputs: out_rgb (N,3); out_depth (N)|NULL; out_weights (N,S)|NULL.
 * One sub-warp (8/16/32 lanes) composites one ray; 128-bit loads when S%4==0.
 * ------------------------------------------------------------------------- */
int nfs_composite_fwd(const float *rgb, const float *density, const float *z_vals,
                      const float *rays_d, const float *noise, float noise_std,
                      int64_t n_rays, int32_t n_samples,
                      int32_t white_bkgd, int32_t packed,
                      float *out_rgb, float *out_depth, float *out_weights,
                      void *stream);

/* The forward with the loss fused into its epilogue: the same compositing, then, while the
 * pixel is still in registers,
 *   sum (rgb - target_rgb)^2            -> F.mse_loss(pred['rgb'], target['rgb'])   nerf_mlp.py:235, train.py:40
 *   sum |depth - target_depth|          -> F.l1_loss(pred['depth'], target['depth']) nerf_mlp.py:240 (target_depth|NULL)
 * are accumulated (fp64 atomics) into loss_sums[32][2] (32 slots the caller zeroes and sums: slot = block % 32;
 * [.][0] squared error, [.][1] absolute depth error), and the upstream gradients of
 *   loss = rgb_weight * mse + depth_weight * l1
 * are written for nfs_composite_bwd:  g_rgb (N,3) = rgb_weight * 2 (rgb - target) / (3 N),
 *   g_depth (N)|NULL = depth_weight * sign(depth - target_depth) / N.
 * Everything else as nfs_composite_fwd (out_weights may be NULL when no resampling pass follows). */
int nfs_composite_loss_fwd(const float *rgb, const float *density, const float *z_vals,
                           const float *rays_d, const float *noise, float noise_std,
                           const float *target_rgb, const float *target_depth,
                           float rgb_weight, float depth_weight,
                           int64_t n_rays, int32_t n_samples,
                           int32_t white_bkgd, int32_t packed,
                           float *out_rgb, float *out_depth, float *out_weights,
                           float *g_rgb, float *g_depth, double *loss_sums,
                           void *stream);

/* Backward of the above by recomputation from the inputs (nothing is saved by
 * the forward).  Closed form of the autograd graph of nerf_mlp.py:181-215:
 *   G_i       = g_rgb.rgb_i + g_depth z_i + g_w_i - white_bkgd * sum_c g_rgb_c
 *   d rgb_i   = w_i g_rgb
 *   d alpha_i = G_i T_i - (sum_{k>i} G_k w_k) / q_i ,  q_i = (1-alpha_i)+1e-10
 *   d dens_i  = d alpha_i * dists_i * exp(-relu(dens_i) dists_i) * [dens_i > 0]
 * g_rgb (N,3); g_depth (N)|NULL; g_weights (N,S)|NULL.
 * d_rgb (N,S,3) and d_density (N,S)  — or, packed, d_rgb = d_rgb_sigma (N,S,4)
 * and d_density NULL.  z_vals / rays_d never receive gradients (no reference
 * caller asks for them). */
int nfs_composite_bwd(const float *rgb, const float *density, const float *z_vals,
                      const float *rays_d, const float *noise, float noise_std,
                      const float *g_rgb, const float *g_depth, const float *g_weights,
                      int64_t n_rays, int32_t n_samples,
                      int32_t white_bkgd, int32_t packed,
                      float *d_rgb, float *d_density,
                      void *stream);

/* ------------------------------------------------------------------------- *
 * K2 — positional encoding
 *   replaces PositionalEncoding.forward   src/models/positional_encoding.py:20-33
 *        and nerf_mlp.PositionalEncoding.forward  src/models/nerf_mlp.py:17-33
 *   out[p] = [x (if include_input), sin(x f_0), cos(x f_0), ..., cos(x f_{L-1})]
 *   x (P,D) -> out (P, D*(2L+include_input)); freqs: L floats (device).
 * ------------------------------------------------------------------------- */
int nfs_posenc_fwd(const float *x, const float *freqs, int64_t n_points,
                   int32_t dim, int32_t n_freqs, int32_t include_input,
                   float *out, void *stream);

/* ------------------------------------------------------------------------- *
 * K4a — stratified sampling along rays
 *   replaces models.ray_sampler.sample_points_along_rays  src/models/ray_sampler.py:47-61
 *        and utils.ray_utils.sample_points_along_rays     src/utils/ray_utils.py:55-84
 *   The S-entry tables z_base / lower / upper are the reference's own
 *   torch.linspace arithmetic (ray_utils.py:59-76) evaluated once on the host
 *   and uploaded; the kernel does, with un-contracted fp32 mul/add,
 *     z   = t_rand ? lower + (upper-lower)*t_rand : z_base      ray_utils.py:79
 *     pts = rays_o + rays_d * z                                ray_utils.py:82
 *   rays_o, rays_d (N,3); t_rand (N,S)|NULL; z_out (N,S); pts_out (N,S,3)|NULL.
 * ------------------------------------------------------------------------- */
int nfs_sample_stratified(const float *rays_o, const float *rays_d,
                          const float *z_base, const float *lower, const float *upper,
                          const float *t_rand, int64_t n_rays, int32_t n_samples,
                          float *z_out, float *pts_out, void *stream);

/* ------------------------------------------------------------------------- *
 * K4b — hierarchical (inverse-CDF) sampling
 *   replaces utils.ray_utils.hierarchical_sampling   src/utils/ray_utils.py:101-143
 *   z_vals (N,M+1) coarse depths (bin edges), weights (N,M), u (N,Ni) with row
 *   stride u_stride (0 = one row broadcast to all rays, the perturb=False
 *   linspace table).  cdf_in (N,M+1)|NULL: when given it is used instead of
 *   the kernel's own cdf (kernel-level parity against the oracle's cdf).
 *   Outputs: z_out (N,M+1+Ni) sorted ascending; pts_out (N,M+1+Ni,3)|NULL;
 *   debug (all optional): cdf_out (N,M+1), idx_out (N,Ni) int64 searchsorted
 *   indices, samples_out (N,Ni) the un-merged fine samples.
 *   One warp per ray; cdf = fp64-accumulated cumsum rounded per entry (ATen CPU
 *   semantics), searchsorted(right=True), un-contracted interpolation, bitonic
 *   merge-sort in shared memory.  M+1+Ni <= 4096.
 * ------------------------------------------------------------------------- */
int nfs_sample_hierarchical(const float *rays_o, const float *rays_d,
                            const float *z_vals, const float *weights,
                            const float *u, int64_t u_stride, const float *cdf_in,
                            int64_t n_rays, int32_t n_bins, int32_t n_importance,
                            float *z_out, float *pts_out,
                            float *cdf_out, int64_t *idx_out, float *samples_out,
                            void *stream);

/* ------------------------------------------------------------------------- *
 * K5 — feature-conditioning gather (the step before the conditioned MLP)
 *   replaces utils.ray_utils.project_points_to_image           src/utils/ray_utils.py:176-210
 *        and SpatialDINOFeatures.sample_features_at_points     src/models/dino_feature_model.py:114-148
 *   points (P,3) world coordinates; pose_inv (4,4) row-major = inverse of the camera-to-world pose
 *   (ray_utils.py:192); features (Hp,Wp,C) fp32 = the (1,Hp,Wp,C) feature map | NULL.
 *     cam = [p,1] . pose_inv^T; x = cam_x / (cam_z + 1e-8) * focal + W/2 (y likewise with H);
 *     points_2d = (x / W) * 2 - 1 ; depths = cam_z ; valid = cam_z > 0
 *     sampled = F.grid_sample(features, points_2d, bilinear, zeros padding, align_corners=False)
 *   Outputs (each optional): points_2d (P,2), depths (P), valid (P) bytes, sampled (P,C).
 *   pose_inv == NULL: `points` already holds normalised image coordinates (P,2); only `sampled` is produced.
 * ------------------------------------------------------------------------- */
int nfs_project_gather(const float *points, const float *pose_inv, float focal, int32_t H, int32_t W,
                       const float *features, int32_t Hp, int32_t Wp, int32_t C, int64_t n_points,
                       float *points_2d, float *depths, unsigned char *valid, float *sampled, void *stream);

/* ------------------------------------------------------------------------- *
 * K3 — NeRF MLP dense layers on tcgen05 tensor cores
 *   replaces the nn.Linear(+ReLU / sigmoid) chains of
 *     nerf_model.NeRFMLP.forward               src/models/nerf_model.py:16-24
 *     nerf_mlp.NeRFWithDINO.forward            src/models/nerf_mlp.py:134-158
 *     NeRFDINOFusion.forward                   src/models/dino_feature_model.py:175-197
 *   and the dgrad / wgrad GEMMs of their autograd backward.
 *
 * nfs_linear_bf16: Y[P,N] = act( X[P,K] . W[N,K]^T + bias[N] ) (* relu mask)
 *   X [P,K], W [N,K] bf16 row-major (K contiguous; K % 64 == 0, K <= 320;
 *   N % 32 == 0, N <= 256 - operands are zero-padded to these shapes), bias fp32
 *   |NULL, fp32 accumulation in TMEM.  act: 0 none, 1 relu, 2 sigmoid on columns
 *   0..2 only ([rgb|sigma] head of nerf_model.py:22-24), 3 sigmoid, 5 softmax over columns 0..1
 *   (the 2-way gate of dino_feature_model.py:165-170,188).
 *   relu_mask_src (bf16 [P,N])|NULL: result *= [relu_mask_src > 0] (ReLU backward
 *   fused into the dgrad epilogue; call with W^T as the weight).
 *   Outputs: y_bf16 [P,N] with row pitch y_pitch elements (0 = N; a wider pitch writes a column
 *   block of a concatenated operand, e.g. [features | encoded directions] of nerf_mlp.py:83)
 *   and/or y_f32 [P,out_cols] (first out_cols columns).
 * ------------------------------------------------------------------------- */
int nfs_linear_bf16(const void *x_bf16, const void *w_bf16, const float *bias,
                    const void *relu_mask_src,
                    int64_t n_points, int32_t k_dim, int32_t n_dim, int32_t act,
                    int32_t out_cols, void *y_bf16, int64_t y_pitch, float *y_f32, void *stream);

/* nfs_wgrad_bf16: D[m,n] += sum_p U[p,m] * V[p,n]   (fp32 red.add into dw[m*ld_m + n*ld_n]),
 *   the wgrad GEMM of Linear backward: dW[n_out,k_in] = sum_p dY[p,n_out] X[p,k_in]
 *   (autograd of nerf_model.py:18 / nerf_mlp.py:60-66,82-84).  U [P,M] (row pitch u_pitch),
 *   V [P,N] (row pitch v_pitch) bf16 row-major; M in {128,256}, N % 64 == 0, N <= 256.
 *   Lanes own consecutive m: pass the operand whose column index is contiguous in dW as U
 *   (ld_m = 1) when its width allows.  colsum (fp32)|NULL also receives += the column sums of
 *   V (colsum_of_v != 0, N entries) or U (M entries): the bias gradient db[n] = sum_p dY[p,n].
 *   Only entries m < m_valid, n < n_valid are written (0 = all): the operands are zero-padded
 *   to the tile shapes, the destination is the un-padded parameter gradient.
 *   The destination must be zeroed (or hold the running gradient) beforehand. */
int nfs_wgrad_bf16(const void *u_bf16, int64_t u_pitch, const void *v_bf16, int64_t v_pitch,
                   int64_t n_points, int32_t m_dim, int32_t n_dim,
                   int32_t m_valid, int32_t n_valid,
                   float *dw, int64_t ld_m, int64_t ld_n,
                   float *colsum, int32_t colsum_of_v, void *stream);

/* nfs_wgrad_multi_bf16: several independent nfs_wgrad_bf16 jobs in ONE launch (the SMs are divided among the jobs
 *   in proportion to their operand bytes).  For models with many small layers (NeRFWithDINO: 21 weight gradients
 *   of a few ten thousand points each) a launch per layer is mostly fixed cost.  jobs: HOST array, read during the
 *   call; every field as the nfs_wgrad_bf16 argument of the same name. */
typedef struct nfs_wgrad_job {
  const void *u_bf16; int64_t u_pitch;
  const void *v_bf16; int64_t v_pitch;
  int64_t n_points;
  int32_t m_dim, n_dim, m_valid, n_valid;
  float *dw; int64_t ld_m, ld_n;
  float *colsum; int32_t colsum_of_v;
} nfs_wgrad_job;
int nfs_wgrad_multi_bf16(const nfs_wgrad_job *jobs, int32_t n_jobs, void *stream);

/* nfs_mlp_chain_rays: nfs_mlp_chain_points[_train] with the SAMPLER fused in as well - point p is sample p % n_samples
 *   of ray p / n_samples at rays_o + rays_d * z_vals[p] (ray_utils.py:82 / ray_sampler.py:58, evaluated un-contracted by
 *   the warps that build the first layer's operand), so the (P,3) positions never exist in HBM.  x_bf16_out / save_bf16 /
 *   relu_bits_out all NULL: inference; x_bf16_out + save_bf16 (+ relu_bits_out): forward of a training step. */
int nfs_mlp_chain_rays(const float *rays_o, const float *rays_d, const float *z_vals, int64_t n_rays,
                       int32_t n_samples, float freq0, int32_t n_octaves, int32_t n_layers,
                       const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                       const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16, void *x_bf16_out,
                       void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                       int32_t out_cols, void *stream);

/* nfs_render_fused_fwd: the whole render path of a ray batch in one call (train.py:188-242 + ray_utils.py:86-143 for the
 *   plain model): stratified depths (nfs_sample_stratified's tables / draws) -> sampler + encoding + MLP in one kernel
 *   (nfs_mlp_chain_rays) -> compositing; with n_importance > 0 also inverse-CDF resampling on weights[..., :-1] (u,
 *   u_stride as nfs_sample_hierarchical) -> MLP over the merged n_coarse + n_importance depths -> compositing.
 *   model: the packed operands of the fused chain (HOST struct, fields as the nfs_mlp_chain arguments of the same name).
 *   Workspace / outputs (device, caller-owned): z_coarse [N,Sc], raw_coarse [N,Sc,4], weights_coarse [N,Sc] (NULL allowed
 *   when n_importance == 0), bin_weights [N,Sc-1] scratch, rgb_coarse [N,3], depth_coarse [N]|NULL; fine pass: z_fine
 *   [N,Sf], raw_fine [N,Sf,4], weights_fine [N,Sf]|NULL, rgb_fine [N,3], depth_fine [N]|NULL (Sf = Sc + n_importance).
 *   Sample positions, encodings and hidden activations never reach HBM. */
typedef struct nfs_chain_model {
  int32_t n_layers;
  const int32_t *k_dims, *n_dims, *acts, *row0;
  const void *w_stack_bf16;
  int32_t w_rows;
  const void *bias_terms_bf16;
  float freq0;
  int32_t n_octaves;
} nfs_chain_model;
int nfs_render_fused_fwd(const nfs_chain_model *model, const float *rays_o, const float *rays_d, int64_t n_rays,
                         int32_t n_coarse, const float *z_base, const float *lower, const float *upper,
                         const float *t_rand, int32_t n_importance, const float *u, int64_t u_stride,
                         int32_t white_bkgd, float *z_coarse, float *raw_coarse, float *weights_coarse,
                         float *bin_weights, float *rgb_coarse, float *depth_coarse, float *z_fine,
                         float *raw_fine, float *weights_fine, float *rgb_fine, float *depth_fine, void *stream);

/* nfs_composite_bwd_dy: the packed backward of nfs_composite_bwd with the derivative of the MLP head folded in
 *   (nerf_model.py:22-24: rgb = sigmoid(.), sigma raw): instead of d(rgb_sigma) [N,S,4] fp32 it writes, per sample p,
 *   dy[p*dy_pitch + 0..2] = d_rgb * rgb * (1 - rgb) and dy[p*dy_pitch + 3] = d_sigma as bf16 - columns 0..3 of the
 *   zero-padded [P, dy_pitch] operand of the MLP's dgrad chain and head weight gradient (the other columns are left
 *   untouched: keep them zero).  Bit-identical to nfs_composite_bwd followed by nfs_act_grad_bf16 (act 2). */
int nfs_composite_bwd_dy(const float *rgb_sigma, const float *z_vals, const float *rays_d, const float *g_rgb,
                         const float *g_depth, const float *g_weights, int64_t n_rays, int32_t n_samples,
                         int32_t white_bkgd, void *dy_bf16, int64_t dy_pitch, void *stream);

/* nfs_mlp_backward_fused: the backward pass of a fused MLP chain in ONE persistent launch - the dgrad chain
 *   (arguments as the backward use of nfs_mlp_chain below: X = dy_bf16 [n_points, k_dims[0]], transposed weights in
 *   reverse order, act 4 = ReLU backward from relu_bits_in, every layer's output stored to dys_bf16
 *   [n_layers, save_rows_per_layer, n_dims[0]]) runs on `producer_pairs` CTA pairs (0 = library default) while the
 *   remaining CTAs compute the weight / bias gradients `jobs` (nfs_wgrad_job, as nfs_wgrad_multi_bf16), consuming the
 *   chain's output quad by quad (512 rows) through L2 as soon as it has been stored.  job_waits[i] != 0 marks a job
 *   whose operands are (partly) produced by this launch's chain (they are read only after the chain has published the
 *   rows); 0 = operands complete before the launch.  quad_flags: >= 2 * ceil(n_points / 512) uint32 of device scratch
 *   (zeroed by the call).  Replaces autograd's backward of nerf_model.NeRFMLP (src/models/nerf_model.py:16-24):
 *   dX_l = (dY_l W_l) * ReLU'(h_l), dW_l = dY_l^T h_{l-1}, db_l = sum_p dY_l, for all points of a training step. */
int nfs_mlp_backward_fused(const void *dy_bf16, int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                           const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                           const void *wt_stack_bf16, int32_t w_rows, const void *relu_bits_in,
                           int64_t bits_rows_per_layer, const int32_t *mask_idx, void *dys_bf16,
                           int64_t save_rows_per_layer, const nfs_wgrad_job *jobs, int32_t n_jobs,
                           const int32_t *job_waits, uint32_t *quad_flags, int32_t producer_pairs, void *stream);

/* nfs_render_fused_fwd_train: nfs_render_fused_fwd for a TRAINING step (train.py:244-292 up to the loss): the same
 *   path - stratified depths -> sampler + encoding + MLP -> compositing [-> resampling -> MLP -> compositing] - with
 *   (a) the chain kernel's training form: the encoded first-layer operand, every hidden activation and the ReLU sign
 *   bits of both passes go to the step's arenas (state: HOST struct; pass i occupies the rows from row0[i], a multiple
 *   of 128, of arenas with `rows_per_layer` rows per layer plane) for nfs_render_fused_bwd, and (b) the rgb MSE of every
 *   pass (train.py:36-44) evaluated in the compositing epilogue (nfs_composite_loss_fwd): g_rgb_* (N,3) receive
 *   d loss / d rgb_map of the pass, loss_out[0] = rgb_weight * (mse_coarse + mse_fine), [1] = mse_coarse, [2] = mse_fine
 *   (one small kernel sums the 32 fp64 partial slots of each pass; loss_sums: 128 doubles of scratch, zeroed here).
 *   Other arguments as nfs_render_fused_fwd (no weights_fine: nothing resamples after the fine pass). */
typedef struct nfs_chain_train {
  void *x_bf16;              /* [rows, k_dims[0]] bf16: receives the encoded operand of the first layer */
  void *save_bf16;           /* [n_layers - 1, rows_per_layer, n_dims[0]] bf16: hidden activations */
  void *relu_bits;           /* [n_layers - 1, rows_per_layer, 8] uint32: ReLU sign bits */
  int64_t rows_per_layer;
  int64_t row0[2];           /* first arena row of the coarse / fine pass */
} nfs_chain_train;
int nfs_render_fused_fwd_train(const nfs_chain_model *model, const nfs_chain_train *state, const float *rays_o,
                               const float *rays_d, int64_t n_rays, int32_t n_coarse, const float *z_base,
                               const float *lower, const float *upper, const float *t_rand, int32_t n_importance,
                               const float *u, int64_t u_stride, int32_t white_bkgd, const float *target_rgb,
                               float rgb_weight, float *z_coarse, float *raw_coarse, float *weights_coarse,
                               float *bin_weights, float *rgb_coarse, float *depth_coarse, float *g_rgb_coarse,
                               float *z_fine, float *raw_fine, float *rgb_fine, float *depth_fine, float *g_rgb_fine,
                               double *loss_sums, float *loss_out, void *stream);

/* nfs_render_fused_bwd: the backward pass of a training step's render path in one call - the counterpart of
 *   nfs_render_fused_fwd for NeRFDINOTrainer.train_step (train.py:244-292: loss.backward() through render_rays):
 *   for every compositing pass of the step (coarse, fine) the compositing backward with the MLP head's derivative
 *   folded in (nfs_composite_bwd_dy) writes columns 0..3 of that pass's rows of the dgrad chain's bf16 operand
 *   dy_bf16 [n_points, dy_pitch] (rows between a pass's last point and its 128-row boundary are zeroed here), then
 *   the whole MLP backward runs as ONE persistent launch (nfs_mlp_backward_fused over all n_points rows: dgrad chain +
 *   every weight / bias gradient).  passes[i].row0: first row of the pass in the step's arenas (multiple of 128; the
 *   passes must not overlap); mlp: HOST struct, fields as the nfs_mlp_backward_fused arguments of the same name.
 *   The d(rgb_sigma) tensors, the hidden gradients' round trip to DRAM and every per-layer launch are gone. */
typedef struct nfs_render_pass {
  const float *rgb_sigma;   /* [n_rays, n_samples, 4] packed network output of the pass */
  const float *z_vals;      /* [n_rays, n_samples] */
  const float *g_rgb;       /* [n_rays, 3] d loss / d rgb_map of the pass */
  const float *g_depth;     /* [n_rays] d loss / d depth_map | NULL */
  int32_t n_samples;
  int64_t row0;
} nfs_render_pass;
typedef struct nfs_chain_backward {
  int32_t n_layers;
  const int32_t *k_dims, *n_dims, *acts, *row0;
  const void *wt_stack_bf16;
  int32_t w_rows;
  const void *relu_bits_in;
  int64_t bits_rows_per_layer;
  const int32_t *mask_idx;
  void *dys_bf16;
  int64_t save_rows_per_layer;
  const nfs_wgrad_job *jobs;
  int32_t n_jobs;
  const int32_t *job_waits;
  uint32_t *quad_flags;
  int32_t producer_pairs;
} nfs_chain_backward;
int nfs_render_fused_bwd(const nfs_render_pass *passes, int32_t n_passes, const float *rays_d, int64_t n_rays,
                         int32_t white_bkgd, void *dy_bf16, int64_t dy_pitch, int64_t n_points,
                         const nfs_chain_backward *mlp, void *stream);

/* nfs_mlp_chain: a whole chain of dense layers in ONE launch (fused multi-layer MLP):
 *   h_0 = X;  h_{l+1} = act_l( h_l . W_l^T + b_l ),  l = 0 .. n_layers-1
 *   Forward use: replaces nerf_model.NeRFMLP.forward (src/models/nerf_model.py:16-24) and the equal-width
 *   sub-chains of NeRFWithDINO (nerf_mlp.py:134-158).  Backward use: the dgrad chain of their autograd
 *   backward (X = dL/d(last pre-activation), W_l = transposed weights in reverse order, act 4 = ReLU backward).
 *   Runs on CTA pairs (tcgen05 cta_group::2); activations stay in shared memory / TMEM between layers; weights
 *   are TMA-streamed from one stacked bf16 tensor w_stack [w_rows, 256] (layer l = rows row0[l] .. row0[l]+N_l,
 *   columns 0..K_l, zero padded); biases stacked the same way as bias_terms_bf16 [w_rows, 8] (row row0[l] + n =
 *   nfs_bias_terms_bf16 of b_l[n]: the bias is added by the tensor core, one K = 16 MMA per tile and layer) or NULL.
 *   K_l, N_l multiples of 64 in [64,256], K_l == N_{l-1}.
 *   acts[l]: 0 none, 1 relu, 2 sigmoid on columns 0..2, 3 sigmoid, 5 softmax over columns 0..1 (output head with
 *   out_cols == 2 only: the gate of dino_feature_model.py:188), 4 ReLU backward: multiply by the sign bit
 *   relu_bits_in[mask_idx[l]][p][n] (layers >= 128 wide).
 *   ReLU mask bits: uint32 [layers, rows_per_layer, 8] = 256 bits per point (1 = the pre-activation was positive);
 *   in word w bit 15 - j (j < 16) is column 32w + 2j and bit 31 - j is column 32w + 2j + 1.  relu_bits_out (forward chain of a training step, NULL
 *   otherwise) receives the bits of every act-1 layer's output, 32 bytes per point instead of the 512-byte
 *   activation row the backward would otherwise re-read; rows per layer = save_rows_per_layer.
 *   out_f32 != NULL: the LAST layer is an output head whose first out_cols columns are written
 *   as fp32 [P,out_cols].  save_bf16 != NULL: [n_saved, save_rows_per_layer, max_l N_l] bf16 (saved layers are
 *   at least 128 wide; a narrower layer fills the first N_l columns of its rows) receives
 *   every non-head layer's output by TMA store (n_saved = n_layers - 1 with a head, else
 *   n_layers); *_rows_per_layer >= P rounded up to 128. */
int nfs_mlp_chain(const void *x_bf16, int64_t n_points, int32_t n_layers,
                  const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts,
                  const int32_t *row0, const void *w_stack_bf16, int32_t w_rows,
                  const void *bias_terms_bf16,
                  const void *relu_bits_in, int64_t bits_rows_per_layer, const int32_t *mask_idx,
                  void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer,
                  float *out_f32, int32_t out_cols, void *stream);

/* nfs_mlp_chain_points: the same chain with K2 fused in (inference): the chain input is the positional encoding
 *   of points (P,3) fp32 - [x, sin(x f_0), cos(x f_0), ..., cos(x f_{L-1}), 0], f_k = freq0 * 2^k, L = n_octaves
 *   <= 10 - computed by the epilogue warps straight into the first layer's shared-memory operand (K_0 must be
 *   64), so neither the fp32 encoding (positional_encoding.py:27-33) nor its bf16 copy ever exists in HBM.
 *   The last layer is the fp32 output head (out_f32 [P,out_cols]); nothing is saved. */
int nfs_mlp_chain_points(const float *points, float freq0, int32_t n_octaves, int64_t n_points,
                         int32_t n_layers, const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts,
                         const int32_t *row0, const void *w_stack_bf16, int32_t w_rows,
                         const void *bias_terms_bf16, float *out_f32, int32_t out_cols, void *stream);

/* nfs_mlp_chain_points_train: nfs_mlp_chain_points for the forward of a TRAINING step: the chain kernel encodes the
 *   points itself, stores the encoded bf16 operand into x_bf16_out [n_points rounded up to 128, 64] (the first
 *   layer's weight gradient reads it), and saves activations / ReLU sign bits exactly like nfs_mlp_chain
 *   (save_bf16, relu_bits_out, save_rows_per_layer).  Rows between n_points and the 128-row boundary receive the
 *   encoding of the origin. */
int nfs_mlp_chain_points_train(const float *points, float freq0, int32_t n_octaves, int64_t n_points,
                               int32_t n_layers, const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts,
                               const int32_t *row0, const void *w_stack_bf16, int32_t w_rows,
                               const void *bias_terms_bf16, void *x_bf16_out, void *save_bf16, void *relu_bits_out,
                               int64_t save_rows_per_layer, float *out_f32, int32_t out_cols, void *stream);

/* ------------------------------------------------------------------------- *
 * K2 fused with the operand cast of the first dense layer
 *   (positional_encoding.py:27-33 / nerf_mlp.py:24-33, the torch.cat with DINO features of
 *   dino_feature_model.py:182,195 and the fp32->bf16 cast feeding nn.Linear):
 *   out row (bf16, k_pad entries) = [ enc(x) * scale_enc | extra * scale_extra | 0 ... ]
 *   enc(x) = [x, sin(x f_0), cos(x f_0), ...] (D*(2L+1)); extra (P,E)|NULL; scale_*|NULL are
 *   the per-point attention gates of dino_feature_model.py:191-192, read at scale_*[p*scale_stride]
 *   (stride 2 = the two columns of the (P,2) softmax output).  Rows are written at
 *   out + p*out_pitch (0 = k_pad), so the block can be a column range of a wider operand.
 *   pow2_bands != 0 asserts freqs[k] == freqs[0] * 2^k (the log-sampled bands both reference encoders
 *   use by default): sin/cos of the higher octaves then come from the double-angle recurrence
 *   (error <= ~2e-4 at 12 octaves, below the bf16 rounding of the output) instead of 2L sincosf.
 * ------------------------------------------------------------------------- */
int nfs_posenc_bf16(const float *x, const float *freqs, const float *extra,
                    const float *scale_enc, const float *scale_extra, int32_t scale_stride,
                    int64_t n_points, int32_t dim, int32_t n_freqs, int32_t extra_dim,
                    int32_t k_pad, int64_t out_pitch, int32_t pow2_bands, void *out_bf16, void *stream);

/* Backward of that gate: dc_bf16 (P rows, pitch dc_pitch) = dL/dc' for c' = [enc(x) g0 | extra g1],
 * gate (P,2) = softmax output -> dlogits_bf16 [P,n_pad] (columns 0..1, rest zero):
 *   dg0 = <dc'[:enc_w], enc(x)>, dg1 = <dc'[enc_w:], extra>, dlogit_i = g_i (dg_i - sum_k g_k dg_k)
 *   (autograd of dino_feature_model.py:188-195; enc(x) is recomputed). */
int nfs_gate_bwd_bf16(const float *x, const float *freqs, const float *extra, const float *gate,
                      const void *dc_bf16, int64_t dc_pitch, int64_t n_points, int32_t dim, int32_t n_freqs,
                      int32_t extra_dim, int32_t n_pad, void *dlogits_bf16, void *stream);

/* fp32 master weight [n_dim,k_dim] -> block (row0,col0) of the zero-padded bf16 operands
 * w_bf16 [n_pad,k_pad] (forward) and wt_bf16 [k_pad,n_pad] (dgrad); either may be NULL.
 * Parameters keep the reference's names / [out,in] fp32 layout;
 * these are the cached operand copies refreshed after an optimizer step. */
int nfs_pack_linear_bf16(const float *w, int32_t n_dim, int32_t k_dim, int32_t n_pad, int32_t k_pad,
                         int32_t row0, int32_t col0, void *w_bf16, void *wt_bf16, void *stream);

/* nfs_pack_stack: one launch refreshes every bf16 operand of a fused chain (nfs_mlp_chain) from the fp32 master
 * parameters after an optimizer step.  table: device int64 [n_entries, 8], row =
 *   [pointer to the fp32 parameter, n_dim, k_dim, w_row0, w_col0, wt_row0 (-1: none), wt_col0, bias_row0]
 *   k_dim > 0: weight [n_dim,k_dim] -> w_stack[(w_row0+n), w_col0+k] and wt_stack[(wt_row0+k), wt_col0+n]
 *   k_dim = 0: bias [n_dim] -> bias_terms[bias_row0+n] (see nfs_bias_terms_bf16)
 * w_stack / wt_stack are [rows, 256] bf16, zero where no parameter lands; max_elems = the largest n_dim*k_dim. */
int nfs_pack_stack(const void *table, int32_t n_entries, int32_t max_elems, void *w_stack_bf16, void *wt_stack_bf16,
                   void *bias_terms_bf16, void *stream);

/* nfs_pack_table: the general form of nfs_pack_stack for models whose bf16 operands live in many tensors
 * (NeRFWithDINO).  table: device int64 [n_entries, 10], row =
 *   [src (fp32 pointer), n_dim, k_dim (0: vector), src_pitch, dst (pointer), dst_pitch, row0, col0, mode, 0]
 *   mode 0: bf16 dst[(row0+n)*dst_pitch + col0+k] = src[n*src_pitch + k]     mode 1: the transpose,
 *           dst[(row0+k)*dst_pitch + col0+n];   mode 2: fp32 dst[row0+n] = src[n];   mode 3: bias terms
 *           (nfs_bias_terms_bf16) dst[row0+n].  max_elems = the largest n_dim*max(k_dim,1). */
int nfs_pack_table(const void *table, int32_t n_entries, int32_t max_elems, void *stream);

/* nfs_g3_operand: K5 + K2 as the PRODUCER of the conditioned model's first operand: row p of
 *   out_bf16 [P,k_pad] = [x, sin(x f_0), cos(x f_0), ..., cos(x f_{L-1}) | bilinear features of the projected point | 0]
 *   (dino_feature_model.py:182 on top of ray_utils.py:176-210 + dino_feature_model.py:114-148 + nerf_mlp.py:24-33), i.e.
 *   nfs_project_gather followed by nfs_posenc_bf16 without the (P,C) fp32 features in between.  Arguments as the
 *   arguments of the same name of those two entries; k_pad % 8 == 0, 3 (2 n_freqs + 1) + C <= k_pad <= 320. */
int nfs_g3_operand(const float *points, const float *pose_inv, float focal, int32_t H, int32_t W,
                   const float *features, int32_t Hp, int32_t Wp, int32_t C, const float *freqs, int32_t n_freqs,
                   int32_t pow2_bands, int64_t n_points, int32_t k_pad, int64_t out_pitch, void *out_bf16, void *stream);

/* nfs_gate_scale_bf16 / nfs_gate_bwd_operand: the softmax gate of NeRFDINOFusion (dino_feature_model.py:188-195) and
 *   its backward, on the bf16 operand c = [enc(x) | f | 0] [P,k_pad] that the first fusion layer consumed:
 *     out[p,j] = c[p,j] * (j < enc_w ? gate[p,0] : gate[p,1])
 *     dlogits[p,0:2] = g_i (dg_i - (g0 dg0 + g1 dg1)),  dg0 = <dc[p,0:enc_w], c[p,0:enc_w]>,  dg1 = <dc[p,enc_w:width], c[p,enc_w:width]>
 *   (bf16 [P,n_pad], zero padded).  The fp32 variants that recompute enc(x) are nfs_posenc_bf16 (with its scale
 *   arguments) and nfs_gate_bwd_bf16. */
int nfs_gate_scale_bf16(const void *c_bf16, int64_t c_pitch, const float *gate, int64_t n_points, int32_t enc_w,
                        int32_t k_pad, void *out_bf16, int64_t out_pitch, void *stream);
int nfs_gate_bwd_operand(const void *c_bf16, int64_t c_pitch, const float *gate, const void *dc_bf16, int64_t dc_pitch,
                         int64_t n_points, int32_t enc_w, int32_t width, int32_t n_pad, void *dlogits_bf16, void *stream);

/* nfs_scatter_add_table: several strided fp32 block adds in ONE launch.  table: device int64 [n_entries, 8], row =
 *   [src, dst, rows, cols, src_ld_r, src_ld_c, dst_ld, clear]:  dst[r*dst_ld + c] += src[r*src_ld_r + c*src_ld_c];
 *   clear != 0 zeroes the source entries afterwards (an accumulator that is reused every step needs no fill).
 * Moves the lane-contiguous accumulators of the weight-gradient kernels (the output head's rows stacked in one
 * operand, the first layer's gradient transposed) into the parameter gradients' own layout (autograd of
 * nerf_model.py:16-24); max_elems = the largest rows*cols.  Blocks must not overlap. */
int nfs_scatter_add_table(const void *table, int32_t n_entries, int64_t max_elems, void *stream);

/* fp32 bias[n] -> terms_bf16 [n, 8] bf16, row i = [hi, mid, lo, 0, 0, 0, 0, 0] with hi + mid + lo = bias[i] to
 * ~2^-24 relative: the bias operand of nfs_mlp_chain (nn.Linear's "+ b", nerf_model.py:16-24, added on the
 * tensor core as ones[128x16] . terms^T).  Refreshed with the packed weights after an optimizer step. */
int nfs_bias_terms_bf16(const float *bias, int32_t n, void *terms_bf16, void *stream);

/* dY (bf16 [P,n_pad], zero padded) = g_out * act'(out) for the fp32 network outputs
 * out, g_out [P,n_cols]: act 0 identity, 1 relu, 2 sigmoid on columns 0..2 (nerf_model.py:22-24),
 * 3 sigmoid (nerf_mlp.py:80).  Rows at dy + p*dy_pitch (0 = n_pad).  First step of the MLP backward. */
int nfs_act_grad_bf16(const float *out, const float *g_out, int64_t n_points, int32_t n_cols,
                      int32_t act, int32_t n_pad, int64_t dy_pitch, void *dy_bf16, void *stream);

/* ------------------------------------------------------------------------- *
 * Fused Adam / AdamW step over one flat fp32 buffer
 *   replaces optim.Adam(...).step()   src/training/train.py:114-118,286   (decoupled = 0)
 *        and optim.AdamW(...).step()  src/training/train_multiscale.py:61 (decoupled = 1)
 *   grad is multiplied by grad_scale first (1/world_size after a sum-allreduce).
 * ------------------------------------------------------------------------- */
int nfs_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                  int64_t n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int32_t step, float grad_scale, int32_t decoupled,
                  void *stream);

/* Same update with the step count and learning rate resident on the device, so that the launch can
 * be replayed from a CUDA graph: *step_counter is incremented first, state (3 floats) =
 * [1 - beta1^t, sqrt(1 - beta2^t), lr]; the first two are rewritten by the call, state[2] (lr,
 * MultiStepLR of train.py:120-124) is owned by the host. */
int nfs_adam_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                      int64_t n, float beta1, float beta2, float eps, float weight_decay,
                      int32_t *step_counter, float *state, float grad_scale, int32_t decoupled,
                      void *stream);

/* ------------------------------------------------------------------------- *
 * K6  Ray generation + batch assembly
 *   models.ray_sampler.get_rays (src/models/ray_sampler.py:4-30), utils.ray_utils.get_rays
 *   (src/utils/ray_utils.py:4-37) and the per-batch gather of train.py:272-278
 *   (rays_o_full.view(-1,3)[idx], rays_d_full.view(-1,3)[idx], target_rgb_full.view(-1,3)[idx]) in one launch:
 *     pixel p = pix_idx[r] (row-major, p = j * width + i; pix_idx NULL: p = r and n_rays = height * width)
 *     dirs   = [(i - width*0.5) / focal, -(j - height*0.5) / focal, -1]
 *     rays_d[r] = sum(dirs * c2w[:3,:3], -1)   rays_o[r] = c2w[:3,3]   target[r] = image[p]
 *   Bit-exact with the reference's CPU arithmetic (true division, products rounded one by one, left-to-right sum).
 *   c2w: device pointer to a row-major (3|4, 4) pose with c2w_row_stride floats per row; image (height*width,3)|NULL;
 *   rays_o / rays_d / target (n_rays,3), each may be NULL (target and image go together).  An index outside
 *   [0, height*width) yields NaN rays (torch would raise).
 * ------------------------------------------------------------------------- */
int nfs_rays_generate(int32_t height, int32_t width, float focal, const float *c2w, int32_t c2w_row_stride,
                      const int64_t *pix_idx, int64_t n_rays, const float *image,
                      float *rays_o, float *rays_d, float *target, void *stream);

/* ---------------------------------------------------------------------------
 * Data-parallel weight update over NVLink peer memory (the reference is single-device:
 * this replaces optimizer.step() of src/training/train.py:286 for a ray-sharded batch).
 *   Every rank owns one exchange buffer = [n_grad fp32 gradient | flags] allocated by nfs_dp_alloc (plain cudaMalloc,
 *   zeroed), exports it with nfs_dp_ipc_export (64-byte CUDA IPC handle, sent to the peers by the host) and maps its
 *   peers' buffers with nfs_dp_ipc_open.  peer_bases: HOST array of `world` device pointers, entry `rank` = the local
 *   buffer.  The model's weight-gradient kernels accumulate into the local buffer's gradient part.
 *   nfs_dp_adam_step: ONE kernel that (1) tells the peers this rank's gradient is complete, (2) waits for theirs,
 *   (3) sums the `world` gradient buffers element-wise, reading the peers' over NVLink, and applies the Adam / AdamW
 *   update of nfs_adam_step_dev to param / exp_avg / exp_avg_sq (replicated weights, every rank computes the same
 *   update), (4) tells the peers it has finished reading.  epoch_dev / cta_counter: one zero-initialised uint32 each.
 *   nfs_dp_wait_readers: waits until every peer has finished reading this rank's gradient of the last exchange; run it
 *   before the gradient buffer is zeroed for the next step.  Both are stream-ordered, capturable in a CUDA graph, and
 *   bounded (a missing peer raises a CUDA error after ~2 s).
 * ------------------------------------------------------------------------- */
uint64_t nfs_dp_flags_offset(int64_t n_grad);
uint64_t nfs_dp_buffer_bytes(int64_t n_grad);
int nfs_dp_alloc(int64_t n_grad, void **base_out);
int nfs_dp_free(void *base);
int nfs_dp_ipc_export(void *base, void *handle64);
int nfs_dp_ipc_open(const void *handle64, void **base_out);
int nfs_dp_ipc_close(void *base);
int nfs_dp_wait_readers(void *const *peer_bases, int32_t world, int32_t rank, int64_t n_grad,
                        const uint32_t *epoch_dev, void *stream);
int nfs_dp_adam_step(float *param, void *const *peer_bases, int32_t world, int32_t rank, float *exp_avg,
                     float *exp_avg_sq, int64_t n, float beta1, float beta2, float eps, float weight_decay,
                     int32_t *step_counter, float *state, float grad_scale, int32_t decoupled,
                     uint32_t *epoch_dev, uint32_t *cta_counter, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NFS_B200_H_ */
