// nfs_render_fused_fwd — the render path of one ray batch behind ONE C call:
//   stratified depths -> [positions + sin/cos encoding + MLP] -> alpha compositing
//   [-> inverse-CDF resampling -> positions + encoding + MLP -> alpha compositing]
// i.e. NeRFDINOTrainer.render_rays (/root/reference/src/training/train.py:188-242) for the plain model, with the
// hierarchical pass of utils.ray_utils.hierarchical_sampling (src/utils/ray_utils.py:86-143).
//
// What never reaches HBM: the sample positions (P,3) (the sampler's `o + d z` is evaluated by the chain kernel's operand
// warps, fused_mlp_body.cuh), the (P,63) encodings and every hidden activation.  What does: the depths z (P floats, also an
// input of the compositing and of the resampling), and the network's packed [r,g,b,sigma] rows (P,4): the compositing
// scan runs over the S samples of a ray, which only coincide with the chain kernel's 128-row tiles when S divides 128 -
// the fine pass (S = 192) does not - so K1 stays its own (HBM-bound, 2 % of the frame) launch.
#include <cuda_bf16.h>

#include "nfs_common.cuh"

using namespace nfs;

extern "C" int nfs_render_fused_fwd(const nfs_chain_model *m, const float *rays_o, const float *rays_d, int64_t n_rays,
                                    int32_t n_coarse, const float *z_base, const float *lower, const float *upper,
                                    const float *t_rand, int32_t n_importance, const float *u, int64_t u_stride,
                                    int32_t white_bkgd, float *z_coarse, float *raw_coarse, float *weights_coarse,
                                    float *bin_weights, float *rgb_coarse, float *depth_coarse, float *z_fine,
                                    float *raw_fine, float *weights_fine, float *rgb_fine, float *depth_fine,
                                    void *stream) {
  const char *fn = "nfs_render_fused_fwd";
  if (!m || !rays_o || !rays_d || !z_base || !z_coarse || !raw_coarse || !rgb_coarse || n_rays < 0 || n_coarse < 2 ||
      n_importance < 0)
    return fail_arg(fn, NFS_E_BADARG, "null pointer / bad sizes");
  if (n_importance > 0 && (!u || !weights_coarse || !bin_weights || !z_fine || !raw_fine || !rgb_fine))
    return fail_arg(fn, NFS_E_BADARG, "the fine pass needs u, weights_coarse, bin_weights, z_fine, raw_fine, rgb_fine");
  if (n_rays == 0) return 0;
  int rc = nfs_sample_stratified(nullptr, nullptr, z_base, lower, upper, t_rand, n_rays, n_coarse, z_coarse, nullptr, stream);
  if (rc) return rc;
  rc = nfs_mlp_chain_rays(rays_o, rays_d, z_coarse, n_rays, n_coarse, m->freq0, m->n_octaves, m->n_layers, m->k_dims,
                          m->n_dims, m->acts, m->row0, m->w_stack_bf16, m->w_rows, m->bias_terms_bf16, nullptr, nullptr,
                          nullptr, 0, raw_coarse, 4, stream);
  if (rc) return rc;
  rc = nfs_composite_fwd(raw_coarse, nullptr, z_coarse, rays_d, nullptr, 0.f, n_rays, n_coarse, white_bkgd, 1, rgb_coarse,
                         depth_coarse, weights_coarse, stream);
  if (rc || n_importance == 0) return rc;
  // weights[..., :-1]: the M = S - 1 bins between the S coarse depths (hierarchical_sampling's calling convention)
  const int M = n_coarse - 1;
  cudaError_t e = cudaMemcpy2DAsync(bin_weights, (size_t)M * 4, weights_coarse, (size_t)n_coarse * 4, (size_t)M * 4,
                                    (size_t)n_rays, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(fn, e);
  rc = nfs_sample_hierarchical(rays_o, rays_d, z_coarse, bin_weights, u, u_stride, nullptr, n_rays, M, n_importance, z_fine,
                               nullptr, nullptr, nullptr, nullptr, stream);
  if (rc) return rc;
  const int S = n_coarse + n_importance;
  rc = nfs_mlp_chain_rays(rays_o, rays_d, z_fine, n_rays, S, m->freq0, m->n_octaves, m->n_layers, m->k_dims, m->n_dims,
                          m->acts, m->row0, m->w_stack_bf16, m->w_rows, m->bias_terms_bf16, nullptr, nullptr, nullptr, 0,
                          raw_fine, 4, stream);
  if (rc) return rc;
  return nfs_composite_fwd(raw_fine, nullptr, z_fine, rays_d, nullptr, 0.f, n_rays, S, white_bkgd, 1, rgb_fine, depth_fine,
                           weights_fine, stream);
}

// nfs_render_fused_bwd — the backward of a training step's render path behind ONE C call: autograd's loss.backward()
// through NeRFDINOTrainer.render_rays (/root/reference/src/training/train.py:188-242, 244-292) for the plain model:
// compositing backward of every pass (closed form, nothing saved, the head's sigmoid / identity derivative folded in,
// written as the bf16 operand of the MLP backward), then dgrad chain + all weight gradients in one persistent launch.
extern "C" int nfs_render_fused_bwd(const nfs_render_pass *passes, int32_t n_passes, const float *rays_d, int64_t n_rays,
                                    int32_t white_bkgd, void *dy_bf16, int64_t dy_pitch, int64_t n_points,
                                    const nfs_chain_backward *m, void *stream) {
  const char *fn = "nfs_render_fused_bwd";
  if (!passes || n_passes < 1 || !rays_d || !dy_bf16 || !m || n_rays < 0 || n_points < 0 || dy_pitch < 4)
    return fail_arg(fn, NFS_E_BADARG, "null pointer / bad sizes");
  if (n_rays == 0 || n_points == 0) return 0;
  for (int i = 0; i < n_passes; ++i) {
    const nfs_render_pass &p = passes[i];
    const long long P = (long long)n_rays * p.n_samples, pad = (P + 127) / 128 * 128;
    if (!p.rgb_sigma || !p.z_vals || !p.g_rgb || p.n_samples < 2 || p.row0 < 0 || (p.row0 & 127) || p.row0 + pad > (n_points + 127) / 128 * 128)
      return fail_arg(fn, NFS_E_BADARG, "a pass needs rgb_sigma, z_vals, g_rgb, n_samples >= 2 and a 128-aligned row range inside n_points");
    __nv_bfloat16 *dy = static_cast<__nv_bfloat16 *>(dy_bf16) + p.row0 * dy_pitch;
    if (pad > P) {                     // zero-gradient rows up to the pass's 128-row boundary
      cudaError_t e = cudaMemsetAsync(dy + P * dy_pitch, 0, (size_t)(pad - P) * (size_t)dy_pitch * 2, (cudaStream_t)stream);
      if (e != cudaSuccess) return fail_cuda(fn, e);
    }
    int rc = nfs_composite_bwd_dy(p.rgb_sigma, p.z_vals, rays_d, p.g_rgb, p.g_depth, nullptr, n_rays, p.n_samples, white_bkgd,
                                  dy, dy_pitch, stream);
    if (rc) return rc;
  }
  return nfs_mlp_backward_fused(dy_bf16, n_points, m->n_layers, m->k_dims, m->n_dims, m->acts, m->row0, m->wt_stack_bf16,
                                m->w_rows, m->relu_bits_in, m->bits_rows_per_layer, m->mask_idx, m->dys_bf16,
                                m->save_rows_per_layer, m->jobs, m->n_jobs, m->job_waits, m->quad_flags, m->producer_pairs,
                                stream);
}

// nfs_render_fused_fwd_train — the forward of a training step's render path behind ONE C call
// (NeRFDINOTrainer.train_step up to the loss, /root/reference/src/training/train.py:244-292 with render_rays :188-242 and
// the rgb MSE of :36-44 on both passes): what nfs_render_fused_fwd does, with the chain kernel saving what the backward
// needs into the step's arenas and the loss evaluated in the compositing epilogue.
namespace nfs {
namespace {
// loss_sums: [2 passes][32 slots][2] fp64 partial sums of nfs_composite_loss_fwd -> out = {total, mse_coarse, mse_fine}
__global__ void loss_finalize_kernel(const double *__restrict__ sums, long long n_rays, int n_passes, float rgb_weight,
                                     float *__restrict__ out) {
  const int lane = threadIdx.x;                     // one warp
  float mse[2] = {0.f, 0.f};
  for (int p = 0; p < n_passes; ++p) {
    double v = sums[(p * 32 + lane) * 2];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    mse[p] = (float)(v / (3.0 * (double)n_rays));   // fp64 sum and quotient, rounded once (ops.composite_loss does the same)
  }
  if (lane == 0) {
    out[1] = mse[0];
    out[2] = mse[1];
    out[0] = n_passes == 2 ? __fadd_rn(__fmul_rn(rgb_weight, mse[0]), __fmul_rn(rgb_weight, mse[1])) : __fmul_rn(rgb_weight, mse[0]);
  }
}
}  // namespace
}  // namespace nfs

extern "C" int nfs_render_fused_fwd_train(const nfs_chain_model *m, const nfs_chain_train *st, const float *rays_o,
                                          const float *rays_d, int64_t n_rays, int32_t n_coarse, const float *z_base,
                                          const float *lower, const float *upper, const float *t_rand,
                                          int32_t n_importance, const float *u, int64_t u_stride, int32_t white_bkgd,
                                          const float *target_rgb, float rgb_weight, float *z_coarse, float *raw_coarse,
                                          float *weights_coarse, float *bin_weights, float *rgb_coarse,
                                          float *depth_coarse, float *g_rgb_coarse, float *z_fine, float *raw_fine,
                                          float *rgb_fine, float *depth_fine, float *g_rgb_fine, double *loss_sums,
                                          float *loss_out, void *stream) {
  const char *fn = "nfs_render_fused_fwd_train";
  if (!m || !st || !rays_o || !rays_d || !z_base || !target_rgb || !z_coarse || !raw_coarse || !rgb_coarse || !g_rgb_coarse ||
      !loss_sums || !loss_out || n_rays < 0 || n_coarse < 2 || n_importance < 0)
    return fail_arg(fn, NFS_E_BADARG, "null pointer / bad sizes");
  if (!st->x_bf16 || !st->save_bf16 || !st->relu_bits || (st->row0[0] & 127) || (st->row0[1] & 127))
    return fail_arg(fn, NFS_E_BADARG, "the arenas (x_bf16, save_bf16, relu_bits) and 128-aligned pass rows are required");
  if (n_importance > 0 && (!u || !weights_coarse || !bin_weights || !z_fine || !raw_fine || !rgb_fine || !g_rgb_fine))
    return fail_arg(fn, NFS_E_BADARG, "the fine pass needs u, weights_coarse, bin_weights, z_fine, raw_fine, rgb_fine, g_rgb_fine");
  if (n_rays == 0) return 0;
  const int n_passes = n_importance > 0 ? 2 : 1;
  cudaError_t e = cudaMemsetAsync(loss_sums, 0, 128 * sizeof(double), (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(fn, e);
  const size_t k0 = (size_t)m->k_dims[0], hp = (size_t)m->n_dims[0];
  auto arena = [&](int pass, void **x, void **save, void **bits) {
    const size_t r0 = (size_t)st->row0[pass];
    *x = static_cast<uint8_t *>(st->x_bf16) + r0 * k0 * 2;
    *save = static_cast<uint8_t *>(st->save_bf16) + r0 * hp * 2;
    *bits = static_cast<uint8_t *>(st->relu_bits) + r0 * 8 * 4;
  };
  void *x, *save, *bits;
  int rc = nfs_sample_stratified(nullptr, nullptr, z_base, lower, upper, t_rand, n_rays, n_coarse, z_coarse, nullptr, stream);
  if (rc) return rc;
  arena(0, &x, &save, &bits);
  rc = nfs_mlp_chain_rays(rays_o, rays_d, z_coarse, n_rays, n_coarse, m->freq0, m->n_octaves, m->n_layers, m->k_dims,
                          m->n_dims, m->acts, m->row0, m->w_stack_bf16, m->w_rows, m->bias_terms_bf16, x, save, bits,
                          st->rows_per_layer, raw_coarse, 4, stream);
  if (rc) return rc;
  rc = nfs_composite_loss_fwd(raw_coarse, nullptr, z_coarse, rays_d, nullptr, 0.f, target_rgb, nullptr, rgb_weight, 0.f,
                              n_rays, n_coarse, white_bkgd, 1, rgb_coarse, depth_coarse,
                              n_importance > 0 ? weights_coarse : nullptr, g_rgb_coarse, nullptr, loss_sums, stream);
  if (rc) return rc;
  if (n_importance > 0) {
    const int M = n_coarse - 1;
    e = cudaMemcpy2DAsync(bin_weights, (size_t)M * 4, weights_coarse, (size_t)n_coarse * 4, (size_t)M * 4, (size_t)n_rays,
                          cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    rc = nfs_sample_hierarchical(rays_o, rays_d, z_coarse, bin_weights, u, u_stride, nullptr, n_rays, M, n_importance, z_fine,
                                 nullptr, nullptr, nullptr, nullptr, stream);
    if (rc) return rc;
    const int S = n_coarse + n_importance;
    arena(1, &x, &save, &bits);
    rc = nfs_mlp_chain_rays(rays_o, rays_d, z_fine, n_rays, S, m->freq0, m->n_octaves, m->n_layers, m->k_dims, m->n_dims,
                            m->acts, m->row0, m->w_stack_bf16, m->w_rows, m->bias_terms_bf16, x, save, bits,
                            st->rows_per_layer, raw_fine, 4, stream);
    if (rc) return rc;
    rc = nfs_composite_loss_fwd(raw_fine, nullptr, z_fine, rays_d, nullptr, 0.f, target_rgb, nullptr, rgb_weight, 0.f, n_rays,
                                S, white_bkgd, 1, rgb_fine, depth_fine, nullptr, g_rgb_fine, nullptr, loss_sums + 64, stream);
    if (rc) return rc;
  }
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_sums, (long long)n_rays, n_passes, rgb_weight, loss_out);
  return check_launch(fn);
}
