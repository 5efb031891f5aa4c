// K3 (fused multi-layer MLP, forward and dgrad): kernel instantiations + C ABI.  The kernel body lives in
// fused_mlp_body.cuh (it is also the producer half of the merged backward kernel, backward.cu).
#include "fused_mlp_body.cuh"

namespace nfs {
namespace {

template <bool kMasked, bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFmThreads, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_save, const __grid_constant__ CUtensorMap tmap_b,
                 const FusedArgs a) {
  chain_body<kMasked, kTrain>(&tmap_x, &tmap_w, &tmap_save, &tmap_b, a, blockIdx.x >> 1, gridDim.x >> 1, nullptr);
}

}  // namespace
}  // namespace nfs

using namespace nfs;

static int g_fm_debug = 0;
static unsigned long long *g_fm_trace = nullptr;
extern "C" void nfs_set_debug_flags(int32_t flags) { g_fm_debug = flags; }
extern "C" void nfs_set_debug_trace(void *buf) { g_fm_trace = (unsigned long long *)buf; }

static int launch_prepared(const char *fn, FusedArgs a, const CUtensorMap &tx, const CUtensorMap &tw, const CUtensorMap &ts,
                           const CUtensorMap &tb, bool train, void *stream, bool dgrad = false) {
  a.dbg = g_fm_debug;
  a.trace = g_fm_trace;
  const size_t smem = kChainSmemBytes;
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(fused_mlp_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fused_mlp_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fused_mlp_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_quads = ((a.P + 127) / 128 + 3) / 4;
  const long long max_pairs = sms / 2;
  const unsigned grid = 2u * (unsigned)(n_quads < max_pairs ? n_quads : max_pairs);   // whole CTA pairs
  const dim3 g(grid), b(kFmThreads);
  if (dgrad) return launch_dep(fn, fused_mlp_kernel<true, false>, g, b, smem, (cudaStream_t)stream, tx, tw, ts, tb, a);
  if (train) return launch_dep(fn, fused_mlp_kernel<false, true>, g, b, smem, (cudaStream_t)stream, tx, tw, ts, tb, a);
  return launch_dep(fn, fused_mlp_kernel<false, false>, g, b, smem, (cudaStream_t)stream, tx, tw, ts, tb, a);
}


static int launch_chain(const char *fn, const void *x_bf16, const float *points, float freq0, int n_octaves,
                        int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                        const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                        const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16,
                        const void *relu_bits_in, int64_t bits_rows_per_layer, const int32_t *mask_idx,
                        void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                        int32_t out_cols, void *stream) {
  FusedArgs a{};
  CUtensorMap tx{}, tw{}, ts{}, tb{};
  int rc = chain_prepare(fn, x_bf16, points, freq0, n_octaves, n_points, n_layers, k_dims, n_dims, acts, row0, w_stack_bf16,
                         w_rows, bias_terms_bf16, relu_bits_in, bits_rows_per_layer, mask_idx, save_bf16, relu_bits_out,
                         save_rows_per_layer, out_f32, out_cols, &a, &tx, &tw, &ts, &tb);
  if (rc == 1) return 0;
  if (rc) return rc;
  return launch_prepared(fn, a, tx, tw, ts, tb, a.save || relu_bits_out, stream, relu_bits_in != nullptr);
}

extern "C" int nfs_mlp_chain(const void *x_bf16, int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                             const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                             const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16,
                             const void *relu_bits_in, int64_t bits_rows_per_layer, const int32_t *mask_idx,
                             void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                             int32_t out_cols, void *stream) {
  return launch_chain("nfs_mlp_chain", x_bf16, nullptr, 0.f, 0, n_points, n_layers, k_dims, n_dims, acts, row0, w_stack_bf16,
                      w_rows, bias_terms_bf16, relu_bits_in, bits_rows_per_layer, mask_idx, save_bf16, relu_bits_out,
                      save_rows_per_layer, out_f32, out_cols, stream);
}

extern "C" int nfs_mlp_chain_points(const float *points, float freq0, int32_t n_octaves, int64_t n_points,
                                    int32_t n_layers, const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts,
                                    const int32_t *row0, const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16,
                                    float *out_f32, int32_t out_cols, void *stream) {
  if (!points || !out_f32) return fail_arg("nfs_mlp_chain_points", NFS_E_BADARG, "null pointer");
  return launch_chain("nfs_mlp_chain_points", nullptr, points, freq0, n_octaves, n_points, n_layers, k_dims, n_dims, acts, row0,
                      w_stack_bf16, w_rows, bias_terms_bf16, nullptr, 0, nullptr, nullptr, nullptr, 0, out_f32, out_cols, stream);
}

extern "C" int nfs_mlp_chain_points_train(const float *points, float freq0, int32_t n_octaves, int64_t n_points,
                                          int32_t n_layers, const int32_t *k_dims, const int32_t *n_dims,
                                          const int32_t *acts, const int32_t *row0, const void *w_stack_bf16,
                                          int32_t w_rows, const void *bias_terms_bf16, void *x_bf16_out, void *save_bf16,
                                          void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                                          int32_t out_cols, void *stream) {
  const char *fn = "nfs_mlp_chain_points_train";
  if (!points || !x_bf16_out || !save_bf16) return fail_arg(fn, NFS_E_BADARG, "null pointer");
  return launch_chain(fn, x_bf16_out, points, freq0, n_octaves, n_points, n_layers, k_dims, n_dims, acts, row0, w_stack_bf16,
                      w_rows, bias_terms_bf16, nullptr, 0, nullptr, save_bf16, relu_bits_out, save_rows_per_layer, out_f32,
                      out_cols, stream);
}

// Sampler + encoding + chain in one launch: the points are rays_o + rays_d * z_vals, evaluated by the warps that build the
// first layer's operand.  x_bf16_out / save_bf16 / relu_bits_out != NULL: forward of a training step (as
// nfs_mlp_chain_points_train); all NULL: inference.
extern "C" int nfs_mlp_chain_rays(const float *rays_o, const float *rays_d, const float *z_vals, int64_t n_rays,
                                  int32_t n_samples, float freq0, int32_t n_octaves, int32_t n_layers,
                                  const int32_t *k_dims, const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                                  const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16, void *x_bf16_out,
                                  void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                                  int32_t out_cols, void *stream) {
  const char *fn = "nfs_mlp_chain_rays";
  if (!rays_o || !rays_d || !z_vals || !out_f32 || n_rays < 0 || n_samples <= 0)
    return fail_arg(fn, NFS_E_BADARG, "null pointer / bad sizes");
  if ((x_bf16_out != nullptr) != (save_bf16 != nullptr))
    return fail_arg(fn, NFS_E_BADARG, "a training forward needs both x_bf16_out and save_bf16");
  FusedArgs a{};
  CUtensorMap tx{}, tw{}, ts{}, tb{};
  int rc = chain_prepare(fn, x_bf16_out, rays_o, freq0, n_octaves, n_rays * (int64_t)n_samples, n_layers, k_dims, n_dims, acts,
                         row0, w_stack_bf16, w_rows, bias_terms_bf16, nullptr, 0, nullptr, save_bf16, relu_bits_out,
                         save_rows_per_layer, out_f32, out_cols, &a, &tx, &tw, &ts, &tb);
  if (rc == 1) return 0;
  if (rc) return rc;
  a.ray_z = z_vals; a.rays_o = rays_o; a.rays_d = rays_d; a.ray_S = n_samples;
  return launch_prepared(fn, a, tx, tw, ts, tb, relu_bits_out != nullptr || save_bf16 != nullptr, stream);
}
