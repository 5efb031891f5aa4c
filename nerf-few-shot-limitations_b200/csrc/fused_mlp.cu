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

// Boxes of the stacked weight tensor and of the saved activations differ from make_tmap_bf16's
// default only in their row count (64 resp. 32).
// points != NULL: in-kernel encoding; x_bf16 is then NULL (inference) or the [rows128, 64] bf16 buffer that RECEIVES the
// encoded operand (training forward).
static int launch_chain(const char *fn, const void *x_bf16, const float *points, float freq0, int n_octaves,
                        int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                        const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                        const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16,
                        const void *relu_bits_in, int64_t bits_rows_per_layer, const int32_t *mask_idx,
                        void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                        int32_t out_cols, void *stream) {
  if (n_points < 0 || n_layers < 2 || n_layers > kFmMaxLayers) return fail_arg(fn, NFS_E_BADARG, "need 2..12 layers");
  if (n_points == 0) return 0;
  if ((!x_bf16 && !points) || !k_dims || !n_dims || !acts || !row0 || !w_stack_bf16 || (!out_f32 && !save_bf16))
    return fail_arg(fn, NFS_E_BADARG, "null pointer");
  FusedArgs a{};
  a.P = n_points; a.n_layers = n_layers; a.has_bias = bias_terms_bf16 != nullptr; a.out = out_f32; a.out_cols = out_cols;
  a.save = save_bf16 != nullptr; a.save_rows = save_rows_per_layer;
  a.head = out_f32 != nullptr;
  a.points = points; a.freq0 = freq0; a.n_octaves = n_octaves;
  a.dbg = g_fm_debug;
  a.trace = g_fm_trace;
  a.bits_in = (const uint32_t *)relu_bits_in; a.bits_rows = bits_rows_per_layer;
  a.bits_out = (uint32_t *)relu_bits_out;
  for (int l = 0; l < n_layers; ++l) {
    a.K[l] = k_dims[l]; a.N[l] = n_dims[l]; a.act[l] = acts[l]; a.row0[l] = row0[l];
    a.mask_idx[l] = mask_idx ? mask_idx[l] : 0;
    if (a.act[l] == 4 && (!relu_bits_in || !mask_idx || a.mask_idx[l] < 0))
      return fail_arg(fn, NFS_E_BADARG, "act 4 (ReLU backward) needs relu_bits_in and mask_idx");
    if ((a.act[l] == 4 || (a.act[l] == 1 && relu_bits_out)) && a.N[l] < 128 && !(out_f32 && l == n_layers - 1))
      return fail_arg(fn, NFS_E_UNSUPPORTED, "ReLU sign bits need layers at least 128 wide");
    if (a.act[l] < 0 || a.act[l] > 4) return fail_arg(fn, NFS_E_BADARG, "act must be 0..4");
    if (a.K[l] % 64 || a.K[l] <= 0 || a.K[l] > 256 || a.N[l] % 64 || a.N[l] <= 0 || a.N[l] > 256 ||
        a.row0[l] < 0 || a.row0[l] + a.N[l] > w_rows)
      return fail_arg(fn, NFS_E_UNSUPPORTED, "layer dims must be multiples of 64 in [64,256] and fit the weight stack");
    if (l > 0 && a.K[l] != a.N[l - 1]) return fail_arg(fn, NFS_E_BADARG, "layer l input width != layer l-1 output width");
  }
  if (a.head && (out_cols <= 0 || out_cols > a.N[n_layers - 1])) return fail_arg(fn, NFS_E_BADARG, "out_cols out of range");
  const long long rows128 = ((n_points + 127) / 128) * 128;
  if (a.save && save_rows_per_layer < rows128)
    return fail_arg(fn, NFS_E_BADARG, "save_rows_per_layer must be >= n_points rounded up to 128");
  if (relu_bits_in && bits_rows_per_layer < rows128)
    return fail_arg(fn, NFS_E_BADARG, "bits_rows_per_layer must be >= n_points rounded up to 128");
  if (relu_bits_out && (relu_bits_in || save_rows_per_layer < rows128))
    return fail_arg(fn, NFS_E_BADARG, "relu_bits_out belongs to a forward chain (no relu_bits_in) with save_rows_per_layer >= rows");
  const int n_saved = a.head ? n_layers - 1 : n_layers;
  if (a.save)
    for (int l = 0; l < n_saved; ++l)
      if (a.N[l] != a.N[0]) return fail_arg(fn, NFS_E_UNSUPPORTED, "saved activations need equal layer widths");

  CUtensorMap tx{}, tw{}, ts{}, tb{};
  int rc = 0;
  if (points != nullptr) {
    if (a.K[0] != 64 || n_octaves < 1 || n_octaves > 10)
      return fail_arg(fn, NFS_E_UNSUPPORTED, "in-kernel encoding needs a 64-wide first layer and 1..10 octaves");
    if (x_bf16 != nullptr) {
      if (!a.save) return fail_arg(fn, NFS_E_BADARG, "the encoded operand is only stored by a training forward (save_bf16)");
      a.x_save = 1;
      rc = tc::make_tmap_bf16(&tx, x_bf16, (uint64_t)rows128, 64, 64, 32, fn);
      if (rc) return rc;
    }
  } else {
    rc = tc::make_tmap_bf16(&tx, x_bf16, (uint64_t)n_points, (uint64_t)a.K[0], (uint64_t)a.K[0], 128, fn);
    if (rc) return rc;
  }
  rc = tc::make_tmap_bf16(&tw, w_stack_bf16, (uint64_t)w_rows, 256, 256, 64, fn);
  if (rc) return rc;
  if (bias_terms_bf16 != nullptr) {
    rc = tc::make_tmap_rows16(&tb, bias_terms_bf16, (uint64_t)w_rows, 128, fn);
    if (rc) return rc;
  }
  if (a.save) {
    rc = tc::make_tmap_bf16(&ts, save_bf16, (uint64_t)(save_rows_per_layer * n_saved), (uint64_t)a.N[0],
                            (uint64_t)a.N[0], 32, fn);
    if (rc) return rc;
  }
  const size_t smem = 2 * kActBytes + kWStages * kWStage + 256 + 256 + 128 * 16;   // tiles, weight ring, barriers, ones, bias operand
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
  const long long n_quads = ((n_points + 127) / 128 + 3) / 4;
  const long long max_pairs = sms / 2;
  const unsigned grid = 2u * (unsigned)(n_quads < max_pairs ? n_quads : max_pairs);   // whole CTA pairs
  if (relu_bits_in != nullptr)
    fused_mlp_kernel<true, false><<<grid, kFmThreads, smem, (cudaStream_t)stream>>>(tx, tw, ts, tb, a);
  else if (a.save || relu_bits_out)
    fused_mlp_kernel<false, true><<<grid, kFmThreads, smem, (cudaStream_t)stream>>>(tx, tw, ts, tb, a);
  else
    fused_mlp_kernel<false, false><<<grid, kFmThreads, smem, (cudaStream_t)stream>>>(tx, tw, ts, tb, a);
  return check_launch(fn);
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
