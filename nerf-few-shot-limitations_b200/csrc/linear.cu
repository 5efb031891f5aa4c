// K3 (dense layers) — Y[P,N] = act( X[P,K] . W[N,K]^T + b ) on tcgen05 tensor cores.
//
// Replaces the nn.Linear(+ReLU / sigmoid) steps of
//   nerf_model.NeRFMLP.forward       /root/reference/src/models/nerf_model.py:16-24
//   NeRFWithDINO.forward             /root/reference/src/models/nerf_mlp.py:134-158
//   NeRFDINOFusion.forward           /root/reference/src/models/dino_feature_model.py:175-197
// and, fed with W^T and the ReLU mask, the dgrad step of their autograd backward.
//
// Weight-stationary persistent kernel, one CTA per SM:
//   * the whole bf16 weight matrix (<= 320 x 256, <= 160 KB) is TMA-loaded ONCE per CTA into
//     shared memory as K/64 slabs of [N x 64] in the canonical K-major SWIZZLE_128B layout;
//   * 128-point activation tiles stream through a ring of [128 x 64] TMA stages;
//   * one elected thread issues tcgen05.mma (M=128, N=N, K=16, bf16 -> fp32) into one of two
//     TMEM accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * four epilogue warps read the accumulator with tcgen05.ld (each thread = one point),
//     add the bias, apply ReLU / sigmoid / the ReLU-backward mask and store bf16 (next layer's
//     operand, also the saved activation) and/or fp32.
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..9 = epilogue: two warps per TMEM lane quadrant, each
// taking every other 32-column chunk (at the few ten thousand points of NeRFWithDINO a CTA sees one or two tiles, and
// the four-warp epilogue - 8 chunks x (tcgen05.ld, 32 bias loads, activation, mask, 4 stores) per thread - was the
// longest part of the launch); the bias is staged in shared memory once per CTA.
// Per layer the kernel is HBM-bound (reads 2K, writes 2N bytes per point against 2KN flop).
#include "tc_common.cuh"

namespace nfs {
namespace {

using namespace tc;

constexpr int kLinThreads = 320;      // warps 0 / 1 = TMA / MMA, 2..9 = epilogue (two per TMEM lane quadrant)
constexpr int kTileM = 128;
constexpr int kSlabBytesX = kTileM * 128;   // [128 x 64] bf16

struct LinearArgs {
  const float *bias;
  const __nv_bfloat16 *mask_src;
  __nv_bfloat16 *y_bf16;
  float *y_f32;
  long long P;
  long long y_pitch;   // row pitch (elements) of y_bf16
  int K, N, act, out_cols;
  int n_stages, tmem_cols;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kLinThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const LinearArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = a.N, ks = a.K >> 6, S = a.n_stages;
  const int w_slab_bytes = N * 128;
  uint8_t *w_smem = smem;
  uint8_t *x_smem = w_smem + ks * w_slab_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(x_smem + S * kSlabBytesX);
  uint64_t *empty = full + S;
  uint64_t *w_full = empty + S;                  // [ks <= 5] weight slab s landed
  uint64_t *tmem_full = w_full + 5;
  uint64_t *tmem_empty = tmem_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
  float *s_bias = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(full) + 256);      // [N] fp32

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_tiles = (a.P + kTileM - 1) / kTileM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 5; ++i) mbar_init(w_full + i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 8); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();                      // everything above is private to the CTA; from here on global memory is touched
  for (int i = threadIdx.x; i < N; i += kLinThreads) s_bias[i] = a.bias != nullptr ? __ldg(a.bias + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                   // TMEM is held: a dependent grid's CTAs may be scheduled behind this one

  if (warp == 0) {
    if (lane == 0) {
      // weights: resident for the whole kernel, one barrier per 64-column slab so that the first tile's MMAs start when
      // the first slab (and the first activation slab, requested right behind it) has landed
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < ks; ++s, ++it) {
          if (tile == (long long)blockIdx.x) {
            mbar_expect_tx(w_full + s, (uint32_t)w_slab_bytes);
            tma_load_2d(w_smem + s * w_slab_bytes, &tmap_w, w_full + s, s * 64, 0);
          }
          const uint32_t stage = it % S, ph = (it / S) & 1;
          mbar_wait(empty + stage, ph ^ 1);
          mbar_expect_tx(full + stage, kSlabBytesX);
          tma_load_2d(x_smem + stage * kSlabBytesX, &tmap_x, full + stage, s * 64, (int)(tile * kTileM));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTileM, N, 0, 0);
      uint32_t it = 0, t = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const uint32_t acc = t & 1, aph = (t >> 1) & 1;
        mbar_wait(tmem_empty + acc, aph ^ 1);     // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)N;
        for (int s = 0; s < ks; ++s, ++it) {
          const uint32_t stage = it % S, ph = (it / S) & 1;
          if (t == 0) mbar_wait(w_full + s, 0);
          mbar_wait(full + stage, ph);
          tc_fence_after();
          const uint32_t xa = smem_u32(x_smem + stage * kSlabBytesX);
          const uint32_t wa = smem_u32(w_smem + s * w_slab_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = umma_desc_sw128(xa + k * 32, 16, 1024);
            const uint64_t bd = umma_desc_sw128(wa + k * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (uint32_t)((s | k) != 0));
          }
          umma_commit(empty + stage);             // stage reusable once these MMAs have read it
        }
        umma_commit(tmem_full + acc);             // accumulator complete
      }
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quadrant this warp may touch
    const int half = (warp - 2) >> 2;             // which of the quadrant's two warps: even or odd 32-column chunks
    const int row_in_tile = q * 32 + lane;
    uint32_t t = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const uint32_t acc = t & 1, aph = (t >> 1) & 1;
      mbar_wait(tmem_full + acc, aph);
      tc_fence_after();
      const long long row = tile * kTileM + row_in_tile;
      const bool ok = row < a.P;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)N;
      for (int c0 = 32 * half; c0 < N; c0 += 64) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (a.bias != nullptr) {
          const float4 *bp = reinterpret_cast<const float4 *>(s_bias + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = bp[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (a.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (a.act == 2) {                  // [sigmoid(rgb) x3 | raw sigma]  nerf_model.py:22-24
          if (c0 == 0) { v[0] = sigmoidf_(v[0]); v[1] = sigmoidf_(v[1]); v[2] = sigmoidf_(v[2]); }
        } else if (a.act == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = sigmoidf_(v[j]);
        } else if (a.act == 5) {                  // 2-way softmax gate  dino_feature_model.py:169,188
          if (c0 == 0) {
            const float mx = fmaxf(v[0], v[1]);
            const float e0 = expf(v[0] - mx), e1 = expf(v[1] - mx);
            const float inv = 1.0f / (e0 + e1);
            v[0] = e0 * inv; v[1] = e1 * inv;
          }
        }
        if (ok) {
          if (a.mask_src != nullptr) {            // ReLU backward: pass where the saved activation is > 0
            const uint4 *mp = reinterpret_cast<const uint4 *>(a.mask_src + row * N + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 m = __ldg(mp + g);
              const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                // bf16 > 0  <=>  sign clear and magnitude non-zero
                if (!((w[h] & 0x7FFFu) != 0 && (w[h] & 0x8000u) == 0)) v[g * 8 + 2 * h] = 0.f;
                if (!((w[h] & 0x7FFF0000u) != 0 && (w[h] & 0x80000000u) == 0)) v[g * 8 + 2 * h + 1] = 0.f;
              }
            }
          }
          if (a.y_bf16 != nullptr) {
            uint4 *yp = reinterpret_cast<uint4 *>(a.y_bf16 + row * a.y_pitch + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              yp[g] = make_uint4(pack_bf16x2(v[8 * g], v[8 * g + 1]), pack_bf16x2(v[8 * g + 2], v[8 * g + 3]),
                                 pack_bf16x2(v[8 * g + 4], v[8 * g + 5]), pack_bf16x2(v[8 * g + 6], v[8 * g + 7]));
          }
          if (a.y_f32 != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < a.out_cols) a.y_f32[row * a.out_cols + c0 + j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_sm_count <= 0)
      g_sm_count = 148;
  }
  return g_sm_count;
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_linear_bf16(const void *x_bf16, const void *w_bf16, const float *bias, const void *relu_mask_src,
                               int64_t n_points, int32_t k_dim, int32_t n_dim, int32_t act, int32_t out_cols,
                               void *y_bf16, int64_t y_pitch, float *y_f32, void *stream) {
  const char *fn = "nfs_linear_bf16";
  if (n_points < 0 || k_dim <= 0 || n_dim <= 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!x_bf16 || !w_bf16 || (!y_bf16 && !y_f32)) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (k_dim % 64 != 0 || n_dim % 32 != 0 || n_dim > 256 || k_dim > 320)
    return fail_arg(fn, NFS_E_UNSUPPORTED, "need K % 64 == 0, K <= 320, N % 32 == 0, N <= 256 (pad the operands)");
  if (act < 0 || act > 5 || act == 4) return fail_arg(fn, NFS_E_BADARG, "act must be 0..3 or 5");
  if (y_pitch == 0) y_pitch = n_dim;
  if (y_pitch < n_dim || (y_pitch & 7)) return fail_arg(fn, NFS_E_BADARG, "y_pitch must be >= N and a multiple of 8");
  if (y_f32 && (out_cols <= 0 || out_cols > n_dim)) return fail_arg(fn, NFS_E_BADARG, "out_cols must be in [1, N]");
  if ((y_bf16 && !aligned16(y_bf16)) || (relu_mask_src && !aligned16(relu_mask_src)))
    return fail_arg(fn, NFS_E_ALIGN, "bf16 tensors must be 16-byte aligned");

  CUtensorMap tx, tw;
  int rc = tc::make_tmap_bf16(&tx, x_bf16, (uint64_t)n_points, (uint64_t)k_dim, (uint64_t)k_dim, kTileM, fn);
  if (rc) return rc;
  rc = tc::make_tmap_bf16(&tw, w_bf16, (uint64_t)n_dim, (uint64_t)k_dim, (uint64_t)k_dim, (uint32_t)n_dim, fn);
  if (rc) return rc;

  LinearArgs a{};
  a.bias = bias; a.mask_src = (const __nv_bfloat16 *)relu_mask_src;
  a.y_bf16 = (__nv_bfloat16 *)y_bf16; a.y_f32 = y_f32;
  a.P = n_points; a.y_pitch = y_pitch; a.K = k_dim; a.N = n_dim; a.act = act; a.out_cols = out_cols;
  const int w_bytes = (k_dim / 64) * n_dim * 128;
  const int budget = 227 * 1024 - 1024 - 256 - 1024 - w_bytes;        // alignment slack, barriers, bias
  int stages = budget / kSlabBytesX;
  if (stages > 8) stages = 8;
  if (stages < 2) return fail_arg(fn, NFS_E_TOOLARGE, "weights leave no room for the activation ring");
  a.n_stages = stages;
  int cols = 32;
  while (cols < 2 * n_dim) cols <<= 1;
  a.tmem_cols = cols;
  const size_t smem = 1024 + (size_t)w_bytes + (size_t)stages * kSlabBytesX + 256 + 1024;
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  const long long n_tiles = (n_points + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < sm_count() ? n_tiles : sm_count());
  return launch_dep(fn, linear_kernel, dim3(grid), dim3(kLinThreads), smem, (cudaStream_t)stream, tx, tw, a);
}
