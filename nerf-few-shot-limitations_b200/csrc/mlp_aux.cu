// K2 fused with the operand cast of the first dense layer, plus the small kernels around
// the tcgen05 GEMMs of K3: weight packing (fp32 master -> zero-padded bf16 W and W^T),
// output-activation backward, and the fused flat Adam step.
//
// Reference lines replaced:
//   positional encoding + cast   /root/reference/src/models/positional_encoding.py:27-33,
//                                /root/reference/src/models/nerf_mlp.py:24-33 (+ torch.cat with the
//                                DINO features, dino_feature_model.py:182,195)
//   sigmoid / relu backward      autograd of nerf_model.py:22-24, nerf_mlp.py:62,80
//   optimizer step               /root/reference/src/training/train.py:114-118,286 (Adam),
//                                /root/reference/src/training/train_multiscale.py:61 (AdamW)
#include "tc_common.cuh"

namespace nfs {
namespace {

// ------------------------------------------------------------------ posenc -> bf16 operand
constexpr int kEncThreads = 256;
constexpr int kEncTile = 64;   // points per CTA

__global__ void __launch_bounds__(kEncThreads)
posenc_bf16_kernel(const float *__restrict__ x, const float *__restrict__ freqs, const float *__restrict__ extra,
                   const float *__restrict__ scale_enc, const float *__restrict__ scale_extra, int scale_stride,
                   long long n_points, int D, int L, int E, int k_pad, long long out_pitch, int pow2_bands,
                   __nv_bfloat16 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __nv_bfloat16 *tile = reinterpret_cast<__nv_bfloat16 *>(s_raw);     // [kEncTile][k_pad]
  const long long p0 = (long long)blockIdx.x * kEncTile;
  const int np = (int)min((long long)kEncTile, n_points - p0);
  const int enc_w = D * (2 * L + 1);

  // zero the padding columns (and everything else; cheap) so the GEMM sees exact zeros
  pdl_trigger();
  for (int t = threadIdx.x; t < np * k_pad / 8; t += kEncThreads) reinterpret_cast<uint4 *>(tile)[t] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  pdl_wait();                                  // (nfs_common.cuh) first global access below

  const int per_pt = L * D;
  if (pow2_bands) {
    // bands f_k = f_0 * 2^k (log sampling, positional_encoding.py:14 / nerf_mlp.py:14): one accurate sincosf per
    // coordinate, then the double-angle recurrence.  Its error doubles per octave (~1e-7 * 2^k <= 2e-4 at
    // L = 12), far below the bf16 rounding (4e-3) of the operand this kernel writes.
    for (int t = threadIdx.x; t < np * D; t += kEncThreads) {
      const int p = t / D, d = t - p * D;
      float s, c;
      sincosf(__fmul_rn(__ldg(x + (p0 + p) * D + d), __ldg(freqs)), &s, &c);
      const float g = scale_enc ? __ldg(scale_enc + (p0 + p) * scale_stride) : 1.f;
      __nv_bfloat16 *row = tile + p * k_pad + D + d;
      for (int k = 0; k < L; ++k) {
        row[2 * k * D] = __float2bfloat16_rn(s * g);
        row[2 * k * D + D] = __float2bfloat16_rn(c * g);
        const float s2 = 2.f * s * c, c2 = 1.f - 2.f * s * s;
        s = s2; c = c2;
      }
    }
  } else
  for (int t = threadIdx.x; t < np * per_pt; t += kEncThreads) {
    const int p = t / per_pt, r = t - p * per_pt;
    const int k = r / D, d = r - k * D;
    const float arg = __fmul_rn(__ldg(x + (p0 + p) * D + d), __ldg(freqs + k));
    float s, c;
    sincosf(arg, &s, &c);
    const float g = scale_enc ? __ldg(scale_enc + (p0 + p) * scale_stride) : 1.f;
    __nv_bfloat16 *row = tile + p * k_pad + D + 2 * k * D;
    row[d] = __float2bfloat16_rn(s * g);
    row[D + d] = __float2bfloat16_rn(c * g);
  }
  for (int t = threadIdx.x; t < np * D; t += kEncThreads) {
    const int p = t / D, d = t - p * D;
    const float g = scale_enc ? __ldg(scale_enc + (p0 + p) * scale_stride) : 1.f;
    tile[p * k_pad + d] = __float2bfloat16_rn(__ldg(x + (p0 + p) * D + d) * g);
  }
  if (extra != nullptr)
    for (int t = threadIdx.x; t < np * E; t += kEncThreads) {
      const int p = t / E, j = t - p * E;
      const float g = scale_extra ? __ldg(scale_extra + (p0 + p) * scale_stride) : 1.f;
      tile[p * k_pad + enc_w + j] = __float2bfloat16_rn(__ldg(extra + (p0 + p) * E + j) * g);
    }
  __syncthreads();
  const uint4 *src = reinterpret_cast<const uint4 *>(tile);
  const int chunks = k_pad / 8;
  for (int t = threadIdx.x; t < np * chunks; t += kEncThreads) {
    const int p = t / chunks, c = t - p * chunks;
    reinterpret_cast<uint4 *>(out + (p0 + p) * out_pitch)[c] = src[t];
  }
}

// ------------------------------------------------------------------ weight packing
// W fp32 [N,K] (row pitch K) -> w16 [n_pad,k_pad] and w16t [k_pad,n_pad], zero padded.
// `row0` places the block at a row offset of the packed matrices (heads stacked in one operand).
__global__ void __launch_bounds__(256)
pack_linear_kernel(const float *__restrict__ w, int N, int K, int n_pad, int k_pad, int row0, int col0,
                   __nv_bfloat16 *__restrict__ w16, __nv_bfloat16 *__restrict__ w16t) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * K) return;
  const int n = idx / K, k = idx - n * K;
  const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(w + idx));
  if (w16) w16[(long long)(row0 + n) * k_pad + col0 + k] = v;
  if (w16t) w16t[(long long)(col0 + k) * n_pad + row0 + n] = v;
}

// ------------------------------------------------------------------ output activation backward
// out, g_out fp32 [P,C] -> dY bf16 [P,n_pad] = g_out * act'(out), zero padded.
//   act 0 identity, 1 relu (out > 0), 2 sigmoid on the first 3 columns, 3 sigmoid on all.
__global__ void __launch_bounds__(256)
act_grad_kernel(const float *__restrict__ out, const float *__restrict__ g_out, long long n_points, int C, int act,
                int n_pad, long long dy_pitch, __nv_bfloat16 *__restrict__ dy) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const int chunks = n_pad / 8;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_points * chunks) return;
  const long long p = t / chunks;
  const int c0 = (int)(t - p * chunks) * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float r = 0.f;
    if (c < C) {
      const float o = __ldg(out + p * C + c), g = __ldg(g_out + p * C + c);
      if (act == 0) r = g;
      else if (act == 1) r = o > 0.f ? g : 0.f;
      else if (act == 2) r = c < 3 ? g * o * (1.f - o) : g;
      else r = g * o * (1.f - o);
    }
    v[j] = r;
  }
  *reinterpret_cast<uint4 *>(dy + p * dy_pitch + c0) =
      make_uint4(tc::pack_bf16x2(v[0], v[1]), tc::pack_bf16x2(v[2], v[3]), tc::pack_bf16x2(v[4], v[5]),
                 tc::pack_bf16x2(v[6], v[7]));
}

// ------------------------------------------------------------------ gate backward
// The fusion block re-weights its own input with a 2-way softmax gate (dino_feature_model.py:188-195):
//   c' = [enc(x) * g0 | extra * g1],  g = softmax(logits).
// Given dc' (bf16, from the dgrad GEMM of the first fusion layer) this produces the gradient of the
// logits, bf16 [P,n_pad] zero padded (the operand of the next dgrad / wgrad GEMMs):
//   dg0 = <dc'[0:enc_w], enc(x)>,  dg1 = <dc'[enc_w:enc_w+E], extra>,
//   dlogit_i = g_i (dg_i - (g0 dg0 + g1 dg1)).
// enc(x) is recomputed, never stored.  One warp per point, lanes across the columns of the row: the dc' row, the extra
// features and the output row are read and written coalesced (one thread per point read its own 256-byte row with 2-byte
// loads, 32 sectors per instruction: 33 us per 32 768 points, all of it L1 wavefronts), each lane evaluates one
// (band, coordinate) pair with an accurate sincosf, and two shuffle trees finish the dot products.
constexpr int kGateBwdWarps = 8;
__global__ void __launch_bounds__(32 * kGateBwdWarps)
gate_bwd_kernel(const float *__restrict__ x, const float *__restrict__ freqs, const float *__restrict__ extra,
                const float *__restrict__ gate, const __nv_bfloat16 *__restrict__ dc, long long dc_pitch,
                long long n_points, int D, int L, int E, int n_pad, __nv_bfloat16 *__restrict__ out) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kGateBwdWarps + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * kGateBwdWarps;
  const int enc_w = D * (2 * L + 1);
  const int k_first = lane / D, d_first = lane - k_first * D;
  for (long long p = warp; p < n_points; p += n_warps) {
    const __nv_bfloat16 *row = dc + p * dc_pitch;
    float dg0 = 0.f, dg1 = 0.f;
    for (int c = lane; c < D; c += 32) dg0 += __bfloat162float(row[c]) * __ldg(x + p * D + c);
    // one (band, coordinate) pair per lane: a single sincosf serves the sin and the cos column (posenc layout: columns
    // D + 2kD + d and D + 2kD + D + d); the pair owned in the first round never changes, so its division is hoisted
    for (int pair = lane, round = 0; pair < L * D; pair += 32, ++round) {
      const int k = round == 0 ? k_first : pair / D;
      const int d = round == 0 ? d_first : pair - k * D;
      float sn, cs;
      sincosf(__fmul_rn(__ldg(x + p * D + d), __ldg(freqs + k)), &sn, &cs);
      const __nv_bfloat16 *e = row + D + 2 * k * D + d;
      dg0 += __bfloat162float(e[0]) * sn + __bfloat162float(e[D]) * cs;
    }
    // image features: up to eight independent loads in flight per lane
    for (int j0 = lane; j0 < E; j0 += 128) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        a[u] = j < E ? __bfloat162float(row[enc_w + j]) : 0.f;
        b[u] = j < E ? __ldg(extra + p * E + j) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dg1 += a[u] * b[u];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dg0 += __shfl_xor_sync(0xffffffffu, dg0, o);
      dg1 += __shfl_xor_sync(0xffffffffu, dg1, o);
    }
    const float g0 = __ldg(gate + 2 * p), g1 = __ldg(gate + 2 * p + 1);
    const float mean = g0 * dg0 + g1 * dg1;
    uint4 *o = reinterpret_cast<uint4 *>(out + p * n_pad);
    for (int c = lane; c < n_pad / 8; c += 32)
      o[c] = c == 0 ? make_uint4(tc::pack_bf16x2(g0 * (dg0 - mean), g1 * (dg1 - mean)), 0, 0, 0) : make_uint4(0, 0, 0, 0);
  }
}

// The same two operations on the bf16 operand c = [enc(x) | f] the first fusion layer consumed, instead of on x and f:
//   gate_scale:       c'[p,j] = c[p,j] * (j < enc_w ? g0 : g1)           (dino_feature_model.py:191-195)
//   gate_bwd_operand: dg0 = <dc'[0:enc_w], c[0:enc_w]>, dg1 = <dc'[enc_w:width], c[enc_w:width]>, dlogits as above
// - no trigonometry, no fp32 feature tensor: both read 2 bytes per element that the forward pass wrote anyway.
__global__ void __launch_bounds__(256)
gate_scale_kernel(const __nv_bfloat16 *__restrict__ c, long long c_pitch, const float *__restrict__ gate, long long n_points,
                  int enc_w, int k_pad, long long out_pitch, __nv_bfloat16 *__restrict__ out) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const int chunks = k_pad >> 3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_points * chunks) return;
  const long long p = t / chunks;
  const int c0 = (int)(t - p * chunks) * 8;
  const float g0 = __ldg(gate + 2 * p), g1 = __ldg(gate + 2 * p + 1);
  const uint4 v = __ldg(reinterpret_cast<const uint4 *>(c + p * c_pitch + c0));
  const uint32_t in[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float lo = __uint_as_float(in[j] << 16), hi = __uint_as_float(in[j] & 0xffff0000u);
    o[j] = tc::pack_bf16x2(lo * (c0 + 2 * j < enc_w ? g0 : g1), hi * (c0 + 2 * j + 1 < enc_w ? g0 : g1));
  }
  *reinterpret_cast<uint4 *>(out + p * out_pitch + c0) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(32 * kGateBwdWarps)
gate_bwd_operand_kernel(const __nv_bfloat16 *__restrict__ c, long long c_pitch, const float *__restrict__ gate,
                        const __nv_bfloat16 *__restrict__ dc, long long dc_pitch, long long n_points, int enc_w, int width,
                        int n_pad, __nv_bfloat16 *__restrict__ out) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * kGateBwdWarps + (threadIdx.x >> 5);
  const long long n_warps = (long long)gridDim.x * kGateBwdWarps;
  const int chunks = (width + 7) >> 3;                     // both rows are zero beyond `width` up to a multiple of 8
  for (long long p = warp; p < n_points; p += n_warps) {
    float dg0 = 0.f, dg1 = 0.f;
    for (int ch = lane; ch < chunks; ch += 32) {
      const uint4 a = __ldg(reinterpret_cast<const uint4 *>(c + p * c_pitch) + ch);
      const uint4 b = __ldg(reinterpret_cast<const uint4 *>(dc + p * dc_pitch) + ch);
      const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t0 = __uint_as_float(av[j] << 16) * __uint_as_float(bv[j] << 16);
        const float t1 = __uint_as_float(av[j] & 0xffff0000u) * __uint_as_float(bv[j] & 0xffff0000u);
        const int col = ch * 8 + 2 * j;
        if (col < enc_w) dg0 += t0; else if (col < width) dg1 += t0;
        if (col + 1 < enc_w) dg0 += t1; else if (col + 1 < width) dg1 += t1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dg0 += __shfl_xor_sync(0xffffffffu, dg0, o);
      dg1 += __shfl_xor_sync(0xffffffffu, dg1, o);
    }
    const float g0 = __ldg(gate + 2 * p), g1 = __ldg(gate + 2 * p + 1);
    const float mean = g0 * dg0 + g1 * dg1;
    uint4 *o = reinterpret_cast<uint4 *>(out + p * n_pad);
    for (int ch = lane; ch < n_pad / 8; ch += 32)
      o[ch] = ch == 0 ? make_uint4(tc::pack_bf16x2(g0 * (dg0 - mean), g1 * (dg1 - mean)), 0, 0, 0) : make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------ Adam / AdamW over a flat buffer
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
            int decoupled, const float *__restrict__ state) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (state != nullptr) { bc1 = __ldg(state); bc2_sqrt = __ldg(state + 1); lr = __ldg(state + 2); }
  float grad = g[i] * gscale, w = p[i];
  if (wd != 0.f) {
    if (decoupled) w *= 1.f - lr * wd;      // AdamW
    else grad += wd * w;                    // Adam (L2)
  }
  const float mi = b1 * m[i] + (1.f - b1) * grad;
  const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = w - (lr / bc1) * (mi / denom);
}

// Device-resident step counter (CUDA-graph replay cannot change a host-side step argument):
// state = [1 - beta1^t, sqrt(1 - beta2^t), lr]; the first two are refreshed here, lr by the host.
__global__ void adam_tick_kernel(int *step, float b1, float b2, float *state) {
  const int t = *step + 1;
  *step = t;
  state[0] = (float)(1.0 - pow((double)b1, (double)t));
  state[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}

// Several strided fp32 block adds in one launch: dst[r*dst_ld + c] += src[r*src_ld_r + c*src_ld_c] per table row
// [src, dst, rows, cols, src_ld_r, src_ld_c, dst_ld, clear] (clear != 0: the source entries are zeroed after they have been
// read).  The weight-gradient kernels write lane-contiguous
// accumulators (the head's rows stacked, the first layer transposed); this moves them into the parameters' own layout
// (five torch adds per training step before).
__global__ void __launch_bounds__(256) scatter_add_table_kernel(const long long *__restrict__ table) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const long long *e = table + 8 * blockIdx.y;
  const float *src = reinterpret_cast<const float *>(e[0]);
  float *dst = reinterpret_cast<float *>(e[1]);
  const long long rows = e[2], cols = e[3], n = rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / cols, c = idx - r * cols;
    float *sp = const_cast<float *>(src) + r * e[4] + c * e[5];
    dst[r * e[6] + c] += *sp;
    if (e[7]) *sp = 0.f;                       // leave the accumulator clean for the next step (no fill launch)
  }
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_posenc_bf16(const float *x, const float *freqs, const float *extra, const float *scale_enc,
                               const float *scale_extra, int32_t scale_stride, int64_t n_points, int32_t dim,
                               int32_t n_freqs, int32_t extra_dim, int32_t k_pad, int64_t out_pitch,
                               int32_t pow2_bands, void *out_bf16, void *stream) {
  const char *fn = "nfs_posenc_bf16";
  if (n_points < 0 || dim <= 0 || n_freqs < 0 || extra_dim < 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!x || !out_bf16 || (n_freqs > 0 && !freqs) || (extra_dim > 0 && !extra))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  const int w = dim * (2 * n_freqs + 1) + extra_dim;
  if (k_pad % 8 != 0 || k_pad < w) return fail_arg(fn, NFS_E_BADARG, "k_pad must be a multiple of 8 and >= the row width");
  if (out_pitch == 0) out_pitch = k_pad;
  if (out_pitch < k_pad || (out_pitch & 7)) return fail_arg(fn, NFS_E_BADARG, "out_pitch must be >= k_pad and a multiple of 8");
  if (scale_stride <= 0) scale_stride = 1;
  if (!aligned16(out_bf16)) return fail_arg(fn, NFS_E_ALIGN, "out must be 16-byte aligned");
  const size_t smem = (size_t)kEncTile * k_pad * 2;
  if (smem > 48 * 1024) return fail_arg(fn, NFS_E_TOOLARGE, "k_pad too large");
  const long long blocks = (n_points + kEncTile - 1) / kEncTile;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many points for one launch");
  return launch_dep(fn, posenc_bf16_kernel, dim3((unsigned)blocks), dim3(kEncThreads), smem, (cudaStream_t)stream,
                    x, freqs, extra_dim > 0 ? extra : nullptr, scale_enc, scale_extra, scale_stride, n_points, dim, n_freqs,
                    extra_dim, k_pad, out_pitch, (pow2_bands != 0 && n_freqs > 1) ? 1 : 0, (__nv_bfloat16 *)out_bf16);
}

extern "C" int nfs_pack_linear_bf16(const float *w, int32_t n_dim, int32_t k_dim, int32_t n_pad, int32_t k_pad,
                                    int32_t row0, int32_t col0, void *w_bf16, void *wt_bf16, void *stream) {
  const char *fn = "nfs_pack_linear_bf16";
  if (n_dim <= 0 || k_dim <= 0 || row0 < 0 || col0 < 0 || row0 + n_dim > n_pad || col0 + k_dim > k_pad)
    return fail_arg(fn, NFS_E_BADARG, "block does not fit the padded matrix");
  if (!w || (!w_bf16 && !wt_bf16)) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  const int total = n_dim * k_dim;
  pack_linear_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w, n_dim, k_dim, n_pad, k_pad, row0, col0, (__nv_bfloat16 *)w_bf16, (__nv_bfloat16 *)wt_bf16);
  return check_launch(fn);
}

// fp32 bias -> the [n, 8] bf16 operand of the chain kernel's bias MMA: row i = [hi, mid, lo, 0, 0, 0, 0, 0] with
// hi + mid + lo == bias[i] to ~2^-24 relative (three round-to-nearest bf16 terms).
__global__ void bias_terms_kernel(const float *__restrict__ bias, int n, uint4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float b = bias[i];
  const __nv_bfloat16 hi = __float2bfloat16_rn(b);
  const float r1 = b - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
  out[i] = make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                      (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
}

// ------------------------------------------------------------------ whole-chain operand refresh
// One launch rebuilds every bf16 operand of a fused chain from the fp32 master parameters after an optimizer step:
// table row e = [src pointer, n_dim, k_dim, w_row0, w_col0, wt_row0 (-1: none), wt_col0, bias_row0] (int64 each);
//   k_dim > 0: weight W [n_dim,k_dim] -> w_stack[(w_row0+n)*256 + w_col0+k] and, transposed, wt_stack[(wt_row0+k)*256 + wt_col0+n]
//   k_dim = 0: bias [n_dim]            -> the three-term bf16 rows bias_terms[bias_row0+n] (see bias_terms_kernel)
__global__ void __launch_bounds__(256)
pack_stack_kernel(const long long *__restrict__ table, __nv_bfloat16 *__restrict__ w_stack,
                  __nv_bfloat16 *__restrict__ wt_stack, uint4 *__restrict__ terms) {
  const long long *e = table + 8 * blockIdx.y;
  const float *src = reinterpret_cast<const float *>(e[0]);
  const int N = (int)e[1], K = (int)e[2];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (K == 0) {
    if (idx >= N || terms == nullptr) return;
    const float b = __ldg(src + idx);
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    const float r1 = b - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    terms[e[7] + idx] = make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                                   (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
    return;
  }
  if (idx >= N * K) return;
  const int n = idx / K, k = idx - n * K;
  const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(src + idx));
  w_stack[(e[3] + n) * 256 + e[4] + k] = v;
  if (e[5] >= 0 && wt_stack != nullptr) wt_stack[(e[5] + k) * 256 + e[6] + n] = v;
}

extern "C" int nfs_pack_stack(const void *table, int32_t n_entries, int32_t max_elems, void *w_stack_bf16,
                              void *wt_stack_bf16, void *bias_terms_bf16, void *stream) {
  const char *fn = "nfs_pack_stack";
  if (n_entries < 0 || max_elems < 0) return fail_arg(fn, NFS_E_BADARG, "negative size");
  if (n_entries == 0 || max_elems == 0) return 0;
  if (!table || !w_stack_bf16) return fail_arg(fn, NFS_E_BADARG, "null pointer");
  if (n_entries > 65535) return fail_arg(fn, NFS_E_TOOLARGE, "too many entries");
  dim3 grid((unsigned)((max_elems + 255) / 256), (unsigned)n_entries);
  pack_stack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)table, (__nv_bfloat16 *)w_stack_bf16,
                                                            (__nv_bfloat16 *)wt_stack_bf16, (uint4 *)bias_terms_bf16);
  return check_launch(fn);
}

// General form of the above for models whose operands live in many tensors (NeRFWithDINO: per-layer W / W^T / bias
// copies for the single-layer kernels plus three stacked chain operands).  Table row (int64 x 10):
//   [src (fp32), n_dim, k_dim (0: vector), src_pitch, dst, dst_pitch, row0, col0, mode, unused]
//   mode 0: bf16 dst[(row0+n)*dst_pitch + col0+k] = src[n*src_pitch + k]
//   mode 1: bf16 dst[(row0+k)*dst_pitch + col0+n] = src[n*src_pitch + k]          (transposed)
//   mode 2: fp32 dst[row0+n] = src[n]                                            (bias copy)
//   mode 3: bias terms (uint4 rows) dst[row0+n] = split(src[n])
__global__ void __launch_bounds__(256) pack_table_kernel(const long long *__restrict__ table) {
  const long long *e = table + 10 * blockIdx.y;
  const float *src = reinterpret_cast<const float *>(e[0]);
  const int N = (int)e[1], K = (int)e[2], mode = (int)e[8];
  const int stride = gridDim.x * blockDim.x;
  if (mode >= 2) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N; idx += stride) {
      const float b = __ldg(src + idx);
      if (mode == 2) { reinterpret_cast<float *>(e[4])[e[6] + idx] = b; continue; }
      const __nv_bfloat16 hi = __float2bfloat16_rn(b);
      const float r1 = b - __bfloat162float(hi);
      const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
      const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
      reinterpret_cast<uint4 *>(e[4])[e[6] + idx] =
          make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16),
                     (uint32_t)__bfloat16_as_ushort(lo), 0u, 0u);
    }
    return;
  }
  __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(e[4]);
  if (mode == 0) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * K; idx += stride) {
      const int n = idx / K, k = idx - n * K;
      dst[(e[6] + n) * e[5] + e[7] + k] = __float2bfloat16_rn(__ldg(src + (long long)n * e[3] + k));
    }
    return;
  }
  // transposed copy through a 32 x 32 shared-memory tile: reads coalesced along k, writes coalesced along n
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8 threads
  const int tiles_k = (K + 31) >> 5, n_tiles = ((N + 31) >> 5) * tiles_k;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int n0 = (t / tiles_k) << 5, k0 = (t % tiles_k) << 5;
#pragma unroll
    for (int r = ty; r < 32; r += 8)
      tile[r][tx] = (n0 + r < N && k0 + tx < K) ? __ldg(src + (long long)(n0 + r) * e[3] + k0 + tx) : 0.f;
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
      if (k0 + r < K && n0 + tx < N) dst[(e[6] + k0 + r) * e[5] + e[7] + n0 + tx] = __float2bfloat16_rn(tile[tx][r]);
    __syncthreads();
  }
}

extern "C" int nfs_pack_table(const void *table, int32_t n_entries, int32_t max_elems, void *stream) {
  const char *fn = "nfs_pack_table";
  if (n_entries < 0 || max_elems < 0) return fail_arg(fn, NFS_E_BADARG, "negative size");
  if (n_entries == 0 || max_elems == 0) return 0;
  if (!table) return fail_arg(fn, NFS_E_BADARG, "null pointer");
  if (n_entries > 65535) return fail_arg(fn, NFS_E_TOOLARGE, "too many entries");
  // most entries are a few hundred elements (biases): a short grid per entry, threads stride over the large ones
  const int bx = (max_elems + 255) / 256;
  dim3 grid((unsigned)(bx < 16 ? bx : 16), (unsigned)n_entries);
  pack_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long *)table);
  return check_launch(fn);
}

extern "C" int nfs_scatter_add_table(const void *table, int32_t n_entries, int64_t max_elems, void *stream) {
  const char *fn = "nfs_scatter_add_table";
  if (n_entries < 0 || max_elems < 0) return fail_arg(fn, NFS_E_BADARG, "negative size");
  if (n_entries == 0 || max_elems == 0) return 0;
  if (!table) return fail_arg(fn, NFS_E_BADARG, "null table");
  const long long bx = (max_elems + 255) / 256;
  dim3 grid((unsigned)(bx < 64 ? bx : 64), (unsigned)n_entries);
  return launch_dep(fn, scatter_add_table_kernel, grid, dim3(256), 0, (cudaStream_t)stream, (const long long *)table);
}

extern "C" int nfs_bias_terms_bf16(const float *bias, int32_t n, void *terms_bf16, void *stream) {
  const char *fn = "nfs_bias_terms_bf16";
  if (n < 0) return fail_arg(fn, NFS_E_BADARG, "negative size");
  if (n == 0) return 0;
  if (!bias || !terms_bf16) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (!aligned16(terms_bf16)) return fail_arg(fn, NFS_E_ALIGN, "terms_bf16 must be 16-byte aligned");
  bias_terms_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(bias, n, (uint4 *)terms_bf16);
  return check_launch(fn);
}

extern "C" int nfs_act_grad_bf16(const float *out, const float *g_out, int64_t n_points, int32_t n_cols, int32_t act,
                                 int32_t n_pad, int64_t dy_pitch, void *dy_bf16, void *stream) {
  const char *fn = "nfs_act_grad_bf16";
  if (n_points < 0 || n_cols <= 0 || n_pad % 8 != 0 || n_pad < n_cols || act < 0 || act > 3)
    return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!out || !g_out || !dy_bf16) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (dy_pitch == 0) dy_pitch = n_pad;
  if (dy_pitch < n_pad || (dy_pitch & 7) || !aligned16(dy_bf16))
    return fail_arg(fn, NFS_E_BADARG, "dy_pitch must be >= n_pad, a multiple of 8, and dy 16-byte aligned");
  const long long threads = n_points * (n_pad / 8);
  const long long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many points for one launch");
  return launch_dep(fn, act_grad_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, out, g_out, n_points,
                    n_cols, act, n_pad, dy_pitch, (__nv_bfloat16 *)dy_bf16);
}

extern "C" int nfs_gate_bwd_bf16(const float *x, const float *freqs, const float *extra, const float *gate,
                                 const void *dc_bf16, int64_t dc_pitch, int64_t n_points, int32_t dim, int32_t n_freqs,
                                 int32_t extra_dim, int32_t n_pad, void *dlogits_bf16, void *stream) {
  const char *fn = "nfs_gate_bwd_bf16";
  if (n_points < 0 || dim <= 0 || n_freqs < 0 || extra_dim < 0 || n_pad < 8 || (n_pad & 7))
    return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!x || !gate || !dc_bf16 || !dlogits_bf16 || (n_freqs > 0 && !freqs) || (extra_dim > 0 && !extra))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (dc_pitch < dim * (2 * n_freqs + 1) + extra_dim) return fail_arg(fn, NFS_E_BADARG, "dc_pitch smaller than the row");
  if (!aligned16(dlogits_bf16)) return fail_arg(fn, NFS_E_ALIGN, "dlogits must be 16-byte aligned");
  // grid-stride over points: at most 8 resident blocks per SM worth of warps
  const long long want = (n_points + kGateBwdWarps - 1) / kGateBwdWarps;
  const long long blocks = want < 148 * 8 ? want : 148 * 8;
  return launch_dep(fn, gate_bwd_kernel, dim3((unsigned)blocks), dim3(32 * kGateBwdWarps), 0, (cudaStream_t)stream,
                    x, freqs, extra, gate, (const __nv_bfloat16 *)dc_bf16, (long long)dc_pitch, (long long)n_points, dim,
                    n_freqs, extra_dim, n_pad, (__nv_bfloat16 *)dlogits_bf16);
}

extern "C" int nfs_gate_scale_bf16(const void *c_bf16, int64_t c_pitch, const float *gate, int64_t n_points, int32_t enc_w,
                                   int32_t k_pad, void *out_bf16, int64_t out_pitch, void *stream) {
  const char *fn = "nfs_gate_scale_bf16";
  if (n_points < 0 || enc_w < 0 || k_pad <= 0 || (k_pad & 7)) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!c_bf16 || !gate || !out_bf16) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (c_pitch == 0) c_pitch = k_pad;
  if (out_pitch == 0) out_pitch = k_pad;
  if (c_pitch < k_pad || out_pitch < k_pad || ((c_pitch | out_pitch) & 7) || !aligned16(c_bf16) || !aligned16(out_bf16))
    return fail_arg(fn, NFS_E_ALIGN, "pitches must be >= k_pad and multiples of 8, tensors 16-byte aligned");
  const long long blocks = (n_points * (k_pad / 8) + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many points for one launch");
  return launch_dep(fn, gate_scale_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream,
                    (const __nv_bfloat16 *)c_bf16, (long long)c_pitch, gate, (long long)n_points, enc_w, k_pad,
                    (long long)out_pitch, (__nv_bfloat16 *)out_bf16);
}

extern "C" int nfs_gate_bwd_operand(const void *c_bf16, int64_t c_pitch, const float *gate, const void *dc_bf16,
                                    int64_t dc_pitch, int64_t n_points, int32_t enc_w, int32_t width, int32_t n_pad,
                                    void *dlogits_bf16, void *stream) {
  const char *fn = "nfs_gate_bwd_operand";
  if (n_points < 0 || enc_w < 0 || width < enc_w || n_pad < 8 || (n_pad & 7)) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!c_bf16 || !gate || !dc_bf16 || !dlogits_bf16) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  const long long w8 = (width + 7) / 8 * 8;
  if (c_pitch < w8 || dc_pitch < w8 || ((c_pitch | dc_pitch) & 7) || !aligned16(c_bf16) || !aligned16(dc_bf16) ||
      !aligned16(dlogits_bf16))
    return fail_arg(fn, NFS_E_ALIGN, "row pitches must cover the width rounded up to 8 and be multiples of 8; 16-byte aligned tensors");
  const long long want = (n_points + kGateBwdWarps - 1) / kGateBwdWarps;
  const long long blocks = want < 148 * 8 ? want : 148 * 8;
  return launch_dep(fn, gate_bwd_operand_kernel, dim3((unsigned)blocks), dim3(32 * kGateBwdWarps), 0, (cudaStream_t)stream,
                    (const __nv_bfloat16 *)c_bf16, (long long)c_pitch, gate, (const __nv_bfloat16 *)dc_bf16,
                    (long long)dc_pitch, (long long)n_points, enc_w, width, n_pad, (__nv_bfloat16 *)dlogits_bf16);
}

extern "C" int nfs_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                             int32_t decoupled, void *stream) {
  const char *fn = "nfs_adam_step";
  if (n < 0 || step <= 0) return fail_arg(fn, NFS_E_BADARG, "n < 0 or step <= 0");
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const long long blocks = (n + 255) / 256;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                 eps, weight_decay, (float)bc1, (float)sqrt(bc2),
                                                                 grad_scale, decoupled, nullptr);
  return check_launch(fn);
}

extern "C" int nfs_adam_step_dev(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t *step_counter,
                                 float *state, float grad_scale, int32_t decoupled, void *stream) {
  const char *fn = "nfs_adam_step_dev";
  if (n < 0) return fail_arg(fn, NFS_E_BADARG, "n < 0");
  if (n == 0) return 0;
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_counter || !state)
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter, beta1, beta2, state);
  int rc = check_launch(fn);
  if (rc) return rc;
  const long long blocks = (n + 255) / 256;
  adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, 0.f, beta1, beta2,
                                                                 eps, weight_decay, 1.f, 1.f, grad_scale, decoupled,
                                                                 state);
  return check_launch(fn);
}
