// K2 — sin/cos positional encoding (standalone, fp32) for sm_100a.
//
// Replaces the 2L x (mul, sin|cos) + cat chain of
//   PositionalEncoding.forward            /root/reference/src/models/positional_encoding.py:27-33
//   nerf_mlp.PositionalEncoding.forward   /root/reference/src/models/nerf_mlp.py:24-33
// Output row: [x (D, optional) | sin(x f_0) (D) | cos(x f_0) (D) | ... | cos(x f_{L-1}) (D)].
//
// One CTA encodes a tile of points.  Each thread evaluates sincosf (accurate
// range reduction, no fast-math: arguments reach 2^(L-1)*|x| ~ 3000) for one
// (point, frequency, coordinate) triple and drops both results into a shared
// tile laid out exactly like the output rows; the tile then leaves as one
// contiguous, fully coalesced 128-bit stream (the rows of a tile are adjacent
// in HBM).  Algorithmic traffic: 4*D in + 4*D*(2L+1) out per point.
#include "nfs_common.cuh"

namespace nfs {
namespace {

constexpr int kPosencThreads = 256;

__global__ void __launch_bounds__(kPosencThreads) posenc_kernel(const float *__restrict__ x,
                                                                const float *__restrict__ freqs,
                                                                long long n_points, int D, int L, int include_input,
                                                                int tile_pts, float *__restrict__ out) {
  extern __shared__ __align__(16) float s_tile[];
  const int W = D * (2 * L + (include_input ? 1 : 0));
  const long long p0 = (long long)blockIdx.x * tile_pts;
  const int np = (int)min((long long)tile_pts, n_points - p0);
  const int off = include_input ? D : 0;

  // (point, freq, coord) triples; coord fastest so x loads coalesce
  const int per_pt = L * D;
  for (int t = threadIdx.x; t < np * per_pt; t += kPosencThreads) {
    const int p = t / per_pt, r = t - p * per_pt;
    const int k = r / D, d = r - k * D;
    const float xv = __ldg(x + (p0 + p) * D + d);
    const float arg = __fmul_rn(xv, __ldg(freqs + k));     // x * freq  (positional_encoding.py:31)
    float s, c;
    sincosf(arg, &s, &c);
    float *row = s_tile + p * W + off + 2 * k * D;
    row[d] = s;
    row[D + d] = c;
  }
  if (include_input)
    for (int t = threadIdx.x; t < np * D; t += kPosencThreads) {
      const int p = t / D, d = t - p * D;
      s_tile[p * W + d] = __ldg(x + (p0 + p) * D + d);
    }
  __syncthreads();

  float *dst = out + p0 * W;
  const int total = np * W;
  if (aligned16_dev(dst) && (total & 3) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(s_tile);
    for (int t = threadIdx.x; t < total / 4; t += kPosencThreads) stg_stream4(dst + 4 * t, s4[t]);
  } else {
    for (int t = threadIdx.x; t < total; t += kPosencThreads) dst[t] = s_tile[t];
  }
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_posenc_fwd(const float *x, const float *freqs, int64_t n_points, int32_t dim,
                              int32_t n_freqs, int32_t include_input, float *out, void *stream) {
  const char *fn = "nfs_posenc_fwd";
  if (n_points < 0 || dim <= 0 || n_freqs < 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  const int W = dim * (2 * n_freqs + (include_input ? 1 : 0));
  if (n_points == 0 || W == 0) return 0;
  if (!x || !out || (n_freqs > 0 && !freqs)) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  // tile: as many points as fit 32 KB of shared memory, at most 64, multiple of 4
  // (so every tile but the last starts 16-byte aligned when `out` is).
  int tile = (32 * 1024) / (4 * W);
  if (tile > 64) tile = 64;
  if (tile >= 4) tile &= ~3;
  if (tile < 1) return fail_arg(fn, NFS_E_TOOLARGE, "encoded row wider than 8192 floats");
  const long long blocks = (n_points + tile - 1) / tile;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many points for one launch");
  posenc_kernel<<<(unsigned)blocks, kPosencThreads, sizeof(float) * (size_t)tile * W, (cudaStream_t)stream>>>(
      x, freqs, n_points, dim, n_freqs, include_input, tile, out);
  return check_launch(fn);
}
