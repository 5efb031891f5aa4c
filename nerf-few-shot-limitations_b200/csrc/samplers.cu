// K4a / K4b — stratified and hierarchical (inverse-CDF) ray samplers for sm_100a.
//
// Replaces
//   models.ray_sampler.sample_points_along_rays   /root/reference/src/models/ray_sampler.py:47-61
//   utils.ray_utils.sample_points_along_rays      /root/reference/src/utils/ray_utils.py:55-84
//   utils.ray_utils.hierarchical_sampling         /root/reference/src/utils/ray_utils.py:101-143
//
// Bit-exactness contract: every fp32 operation the reference
// performs as a separate ATen op is a separately rounded __f*_rn here (never
// contracted into an FMA); the S-entry linspace tables come from torch on the
// host; cumsum accumulates in fp64 and rounds per entry like ATen's CPU kernel.
#include "nfs_common.cuh"

namespace nfs {
namespace {

// ---------------------------------------------------------------- stratified
// one thread = 4 consecutive samples of one ray (VEC) or one sample (scalar)
template <bool VEC>
__global__ void __launch_bounds__(256) stratified_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                         const float *__restrict__ z_base, const float *__restrict__ lower,
                                                         const float *__restrict__ upper, const float *__restrict__ t_rand,
                                                         long long n_rays, int S, float *__restrict__ z_out,
                                                         float *__restrict__ pts_out) {
  constexpr int V = VEC ? 4 : 1;
  const int per_ray = S / V;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n_rays * per_ray) return;
  const long long ray = tid / per_ray;
  const int s0 = (int)(tid - ray * per_ray) * V;
  const long long base = ray * (long long)S + s0;

  float z[V];
  if (t_rand != nullptr) {
    float t[V];
    if constexpr (VEC) { const float4 tt = ldg_stream4(t_rand + base); t[0] = tt.x; t[1] = tt.y; t[2] = tt.z; t[3] = tt.w; }
    else t[0] = ldg_stream1(t_rand + base);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const float lo = __ldg(lower + s0 + j), up = __ldg(upper + s0 + j);
      z[j] = __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), t[j]));      // ray_utils.py:79
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) z[j] = __ldg(z_base + s0 + j);       // ray_utils.py:68
  }
  if constexpr (VEC) stg_stream4(z_out + base, make_float4(z[0], z[1], z[2], z[3]));
  else z_out[base] = z[0];

  if (pts_out != nullptr) {
    const float ox = __ldg(rays_o + ray * 3), oy = __ldg(rays_o + ray * 3 + 1), oz = __ldg(rays_o + ray * 3 + 2);
    const float dx = __ldg(rays_d + ray * 3), dy = __ldg(rays_d + ray * 3 + 1), dz = __ldg(rays_d + ray * 3 + 2);
    float p[3 * V];
#pragma unroll
    for (int j = 0; j < V; ++j) {                                     // ray_utils.py:82
      p[3 * j] = __fadd_rn(ox, __fmul_rn(dx, z[j]));
      p[3 * j + 1] = __fadd_rn(oy, __fmul_rn(dy, z[j]));
      p[3 * j + 2] = __fadd_rn(oz, __fmul_rn(dz, z[j]));
    }
    float *pp = pts_out + base * 3;
    if constexpr (VEC) {
      stg_stream4(pp, make_float4(p[0], p[1], p[2], p[3]));
      stg_stream4(pp + 4, make_float4(p[4], p[5], p[6], p[7]));
      stg_stream4(pp + 8, make_float4(p[8], p[9], p[10], p[11]));
    } else {
      pp[0] = p[0]; pp[1] = p[1]; pp[2] = p[2];
    }
  }
}

// -------------------------------------------------------------- hierarchical
// one warp = one ray.  Shared memory per warp: cdf[M+1] | zc[M+1] | smp[P2] | buf[M+1+Ni]  (P2 = pow2 >= Ni)
// The reference sorts cat([z_vals, samples]) (ray_utils.py:138).  z_vals is already sorted and the inverse-CDF
// samples are non-decreasing in u, so with sorted draws (perturb=False: u = linspace) both lists are sorted and the
// output is a MERGE: every element's position is its own index plus its rank in the other list (one binary search,
// ~8 steps), instead of a 256-key bitonic network (288 compare-exchange sweeps - the kernel was instruction-bound
// on it: 462 us per 65 536 rays, 4.2 % of a full-frame render).  Random draws (training) first sort the Ni samples
// (bitonic over P2 keys).  Either way the result is the sorted multiset, bit-identical to torch.sort's values.
__global__ void __launch_bounds__(128) hierarchical_kernel(
    const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ z_vals,
    const float *__restrict__ weights, const float *__restrict__ u, long long u_stride,
    const float *__restrict__ cdf_in, long long n_rays, int M, int Ni, int P2, float *__restrict__ z_out,
    float *__restrict__ pts_out, float *__restrict__ cdf_out, long long *__restrict__ idx_out,
    float *__restrict__ samples_out) {
  extern __shared__ float smem[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long ray = (long long)blockIdx.x * warps + warp;
  if (ray >= n_rays) return;   // whole warp leaves together
  const int total = M + 1 + Ni;
  const int per_warp = 2 * (M + 1) + P2 + total;
  float *cdf = smem + (size_t)warp * per_warp;
  float *zc = cdf + (M + 1);
  float *smp = zc + (M + 1);
  float *buf = smp + P2;

  for (int i = lane; i <= M; i += 32) zc[i] = __ldg(z_vals + ray * (long long)(M + 1) + i);

  if (cdf_in != nullptr) {
    for (int i = lane; i <= M; i += 32) cdf[i] = __ldg(cdf_in + ray * (long long)(M + 1) + i);
  } else {
    // weights + 1e-5 ; pdf = w / sum(w) ; cdf = [0, cumsum(pdf)]     ray_utils.py:105-110
    double part = 0.0;
    for (int i = lane; i < M; i += 32) {
      const float w = __fadd_rn(__ldg(weights + ray * (long long)M + i), 1e-5f);
      buf[i] = w;
      part += (double)w;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(kFullMask, part, d);
    const float wsum = (float)part;
    __syncwarp();
    // blocked layout for the scan: lane owns [lane*per, lane*per+per)
    const int per = (M + 31) / 32;
    const int b = lane * per, e = min(M, b + per);
    double run = 0.0;
    for (int i = b; i < e; ++i) run += (double)__fdiv_rn(buf[i], wsum);
    double inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double t = __shfl_up_sync(kFullMask, inc, d);
      if (lane >= d) inc += t;
    }
    double acc = inc - run;   // exclusive prefix (exact: fp32 addends)
    for (int i = b; i < e; ++i) {
      acc += (double)__fdiv_rn(buf[i], wsum);
      cdf[i + 1] = (float)acc;   // fp64 accumulate, round per entry = ATen CPU cumsum
    }
    if (lane == 0) cdf[0] = 0.f;
  }
  __syncwarp();
  if (cdf_out != nullptr)
    for (int i = lane; i <= M; i += 32) cdf_out[ray * (long long)(M + 1) + i] = cdf[i];

  const float *urow = u + ray * u_stride;
  bool sorted = true;
  float prev_last = -__int_as_float(0x7f800000);
  for (int k0 = 0; k0 < Ni; k0 += 32) {
    const int k = k0 + lane;
    float smp_k = __int_as_float(0x7f800000);
    if (k < Ni) {
      const float uu = __ldg(urow + k);
      // searchsorted(cdf, u, right=True): number of entries <= u     ray_utils.py:121
      int lo = 0, hi = M + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      const int below = max(0, lo - 1), above = min(M, lo);             // :122-123
      const float c0 = cdf[below], c1 = cdf[above];
      const float b0 = zc[below], b1 = zc[above];
      float den = __fsub_rn(c1, c0);                                    // :132
      if (den < 1e-5f) den = 1.0f;                                      // :133
      const float t = __fdiv_rn(__fsub_rn(uu, c0), den);                // :134
      smp_k = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));           // :135
      if (idx_out != nullptr) idx_out[ray * (long long)Ni + k] = lo;
      if (samples_out != nullptr) samples_out[ray * (long long)Ni + k] = smp_k;
    }
    smp[k] = smp_k;                                                     // k < P2: rounds of 32 never pass a power of two >= 32
    // non-decreasing so far?  (compare with the previous element: the lane below, or the last lane of the previous round)
    float before = __shfl_up_sync(kFullMask, smp_k, 1);
    if (lane == 0) before = prev_last;
    if (k < Ni && !(before <= smp_k)) sorted = false;
    prev_last = __shfl_sync(kFullMask, smp_k, 31);
  }
  for (int i = ((Ni + 31) & ~31) + lane; i < P2; i += 32) smp[i] = __int_as_float(0x7f800000);
  sorted = __all_sync(kFullMask, sorted);
  __syncwarp();

  if (!sorted) {
    // bitonic sort of the P2 sample keys, ascending
    for (int k = 2; k <= P2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P2; i += 32) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float x = smp[i], y = smp[ixj];
            const bool asc = (i & k) == 0;
            if ((x > y) == asc) { smp[i] = y; smp[ixj] = x; }
          }
        }
        __syncwarp();
      }
    }
  }

  // The merge below needs the coarse depths ascending, as every caller in the reference provides them.  The reference
  // itself sorts the concatenation (ray_utils.py:139), so an unsorted row must still give the sorted multiset: check,
  // and sort the row in place when needed (odd-even transposition; the sampling above used the original order, as the
  // reference's gather does).  NaN depths are outside the contract (torch.sort puts them last; NFS_DEBUG_CHECKS=1
  // validates the input in ops.sample_hierarchical).
  {
    bool z_sorted = true;
    for (int i = lane; i < M; i += 32)
      if (!(zc[i] <= zc[i + 1])) z_sorted = false;
    if (!__all_sync(kFullMask, z_sorted)) {
      for (int r = 0; r <= M; ++r) {
        for (int i = (r & 1) + 2 * lane; i < M; i += 64) {
          const float x = zc[i], y = zc[i + 1];
          if (x > y) { zc[i] = y; zc[i + 1] = x; }
        }
        __syncwarp();
      }
    }
  }

  // merge: position = own index + rank in the other list (coarse depths first among equals)
  for (int i = lane; i <= M; i += 32) {
    const float v = zc[i];
    int lo = 0, hi = Ni;                                 // number of samples < v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (smp[mid] < v) lo = mid + 1; else hi = mid;
    }
    buf[i + lo] = v;
  }
  for (int k = lane; k < Ni; k += 32) {
    const float v = smp[k];
    int lo = 0, hi = M + 1;                              // number of coarse depths <= v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (zc[mid] <= v) lo = mid + 1; else hi = mid;
    }
    buf[k + lo] = v;
  }
  __syncwarp();

  const float ox = __ldg(rays_o + ray * 3), oy = __ldg(rays_o + ray * 3 + 1), oz = __ldg(rays_o + ray * 3 + 2);
  const float dx = __ldg(rays_d + ray * 3), dy = __ldg(rays_d + ray * 3 + 1), dz = __ldg(rays_d + ray * 3 + 2);
  for (int i = lane; i < total; i += 32) z_out[ray * (long long)total + i] = buf[i];
  if (pts_out != nullptr) {
    float *pp = pts_out + ray * (long long)total * 3;
    for (int i = lane; i < 3 * total; i += 32) {                      // :141
      const int s = i / 3, c = i - 3 * s;
      const float o = c == 0 ? ox : (c == 1 ? oy : oz);
      const float d = c == 0 ? dx : (c == 1 ? dy : dz);
      pp[i] = __fadd_rn(o, __fmul_rn(d, buf[s]));
    }
  }
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_sample_stratified(const float *rays_o, const float *rays_d, const float *z_base,
                                     const float *lower, const float *upper, const float *t_rand,
                                     int64_t n_rays, int32_t n_samples, float *z_out, float *pts_out,
                                     void *stream) {
  const char *fn = "nfs_sample_stratified";
  if (n_rays < 0 || n_samples <= 0) return fail_arg(fn, NFS_E_BADARG, "n_rays < 0 or n_samples <= 0");
  if (n_rays == 0) return 0;
  if (!z_out || !z_base || (t_rand && (!lower || !upper)) || (pts_out && (!rays_o || !rays_d)))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  const bool vec = (n_samples % 4 == 0) && aligned16(z_out) && (!t_rand || aligned16(t_rand)) &&
                   (!pts_out || aligned16(pts_out));
  const long long threads = n_rays * (long long)(vec ? n_samples / 4 : n_samples);
  const long long blocks = (threads + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many samples for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  if (vec)
    stratified_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(rays_o, rays_d, z_base, lower, upper, t_rand, n_rays, n_samples, z_out, pts_out);
  else
    stratified_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(rays_o, rays_d, z_base, lower, upper, t_rand, n_rays, n_samples, z_out, pts_out);
  return check_launch(fn);
}

extern "C" int nfs_sample_hierarchical(const float *rays_o, const float *rays_d, const float *z_vals,
                                       const float *weights, const float *u, int64_t u_stride,
                                       const float *cdf_in, int64_t n_rays, int32_t n_bins,
                                       int32_t n_importance, float *z_out, float *pts_out, float *cdf_out,
                                       int64_t *idx_out, float *samples_out, void *stream) {
  const char *fn = "nfs_sample_hierarchical";
  if (n_rays < 0 || n_bins <= 0 || n_importance < 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_rays == 0) return 0;
  if (!z_vals || (!weights && !cdf_in) || (n_importance > 0 && !u) || !z_out || (pts_out && (!rays_o || !rays_d)))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (!rays_o || !rays_d) return fail_arg(fn, NFS_E_BADARG, "rays_o / rays_d are required");
  const int total = n_bins + 1 + n_importance;
  if (total > 4096) return fail_arg(fn, NFS_E_TOOLARGE, "n_bins + 1 + n_importance > 4096");
  int p2 = 32;                                  // samples are processed in rounds of 32; bitonic needs a power of two
  while (p2 < n_importance) p2 <<= 1;
  // scratch for w+1e-5 lives in buf (M <= total entries)
  const size_t per_warp = sizeof(float) * (size_t)(2 * (n_bins + 1) + p2 + total);
  int warps = 4;
  while (warps > 1 && per_warp * warps > 48 * 1024) warps >>= 1;
  if (per_warp * warps > 48 * 1024) return fail_arg(fn, NFS_E_TOOLARGE, "shared memory");
  const long long blocks = (n_rays + warps - 1) / warps;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many rays for one launch");
  hierarchical_kernel<<<(unsigned)blocks, warps * 32, per_warp * warps, (cudaStream_t)stream>>>(
      rays_o, rays_d, z_vals, weights, u, (long long)u_stride, cdf_in, n_rays, n_bins, n_importance, p2, z_out,
      pts_out, cdf_out, (long long *)idx_out, samples_out);
  return check_launch(fn);
}
