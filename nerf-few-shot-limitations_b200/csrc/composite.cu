// K1 — alpha compositing forward / backward for sm_100a.
//
// Replaces the ~20 (forward) + ~30 (autograd backward) full-tensor ATen passes of
//   VolumeRenderer.forward          /root/reference/src/models/nerf_mlp.py:165-215
//   volume_render_radiance          /root/reference/src/models/volume_renderer.py:4-43
// with one pass over the inputs per direction.
//
// Mapping: a sub-warp ("group") of G lanes owns one ray; each lane owns 4
// CONSECUTIVE samples per chunk, so z / density / weights move as one 128-bit
// access per lane and rgb as three (the 48 contiguous bytes of 4 samples), and a
// warp covers 32/G adjacent rays whose rows are contiguous in memory -> every
// request is a fully used run of 128 B lines.  Transmittance is a lane-local
// product followed by a log2(G)-step shuffle scan; the backward needs the
// mirror-image suffix sum and runs the same scan with shfl_down.  Rays longer
// than one chunk (S > 4G) carry T forward between chunks; the backward first
// records the chunk-entry transmittances in shared memory (z/density only) and
// then walks the chunks in reverse so the suffix sum is carried the other way.
//
// The kernel is HBM-bound: 24*S+28 B/ray forward, 36*S+28 (+4*S with g_weights)
// backward (DESIGN.md "K1").
#include "nfs_common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace nfs {
namespace {

struct CompositeArgs {
  const float *rgb, *density, *z, *rays_d, *noise;
  float noise_std;
  long long n_rays;
  int S;
  int white;
  // forward outputs
  float *out_rgb, *out_depth, *out_w;
  // backward inputs / outputs
  const float *g_rgb, *g_depth, *g_w;
  float *d_rgb, *d_density;
  // packed backward only (nfs_composite_bwd_dy): instead of d(rgb_sigma) as fp32 the kernel writes the gradient of the
  // MLP head's pre-activations as the bf16 GEMM operand of the dgrad chain / head weight gradient:
  // dy[p * dy_pitch + 0..2] = d_rgb * rgb (1 - rgb) (the head's sigmoid), dy[.. + 3] = d_sigma
  __nv_bfloat16 *dy;
  long long dy_pitch;
  // fused loss epilogue of the forward (nfs_composite_loss_fwd); target == nullptr: plain forward
  const float *target, *target_depth;
  float rgb_coef, depth_coef;          // rgb_weight * 2 / (3 N), depth_weight / N
  float *g_rgb_out, *g_depth_out;
  double *loss_sums;                   // [kLossSlots][2]: sum (rgb - target)^2, sum |depth - target_depth|
};

// The quotient R / q of the backward's closed form, q in (0, 1].  Three builds (scripts/dev/ab_k1_div.py, 2^20 rays x 64,
// d_density against the reference's fp32 autograd with the 3 % floor / time of the staged backward on one box):
//   __fdividef (2 ulp)                      1.01e-5   0.400 ms      (round 1)
//   approximate reciprocal + one Newton step  (default)             one MUFU + three FMA-pipe instructions, <= 1 ulp
//   __fdiv_rn (-DNFS_K1_EXACT_DIV)          0.71e-5   0.433 ms      (IEEE sequence, ~10 instructions + slow path)
__device__ __forceinline__ float k1_div(float r, float q) {
#if defined(NFS_K1_EXACT_DIV)
  return __fdiv_rn(r, q);
#elif defined(NFS_K1_FAST_DIV)
  return __fdividef(r, q);
#else
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(q));
  const float x0 = r * rc;
  return fmaf(fmaf(-x0, q, r), rc, x0);          // x0 + (r - x0 q) / q: the residual is exact in the FMA
#endif
}
#define NFS_K1_DIV(a, b) k1_div((a), (b))

constexpr int kLossSlots = 32;

constexpr int kBlock = 256;

// One lane's 4 consecutive samples of one ray.
struct Lane4 {
  float z[4], sg[4], col[12];
};

template <bool ALIGNED, bool PACKED, bool WITH_RGB>
__device__ __forceinline__ void load_lane(const CompositeArgs &a, long long ray, int s0, bool ray_ok,
                                          Lane4 &v) {
  const int S = a.S;
#pragma unroll
  for (int j = 0; j < 4; ++j) { v.z[j] = 0.f; v.sg[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < 12; ++j) v.col[j] = 0.f;
  if (!ray_ok || s0 >= S) return;
  const long long base = ray * (long long)S + s0;
  if (ALIGNED) {
    const float4 zz = ldg_stream4(a.z + base);
    v.z[0] = zz.x; v.z[1] = zz.y; v.z[2] = zz.z; v.z[3] = zz.w;
    if (PACKED) {
      if (WITH_RGB) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = ldg_stream4(a.rgb + (base + j) * 4);
          v.col[3 * j] = t.x; v.col[3 * j + 1] = t.y; v.col[3 * j + 2] = t.z; v.sg[j] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v.sg[j] = ldg_stream1(a.rgb + (base + j) * 4 + 3);
      }
    } else {
      const float4 dd = ldg_stream4(a.density + base);
      v.sg[0] = dd.x; v.sg[1] = dd.y; v.sg[2] = dd.z; v.sg[3] = dd.w;
      if (WITH_RGB) {
        const float *cp = a.rgb + base * 3;
        const float4 c0 = ldg_stream4(cp), c1 = ldg_stream4(cp + 4), c2 = ldg_stream4(cp + 8);
        v.col[0] = c0.x; v.col[1] = c0.y; v.col[2] = c0.z; v.col[3] = c0.w;
        v.col[4] = c1.x; v.col[5] = c1.y; v.col[6] = c1.z; v.col[7] = c1.w;
        v.col[8] = c2.x; v.col[9] = c2.y; v.col[10] = c2.z; v.col[11] = c2.w;
      }
    }
    if (a.noise != nullptr) {
      const float4 nz = ldg_stream4(a.noise + base);
      // density + randn * noise_std, two roundings (nerf_mlp.py:189-190)
      v.sg[0] = __fadd_rn(v.sg[0], __fmul_rn(nz.x, a.noise_std));
      v.sg[1] = __fadd_rn(v.sg[1], __fmul_rn(nz.y, a.noise_std));
      v.sg[2] = __fadd_rn(v.sg[2], __fmul_rn(nz.z, a.noise_std));
      v.sg[3] = __fadd_rn(v.sg[3], __fmul_rn(nz.w, a.noise_std));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (s0 + j < S) {
        v.z[j] = ldg_stream1(a.z + base + j);
        if (PACKED) {
          const float *t = a.rgb + (base + j) * 4;
          if (WITH_RGB) { v.col[3 * j] = ldg_stream1(t); v.col[3 * j + 1] = ldg_stream1(t + 1); v.col[3 * j + 2] = ldg_stream1(t + 2); }
          v.sg[j] = ldg_stream1(t + 3);
        } else {
          v.sg[j] = ldg_stream1(a.density + base + j);
          if (WITH_RGB) {
            const float *t = a.rgb + (base + j) * 3;
            v.col[3 * j] = ldg_stream1(t); v.col[3 * j + 1] = ldg_stream1(t + 1); v.col[3 * j + 2] = ldg_stream1(t + 2);
          }
        }
        if (a.noise != nullptr)
          v.sg[j] = __fadd_rn(v.sg[j], __fmul_rn(ldg_stream1(a.noise + base + j), a.noise_std));
      }
    }
  }
}

// alpha / q / exp term of the lane's 4 samples.  `zn` = z of the sample after
// the lane's last one.  Invalid samples get alpha 0, q 1 (they vanish from
// every product and sum).
struct Alpha4 {
  float alpha[4], q[4], e[4], dist[4];
};

__device__ __forceinline__ void alpha_lane(const Lane4 &v, float zn, int s0, int S, bool ray_ok,
                                           float dnorm, Alpha4 &o) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = s0 + j;
    const float znext = (j < 3) ? v.z[j + 1] : zn;
    float d = (s == S - 1) ? 1e10f : (znext - v.z[j]);   // nerf_mlp.py:181-182
    d = d * dnorm;                                        // :185
    const float x = fmaxf(v.sg[j], 0.f) * d;              // relu(density) * dists  :193
    const float e = expf(-x);
    const float al = 1.0f - e;
    const float q = (1.0f - al) + 1e-10f;                 // :197
    const bool ok = ray_ok && s < S;
    o.alpha[j] = ok ? al : 0.f;
    o.q[j] = ok ? q : 1.f;
    o.e[j] = ok ? e : 0.f;
    o.dist[j] = ok ? d : 0.f;
  }
}

// z of the sample that follows the lane's last sample.
template <int G>
__device__ __forceinline__ float next_z(const CompositeArgs &a, const Lane4 &v, long long ray, int c0,
                                        int gl, bool ray_ok) {
  float zn = __shfl_down_sync(kFullMask, v.z[0], 1, G);
  if (gl == G - 1) {
    const int sn = c0 + 4 * G;
    zn = (ray_ok && sn < a.S) ? __ldg(a.z + ray * (long long)a.S + sn) : 0.f;
  }
  return zn;
}

// exclusive prefix product of `total` over the G lanes of a group; also returns
// the group's full product in `all`.
template <int G>
__device__ __forceinline__ float group_excl_prod(float total, int gl, float &all) {
  float inc = total;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    const float t = __shfl_up_sync(kFullMask, inc, d, G);
    if (gl >= d) inc *= t;
  }
  float exc = __shfl_up_sync(kFullMask, inc, 1, G);
  if (gl == 0) exc = 1.f;
  all = __shfl_sync(kFullMask, inc, G - 1, G);
  return exc;
}

// exclusive suffix sum over the lanes of a group (lanes after me); `all` = group sum.
template <int G>
__device__ __forceinline__ float group_excl_suffix_sum(float total, int gl, float &all) {
  float inc = total;
#pragma unroll
  for (int d = 1; d < G; d <<= 1) {
    const float t = __shfl_down_sync(kFullMask, inc, d, G);
    if (gl + d < G) inc += t;
  }
  float exc = __shfl_down_sync(kFullMask, inc, 1, G);
  if (gl == G - 1) exc = 0.f;
  all = __shfl_sync(kFullMask, inc, 0, G);
  return exc;
}

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v += __shfl_xor_sync(kFullMask, v, d, G);
  return v;
}

// Loss epilogue (block-uniform call): F.mse_loss(rgb, target) (nerf_mlp.py:235, train.py:40) and
// F.l1_loss(depth, target_depth) (nerf_mlp.py:240) evaluated where the pixel is still in registers; the upstream
// gradients of the compositing backward leave this kernel instead of a chain of torch kernels.  `leader`: this
// thread holds a finished ray.  Contains a __syncthreads().
__device__ __forceinline__ void loss_epilogue(const CompositeArgs &a, long long ray, bool leader, float acc_r, float acc_g,
                                              float acc_b, float acc_d, long long slot) {
  float se = 0.f, ae = 0.f;
  if (leader) {
    const float d0 = acc_r - __ldg(a.target + ray * 3), d1 = acc_g - __ldg(a.target + ray * 3 + 1),
                d2 = acc_b - __ldg(a.target + ray * 3 + 2);
    se = d0 * d0 + d1 * d1 + d2 * d2;
    a.g_rgb_out[ray * 3] = a.rgb_coef * d0; a.g_rgb_out[ray * 3 + 1] = a.rgb_coef * d1;
    a.g_rgb_out[ray * 3 + 2] = a.rgb_coef * d2;
    if (a.target_depth != nullptr) {
      const float dd = acc_d - __ldg(a.target_depth + ray);
      ae = fabsf(dd);
      a.g_depth_out[ray] = dd > 0.f ? a.depth_coef : (dd < 0.f ? -a.depth_coef : 0.f);   // sign(dd), l1_loss backward
    }
  }
  __shared__ float red[2][kBlock / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    ae += __shfl_xor_sync(0xffffffffu, ae, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = se; red[1][threadIdx.x >> 5] = ae; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) t += (double)red[threadIdx.x][w];
    if (threadIdx.x == 0 || a.target_depth != nullptr)
      atomicAdd(a.loss_sums + 2 * (slot % kLossSlots) + threadIdx.x, t);
  }
}

// ------------------------------- forward -----------------------------------
template <int G, bool ALIGNED, bool PACKED>
__global__ void __launch_bounds__(kBlock) composite_fwd_kernel(const CompositeArgs a) {
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kRaysPerBlock = (kBlock / 32) * kGroupsPerWarp;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const long long ray = (long long)blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5) * kGroupsPerWarp + lane / G;
  const bool ray_ok = ray < a.n_rays;
  const int S = a.S;

  float dnorm = 0.f;
  if (ray_ok) {
    const float dx = __ldg(a.rays_d + ray * 3), dy = __ldg(a.rays_d + ray * 3 + 1), dz = __ldg(a.rays_d + ray * 3 + 2);
    dnorm = sqrtf(dx * dx + dy * dy + dz * dz);   // torch.norm(rays_d, dim=-1)  nerf_mlp.py:185
  }

  float t_carry = 1.f;
  float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;

  for (int c0 = 0; c0 < S; c0 += 4 * G) {
    const int s0 = c0 + 4 * gl;
    Lane4 v;
    load_lane<ALIGNED, PACKED, true>(a, ray, s0, ray_ok, v);
    const float zn = next_z<G>(a, v, ray, c0, gl, ray_ok);
    Alpha4 al;
    alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);

    // exclusive cumprod of q (nerf_mlp.py:196-199)
    const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
    float chunk_all;
    const float tb = t_carry * group_excl_prod<G>(p4, gl, chunk_all);
    const float T[4] = {tb, tb * p1, tb * p2, tb * p3};
    t_carry *= chunk_all;

    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = al.alpha[j] * T[j];                     // :202
      acc_r += w[j] * v.col[3 * j];                  // :205
      acc_g += w[j] * v.col[3 * j + 1];
      acc_b += w[j] * v.col[3 * j + 2];
      acc_d += w[j] * v.z[j];                        // :208
      acc_w += w[j];
    }
    if (a.out_w != nullptr && ray_ok && s0 < S) {
      float *wp = a.out_w + ray * (long long)S + s0;
      if (ALIGNED) {
        stg_stream4(wp, make_float4(w[0], w[1], w[2], w[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (s0 + j < S) wp[j] = w[j];
      }
    }
  }

  acc_r = group_sum<G>(acc_r);
  acc_g = group_sum<G>(acc_g);
  acc_b = group_sum<G>(acc_b);
  acc_d = group_sum<G>(acc_d);
  acc_w = group_sum<G>(acc_w);
  if (ray_ok && gl == 0) {
    if (a.white) {                                   // :211-213
      const float bg = 1.0f - acc_w;
      acc_r += bg; acc_g += bg; acc_b += bg;
    }
    a.out_rgb[ray * 3] = acc_r; a.out_rgb[ray * 3 + 1] = acc_g; a.out_rgb[ray * 3 + 2] = acc_b;
    if (a.out_depth != nullptr) a.out_depth[ray] = acc_d;
  }
  if (a.target != nullptr) loss_epilogue(a, ray, ray_ok && gl == 0, acc_r, acc_g, acc_b, acc_d, blockIdx.x);
}

// ------------------------------- backward ----------------------------------
template <int G>
__device__ __forceinline__ float group_prod(float v) {
#pragma unroll
  for (int d = G / 2; d >= 1; d >>= 1) v *= __shfl_xor_sync(kFullMask, v, d, G);
  return v;
}

template <int G, bool ALIGNED, bool PACKED>
__global__ void __launch_bounds__(kBlock, 5) composite_bwd_kernel(const CompositeArgs a, const int n_chunks) {
  extern __shared__ float s_carry[];   // [groups per block][n_chunks]; touched only when n_chunks > 1
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kRaysPerBlock = (kBlock / 32) * kGroupsPerWarp;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int group_in_block = (threadIdx.x >> 5) * kGroupsPerWarp + lane / G;
  const long long ray = (long long)blockIdx.x * kRaysPerBlock + group_in_block;
  const bool ray_ok = ray < a.n_rays;
  const int S = a.S;

  float dnorm = 0.f, gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f;
  if (ray_ok) {
    const float dx = __ldg(a.rays_d + ray * 3), dy = __ldg(a.rays_d + ray * 3 + 1), dz = __ldg(a.rays_d + ray * 3 + 2);
    dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    gr = __ldg(a.g_rgb + ray * 3); gg = __ldg(a.g_rgb + ray * 3 + 1); gb = __ldg(a.g_rgb + ray * 3 + 2);
    if (a.g_depth != nullptr) gd = __ldg(a.g_depth + ray);
  }
  // white_bkgd adds (1 - sum_i w_i) to every channel: d/dw_i = -(g_r + g_g + g_b)
  const float g_bg = a.white ? (gr + gg + gb) : 0.f;

  float *my_carry = s_carry + group_in_block * n_chunks;
  if (n_chunks > 1) {
    // phase A: transmittance at the entry of every chunk (z and density only; L1/L2 re-read below)
    float t_carry = 1.f;
    for (int c = 0; c < n_chunks; ++c) {
      const int c0 = c * 4 * G, s0 = c0 + 4 * gl;
      if (gl == 0) my_carry[c] = t_carry;
      Lane4 v;
      load_lane<ALIGNED, PACKED, false>(a, ray, s0, ray_ok, v);
      const float zn = next_z<G>(a, v, ray, c0, gl, ray_ok);
      Alpha4 al;
      alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);
      t_carry *= group_prod<G>(al.q[0] * al.q[1] * al.q[2] * al.q[3]);
    }
    __syncwarp();
  }

  // phase B: chunks in reverse; r_carry = sum of G_k w_k over all later chunks
  float r_carry = 0.f;
  for (int c = n_chunks - 1; c >= 0; --c) {
    const int c0 = c * 4 * G, s0 = c0 + 4 * gl;
    Lane4 v;
    load_lane<ALIGNED, PACKED, true>(a, ray, s0, ray_ok, v);
    float gw_in[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.g_w != nullptr && ray_ok && s0 < S) {
      const float *gp = a.g_w + ray * (long long)S + s0;
      if (ALIGNED) {
        const float4 t = ldg_stream4(gp);
        gw_in[0] = t.x; gw_in[1] = t.y; gw_in[2] = t.z; gw_in[3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (s0 + j < S) gw_in[j] = ldg_stream1(gp + j);
      }
    }
    const float zn = next_z<G>(a, v, ray, c0, gl, ray_ok);
    Alpha4 al;
    alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);

    const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
    float chunk_all;
    const float t_in = (n_chunks > 1) ? my_carry[c] : 1.f;
    const float tb = t_in * group_excl_prod<G>(p4, gl, chunk_all);
    const float T[4] = {tb, tb * p1, tb * p2, tb * p3};

    float w[4], Gi[4], gwk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = al.alpha[j] * T[j];
      Gi[j] = gr * v.col[3 * j] + gg * v.col[3 * j + 1] + gb * v.col[3 * j + 2] + gd * v.z[j] + gw_in[j] - g_bg;
      gwk[j] = Gi[j] * w[j];
    }
    // exclusive suffix sums inside the lane, then across the lanes, then across the chunks
    const float e3 = 0.f, e2 = gwk[3], e1 = e2 + gwk[2], e0 = e1 + gwk[1];
    float chunk_sum;
    const float rb = r_carry + group_excl_suffix_sum<G>(e0 + gwk[0], gl, chunk_sum);
    const float R[4] = {rb + e0, rb + e1, rb + e2, rb + e3};
    r_carry += chunk_sum;

    float ds[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dalpha = Gi[j] * T[j] - NFS_K1_DIV(R[j], al.q[j]);   // q in (0, 1]: 2-ulp quotient, 1/5 of the IEEE sequence
      // d alpha / d density = dists * exp(-relu(density) dists) * [density > 0]
      ds[j] = (v.sg[j] > 0.f) ? dalpha * al.dist[j] * al.e[j] : 0.f;
    }

    if (ray_ok && s0 < S) {
      const long long base = ray * (long long)S + s0;
      if (PACKED && a.dy != nullptr) {
        // the arithmetic of act_grad_kernel (act 2) on the values the fp32 route would have stored: bit-identical dY
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ALIGNED || s0 + j < S) {
            const float r = v.col[3 * j], g = v.col[3 * j + 1], b = v.col[3 * j + 2];
            const float dr = (w[j] * gr) * r * (1.f - r), dg = (w[j] * gg) * g * (1.f - g), db = (w[j] * gb) * b * (1.f - b);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(dr, dg), hi = __floats2bfloat162_rn(db, ds[j]);
            *reinterpret_cast<uint2 *>(a.dy + (base + j) * a.dy_pitch) =
                make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
          }
        }
      } else if (PACKED) {
        float *op = a.d_rgb + base * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ALIGNED || s0 + j < S) {
            const float4 o = make_float4(w[j] * gr, w[j] * gg, w[j] * gb, ds[j]);
            if (ALIGNED) stg_stream4(op + 4 * j, o);
            else { op[4 * j] = o.x; op[4 * j + 1] = o.y; op[4 * j + 2] = o.z; op[4 * j + 3] = o.w; }
          }
        }
      } else if (ALIGNED) {
        float *op = a.d_rgb + base * 3;
        stg_stream4(op,     make_float4(w[0] * gr, w[0] * gg, w[0] * gb, w[1] * gr));
        stg_stream4(op + 4, make_float4(w[1] * gg, w[1] * gb, w[2] * gr, w[2] * gg));
        stg_stream4(op + 8, make_float4(w[2] * gb, w[3] * gr, w[3] * gg, w[3] * gb));
        stg_stream4(a.d_density + base, make_float4(ds[0], ds[1], ds[2], ds[3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (s0 + j < S) {
            float *op = a.d_rgb + (base + j) * 3;
            op[0] = w[j] * gr; op[1] = w[j] * gg; op[2] = w[j] * gb;
            a.d_density[base + j] = ds[j];
          }
        }
      }
    }
  }
}

// ------------------------------- backward, staged ---------------------------
// Same arithmetic as composite_bwd_kernel for the common case (row-major rgb / density / z, S <= 4G, 16-byte
// aligned, no noise), with the loads decoupled from the compute: persistent blocks walk tiles of
// kStageRays rays; one thread streams each tile's three (four with g_weights) contiguous input blocks into a
// 2-stage shared-memory ring with bulk async copies (cp.async.bulk + mbarrier complete_tx), the warps read
// their samples from shared memory.  The long dependent chains of the backward (two shuffle scans, exp,
// quotient) then never wait on DRAM latency, and ~40 KB of loads per block are always in flight.
constexpr int kStageRays = 16;       // 8 warps x (32 / G) rays at G = 16; G = 8 -> 32, G = 32 -> 8 (see launch)
constexpr int kBwdStages = 2;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

template <int G>
__global__ void __launch_bounds__(kBlock, 4) composite_bwd_staged_kernel(const CompositeArgs a, const long long n_tiles) {
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kRays = (kBlock / 32) * kGroupsPerWarp;          // rays per tile
  extern __shared__ __align__(128) unsigned char s_stage[];
  const int S = a.S;
  const bool has_gw = a.g_w != nullptr;
  const int stage_floats = kRays * S * (has_gw ? 6 : 5);          // rgb 3S | density S | z S | [g_w S] per ray
  float *ring = reinterpret_cast<float *>(s_stage);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + kBwdStages * stage_floats);
  const int lane = threadIdx.x & 31, gl = lane % G;
  const int group_in_block = (threadIdx.x >> 5) * kGroupsPerWarp + lane / G;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](long long tile, int stage) {                    // thread 0 only
    const long long r0 = tile * kRays;
    const long long nr = (a.n_rays - r0) < kRays ? (a.n_rays - r0) : kRays;
    const uint32_t row = (uint32_t)(nr * S * 4);
    float *st = ring + stage * stage_floats;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(full + stage)),
                 "r"(row * (has_gw ? 6u : 5u)) : "memory");
    bulk_g2s(st, a.rgb + r0 * S * 3, row * 3, full + stage);
    bulk_g2s(st + kRays * S * 3, a.density + r0 * S, row, full + stage);
    bulk_g2s(st + kRays * S * 4, a.z + r0 * S, row, full + stage);
    if (has_gw) bulk_g2s(st + kRays * S * 5, a.g_w + r0 * S, row, full + stage);
  };

  const long long first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0)
    for (int i = 0; i < kBwdStages; ++i)
      if (first + i * step < n_tiles) issue(first + i * step, i);

  uint32_t it = 0;
  for (long long tile = first; tile < n_tiles; tile += step, ++it) {
    const int stage = it % kBwdStages;
    const uint32_t parity = (it / kBwdStages) & 1;
    const long long ray = tile * kRays + group_in_block;
    const bool ray_ok = ray < a.n_rays;
    // per-ray scalars straight from global memory (28 bytes per ray) while the tile lands
    float dnorm = 0.f, gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f;
    if (ray_ok) {
      const float dx = __ldg(a.rays_d + ray * 3), dy = __ldg(a.rays_d + ray * 3 + 1), dz = __ldg(a.rays_d + ray * 3 + 2);
      dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
      gr = __ldg(a.g_rgb + ray * 3); gg = __ldg(a.g_rgb + ray * 3 + 1); gb = __ldg(a.g_rgb + ray * 3 + 2);
      if (a.g_depth != nullptr) gd = __ldg(a.g_depth + ray);
    }
    const float g_bg = a.white ? (gr + gg + gb) : 0.f;
    {                                                               // wait for the tile
      uint32_t ok = 0, spin = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr(full + stage)), "r"(parity) : "memory");
        if (++spin > (1u << 26)) { printf("nfs_b200: composite_bwd tile wait timed out (block %d)\n", (int)blockIdx.x); __trap(); }
      }
    }
    const float *st = ring + stage * stage_floats;
    const int s0 = 4 * gl;
    Lane4 v;
    float gw_in[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) { v.z[j] = 0.f; v.sg[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < 12; ++j) v.col[j] = 0.f;
    if (ray_ok && s0 < S) {
      const int o = group_in_block * S + s0;
      const float4 c0 = *reinterpret_cast<const float4 *>(st + o * 3);
      const float4 c1 = *reinterpret_cast<const float4 *>(st + o * 3 + 4);
      const float4 c2 = *reinterpret_cast<const float4 *>(st + o * 3 + 8);
      const float4 dd = *reinterpret_cast<const float4 *>(st + kRays * S * 3 + o);
      const float4 zz = *reinterpret_cast<const float4 *>(st + kRays * S * 4 + o);
      v.col[0] = c0.x; v.col[1] = c0.y; v.col[2] = c0.z; v.col[3] = c0.w;
      v.col[4] = c1.x; v.col[5] = c1.y; v.col[6] = c1.z; v.col[7] = c1.w;
      v.col[8] = c2.x; v.col[9] = c2.y; v.col[10] = c2.z; v.col[11] = c2.w;
      v.sg[0] = dd.x; v.sg[1] = dd.y; v.sg[2] = dd.z; v.sg[3] = dd.w;
      v.z[0] = zz.x; v.z[1] = zz.y; v.z[2] = zz.z; v.z[3] = zz.w;
      if (has_gw) {
        const float4 t = *reinterpret_cast<const float4 *>(st + kRays * S * 5 + o);
        gw_in[0] = t.x; gw_in[1] = t.y; gw_in[2] = t.z; gw_in[3] = t.w;
      }
    }
    float zn = __shfl_down_sync(kFullMask, v.z[0], 1, G);          // single chunk: the last lane's successor
    if (gl == G - 1) zn = 0.f;                                      // does not exist (its distance is 1e10)
    Alpha4 al;
    alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);

    const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
    float chunk_all;
    const float tb = group_excl_prod<G>(p4, gl, chunk_all);
    const float T[4] = {tb, tb * p1, tb * p2, tb * p3};
    float w[4], Gi[4], gwk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = al.alpha[j] * T[j];
      Gi[j] = gr * v.col[3 * j] + gg * v.col[3 * j + 1] + gb * v.col[3 * j + 2] + gd * v.z[j] + gw_in[j] - g_bg;
      gwk[j] = Gi[j] * w[j];
    }
    const float e3 = 0.f, e2 = gwk[3], e1 = e2 + gwk[2], e0 = e1 + gwk[1];
    float chunk_sum;
    const float rb = group_excl_suffix_sum<G>(e0 + gwk[0], gl, chunk_sum);
    const float R[4] = {rb + e0, rb + e1, rb + e2, rb + e3};
    float ds[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dalpha = Gi[j] * T[j] - NFS_K1_DIV(R[j], al.q[j]);
      ds[j] = (v.sg[j] > 0.f) ? dalpha * al.dist[j] * al.e[j] : 0.f;
    }
    if (ray_ok && s0 < S) {
      const long long base = ray * (long long)S + s0;
      float *op = a.d_rgb + base * 3;
      stg_stream4(op,     make_float4(w[0] * gr, w[0] * gg, w[0] * gb, w[1] * gr));
      stg_stream4(op + 4, make_float4(w[1] * gg, w[1] * gb, w[2] * gr, w[2] * gg));
      stg_stream4(op + 8, make_float4(w[2] * gb, w[3] * gr, w[3] * gg, w[3] * gb));
      stg_stream4(a.d_density + base, make_float4(ds[0], ds[1], ds[2], ds[3]));
    }
    __syncthreads();                                                // everyone has read this stage
    if (threadIdx.x == 0 && tile + kBwdStages * step < n_tiles) issue(tile + kBwdStages * step, stage);
  }
}

template <int G>
int launch_bwd_staged(const CompositeArgs &a, cudaStream_t st) {
  constexpr int kRays = (kBlock / 32) * (32 / G);
  const long long n_tiles = (a.n_rays + kRays - 1) / kRays;
  const size_t smem = sizeof(float) * (size_t)kBwdStages * kRays * a.S * (a.g_w ? 6 : 5) + 64;
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(composite_bwd_staged_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return fail_cuda("nfs_composite_bwd", e);
    attr_once.mark(attr_dev);
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long max_blocks = (long long)sms * 4;
  const unsigned grid = (unsigned)(n_tiles < max_blocks ? n_tiles : max_blocks);
  composite_bwd_staged_kernel<G><<<grid, kBlock, smem, st>>>(a, n_tiles);
  return check_launch("nfs_composite_bwd");
}

// ------------------------------- forward, staged ----------------------------
// The forward with the same load pipeline as composite_bwd_staged_kernel (persistent blocks, bulk async copies of
// each tile's three contiguous input blocks into a 2-stage shared-memory ring): identical arithmetic to
// composite_fwd_kernel (bit-identical results), but ~40 KB of loads per block always in flight instead of one
// round trip per block.
template <int G>
__global__ void __launch_bounds__(kBlock, 4) composite_fwd_staged_kernel(const CompositeArgs a, const long long n_tiles) {
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kRays = (kBlock / 32) * kGroupsPerWarp;          // rays per tile
  extern __shared__ __align__(128) unsigned char s_stage[];
  const int S = a.S;
  const int stage_floats = kRays * S * 5;                         // rgb 3S | density S | z S per ray
  float *ring = reinterpret_cast<float *>(s_stage);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + kBwdStages * stage_floats);
  const int lane = threadIdx.x & 31, gl = lane % G;
  const int group_in_block = (threadIdx.x >> 5) * kGroupsPerWarp + lane / G;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](long long tile, int stage) {                    // thread 0 only
    const long long r0 = tile * kRays;
    const long long nr = (a.n_rays - r0) < kRays ? (a.n_rays - r0) : kRays;
    const uint32_t row = (uint32_t)(nr * S * 4);
    float *st = ring + stage * stage_floats;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(full + stage)), "r"(row * 5u) : "memory");
    bulk_g2s(st, a.rgb + r0 * S * 3, row * 3, full + stage);
    bulk_g2s(st + kRays * S * 3, a.density + r0 * S, row, full + stage);
    bulk_g2s(st + kRays * S * 4, a.z + r0 * S, row, full + stage);
  };

  const long long first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0)
    for (int i = 0; i < kBwdStages; ++i)
      if (first + i * step < n_tiles) issue(first + i * step, i);

  uint32_t it = 0;
  for (long long tile = first; tile < n_tiles; tile += step, ++it) {
    const int stage = it % kBwdStages;
    const uint32_t parity = (it / kBwdStages) & 1;
    const long long ray = tile * kRays + group_in_block;
    const bool ray_ok = ray < a.n_rays;
    float dnorm = 0.f;
    if (ray_ok) {
      const float dx = __ldg(a.rays_d + ray * 3), dy = __ldg(a.rays_d + ray * 3 + 1), dz = __ldg(a.rays_d + ray * 3 + 2);
      dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    }
    {                                                               // wait for the tile
      uint32_t ok = 0, spin = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr(full + stage)), "r"(parity) : "memory");
        if (++spin > (1u << 26)) { printf("nfs_b200: composite_fwd tile wait timed out (block %d)\n", (int)blockIdx.x); __trap(); }
      }
    }
    const float *st = ring + stage * stage_floats;
    const int s0 = 4 * gl;
    Lane4 v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v.z[j] = 0.f; v.sg[j] = 0.f; }
#pragma unroll
    for (int j = 0; j < 12; ++j) v.col[j] = 0.f;
    if (ray_ok && s0 < S) {
      const int o = group_in_block * S + s0;
      const float4 c0 = *reinterpret_cast<const float4 *>(st + o * 3);
      const float4 c1 = *reinterpret_cast<const float4 *>(st + o * 3 + 4);
      const float4 c2 = *reinterpret_cast<const float4 *>(st + o * 3 + 8);
      const float4 dd = *reinterpret_cast<const float4 *>(st + kRays * S * 3 + o);
      const float4 zz = *reinterpret_cast<const float4 *>(st + kRays * S * 4 + o);
      v.col[0] = c0.x; v.col[1] = c0.y; v.col[2] = c0.z; v.col[3] = c0.w;
      v.col[4] = c1.x; v.col[5] = c1.y; v.col[6] = c1.z; v.col[7] = c1.w;
      v.col[8] = c2.x; v.col[9] = c2.y; v.col[10] = c2.z; v.col[11] = c2.w;
      v.sg[0] = dd.x; v.sg[1] = dd.y; v.sg[2] = dd.z; v.sg[3] = dd.w;
      v.z[0] = zz.x; v.z[1] = zz.y; v.z[2] = zz.z; v.z[3] = zz.w;
    }
    float zn = __shfl_down_sync(kFullMask, v.z[0], 1, G);          // single chunk (S <= 4G)
    if (gl == G - 1) zn = 0.f;
    Alpha4 al;
    alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);
    const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
    float chunk_all;
    const float tb = 1.f * group_excl_prod<G>(p4, gl, chunk_all);
    const float T[4] = {tb, tb * p1, tb * p2, tb * p3};
    float w[4];
    float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w[j] = al.alpha[j] * T[j];
      acc_r += w[j] * v.col[3 * j];
      acc_g += w[j] * v.col[3 * j + 1];
      acc_b += w[j] * v.col[3 * j + 2];
      acc_d += w[j] * v.z[j];
      acc_w += w[j];
    }
    if (a.out_w != nullptr && ray_ok && s0 < S)
      stg_stream4(a.out_w + ray * (long long)S + s0, make_float4(w[0], w[1], w[2], w[3]));
    acc_r = group_sum<G>(acc_r);
    acc_g = group_sum<G>(acc_g);
    acc_b = group_sum<G>(acc_b);
    acc_d = group_sum<G>(acc_d);
    acc_w = group_sum<G>(acc_w);
    if (ray_ok && gl == 0) {
      if (a.white) {
        const float bg = 1.0f - acc_w;
        acc_r += bg; acc_g += bg; acc_b += bg;
      }
      a.out_rgb[ray * 3] = acc_r; a.out_rgb[ray * 3 + 1] = acc_g; a.out_rgb[ray * 3 + 2] = acc_b;
      if (a.out_depth != nullptr) a.out_depth[ray] = acc_d;
    }
    if (a.target != nullptr) loss_epilogue(a, ray, ray_ok && gl == 0, acc_r, acc_g, acc_b, acc_d, tile);
    __syncthreads();                                                // everyone has read this stage
    if (threadIdx.x == 0 && tile + kBwdStages * step < n_tiles) issue(tile + kBwdStages * step, stage);
  }
}

template <int G>
int launch_fwd_staged(const CompositeArgs &a, cudaStream_t st) {
  constexpr int kRays = (kBlock / 32) * (32 / G);
  const long long n_tiles = (a.n_rays + kRays - 1) / kRays;
  const size_t smem = sizeof(float) * (size_t)kBwdStages * kRays * a.S * 5 + 64;
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(composite_fwd_staged_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return fail_cuda("nfs_composite_fwd", e);
    attr_once.mark(attr_dev);
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long max_blocks = (long long)sms * 4;
  const unsigned grid = (unsigned)(n_tiles < max_blocks ? n_tiles : max_blocks);
  composite_fwd_staged_kernel<G><<<grid, kBlock, smem, st>>>(a, n_tiles);
  return check_launch("nfs_composite_fwd");
}

// ------------------------------- staged kernels for long rays (S > 128) ------
// The fine pass of BASELINE configs 3 / 5 composites 192 samples per ray: more than one 4 x 32-sample chunk, which the
// generic backward handles by reading z / density twice from global memory (phase A: chunk-entry transmittances,
// phase B: the chunks in reverse).  Here the same load pipeline as the staged kernels above brings a tile of 8 rays
// (one warp per ray) into a 2-stage shared-memory ring and BOTH phases read from there: every input byte crosses the
// memory system once.  Same arithmetic, chunk by chunk, as composite_fwd_kernel / composite_bwd_kernel<32>
// (bit-identical results).
constexpr int kLongG = 32;
constexpr int kLongRays = kBlock / 32;           // rays per tile
constexpr int kLongMaxChunks = 8;                // S <= 1024

// one lane's 4 samples of chunk c0 of the warp's ray, from the stage
__device__ __forceinline__ void load_lane_stage(const float *st, int S, int ray_in_tile, int s0, bool ok, bool with_rgb,
                                                Lane4 &v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) { v.z[j] = 0.f; v.sg[j] = 0.f; }
#pragma unroll
  for (int j = 0; j < 12; ++j) v.col[j] = 0.f;
  if (!ok || s0 >= S) return;
  const int o = ray_in_tile * S + s0;
  const float4 dd = *reinterpret_cast<const float4 *>(st + kLongRays * S * 3 + o);
  const float4 zz = *reinterpret_cast<const float4 *>(st + kLongRays * S * 4 + o);
  v.sg[0] = dd.x; v.sg[1] = dd.y; v.sg[2] = dd.z; v.sg[3] = dd.w;
  v.z[0] = zz.x; v.z[1] = zz.y; v.z[2] = zz.z; v.z[3] = zz.w;
  if (with_rgb) {
    const float4 c0 = *reinterpret_cast<const float4 *>(st + o * 3);
    const float4 c1 = *reinterpret_cast<const float4 *>(st + o * 3 + 4);
    const float4 c2 = *reinterpret_cast<const float4 *>(st + o * 3 + 8);
    v.col[0] = c0.x; v.col[1] = c0.y; v.col[2] = c0.z; v.col[3] = c0.w;
    v.col[4] = c1.x; v.col[5] = c1.y; v.col[6] = c1.z; v.col[7] = c1.w;
    v.col[8] = c2.x; v.col[9] = c2.y; v.col[10] = c2.z; v.col[11] = c2.w;
  }
}
__device__ __forceinline__ float next_z_stage(const float *st, const Lane4 &v, int S, int ray_in_tile, int c0, int lane, bool ok) {
  float zn = __shfl_down_sync(kFullMask, v.z[0], 1, kLongG);
  if (lane == kLongG - 1) {
    const int sn = c0 + 4 * kLongG;
    zn = (ok && sn < S) ? st[kLongRays * S * 4 + ray_in_tile * S + sn] : 0.f;
  }
  return zn;
}

template <bool FWD>
__global__ void __launch_bounds__(kBlock, 3) composite_long_staged_kernel(const CompositeArgs a, const long long n_tiles) {
  extern __shared__ __align__(128) unsigned char s_stage[];
  const int S = a.S;
  const bool has_gw = !FWD && a.g_w != nullptr;
  const int per_sample = FWD ? 5 : (has_gw ? 6 : 5);              // rgb 3 | density | z | [g_w]
  const int stage_floats = kLongRays * S * per_sample;
  float *ring = reinterpret_cast<float *>(s_stage);
  uint64_t *full = reinterpret_cast<uint64_t *>(ring + kBwdStages * stage_floats);
  float *t_entry = reinterpret_cast<float *>(full + kBwdStages);  // [kLongRays][kLongMaxChunks] chunk-entry transmittances
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_chunks = (S + 4 * kLongG - 1) / (4 * kLongG);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kBwdStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](long long tile, int stage) {                    // thread 0 only
    const long long r0 = tile * kLongRays;
    const long long nr = (a.n_rays - r0) < kLongRays ? (a.n_rays - r0) : kLongRays;
    const uint32_t row = (uint32_t)(nr * S * 4);
    float *st = ring + stage * stage_floats;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(full + stage)),
                 "r"(row * (uint32_t)per_sample) : "memory");
    bulk_g2s(st, a.rgb + r0 * S * 3, row * 3, full + stage);
    bulk_g2s(st + kLongRays * S * 3, a.density + r0 * S, row, full + stage);
    bulk_g2s(st + kLongRays * S * 4, a.z + r0 * S, row, full + stage);
    if (has_gw) bulk_g2s(st + kLongRays * S * 5, a.g_w + r0 * S, row, full + stage);
  };

  const long long first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0)
    for (int i = 0; i < kBwdStages; ++i)
      if (first + i * step < n_tiles) issue(first + i * step, i);

  uint32_t it = 0;
  for (long long tile = first; tile < n_tiles; tile += step, ++it) {
    const int stage = it % kBwdStages;
    const uint32_t parity = (it / kBwdStages) & 1;
    const long long ray = tile * kLongRays + warp;
    const bool ray_ok = ray < a.n_rays;
    float dnorm = 0.f, gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f;
    if (ray_ok) {
      const float dx = __ldg(a.rays_d + ray * 3), dy = __ldg(a.rays_d + ray * 3 + 1), dz = __ldg(a.rays_d + ray * 3 + 2);
      dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
      if (!FWD) {
        gr = __ldg(a.g_rgb + ray * 3); gg = __ldg(a.g_rgb + ray * 3 + 1); gb = __ldg(a.g_rgb + ray * 3 + 2);
        if (a.g_depth != nullptr) gd = __ldg(a.g_depth + ray);
      }
    }
    {                                                               // wait for the tile
      uint32_t ok = 0, spin = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_addr(full + stage)), "r"(parity) : "memory");
        if (++spin > (1u << 26)) { printf("nfs_b200: composite (long rays) tile wait timed out (block %d)\n", (int)blockIdx.x); __trap(); }
      }
    }
    const float *st = ring + stage * stage_floats;

    if (FWD) {
      float t_carry = 1.f;
      float acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_d = 0.f, acc_w = 0.f;
      for (int c0 = 0; c0 < S; c0 += 4 * kLongG) {
        const int s0 = c0 + 4 * lane;
        Lane4 v;
        load_lane_stage(st, S, warp, s0, ray_ok, true, v);
        const float zn = next_z_stage(st, v, S, warp, c0, lane, ray_ok);
        Alpha4 al;
        alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);
        const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
        float chunk_all;
        const float tb = t_carry * group_excl_prod<kLongG>(p4, lane, chunk_all);
        const float T[4] = {tb, tb * p1, tb * p2, tb * p3};
        t_carry *= chunk_all;
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[j] = al.alpha[j] * T[j];
          acc_r += w[j] * v.col[3 * j];
          acc_g += w[j] * v.col[3 * j + 1];
          acc_b += w[j] * v.col[3 * j + 2];
          acc_d += w[j] * v.z[j];
          acc_w += w[j];
        }
        if (a.out_w != nullptr && ray_ok && s0 < S)
          stg_stream4(a.out_w + ray * (long long)S + s0, make_float4(w[0], w[1], w[2], w[3]));
      }
      acc_r = group_sum<kLongG>(acc_r);
      acc_g = group_sum<kLongG>(acc_g);
      acc_b = group_sum<kLongG>(acc_b);
      acc_d = group_sum<kLongG>(acc_d);
      acc_w = group_sum<kLongG>(acc_w);
      if (ray_ok && lane == 0) {
        if (a.white) {
          const float bg = 1.0f - acc_w;
          acc_r += bg; acc_g += bg; acc_b += bg;
        }
        a.out_rgb[ray * 3] = acc_r; a.out_rgb[ray * 3 + 1] = acc_g; a.out_rgb[ray * 3 + 2] = acc_b;
        if (a.out_depth != nullptr) a.out_depth[ray] = acc_d;
      }
      if (a.target != nullptr) loss_epilogue(a, ray, ray_ok && lane == 0, acc_r, acc_g, acc_b, acc_d, tile);
    } else {
      const float g_bg = a.white ? (gr + gg + gb) : 0.f;
      float *my_t = t_entry + warp * kLongMaxChunks;
      // phase A: transmittance at the entry of every chunk (z and density, from the stage)
      float t_carry = 1.f;
      for (int c = 0; c < n_chunks; ++c) {
        const int c0 = c * 4 * kLongG, s0 = c0 + 4 * lane;
        if (lane == 0) my_t[c] = t_carry;
        Lane4 v;
        load_lane_stage(st, S, warp, s0, ray_ok, false, v);
        const float zn = next_z_stage(st, v, S, warp, c0, lane, ray_ok);
        Alpha4 al;
        alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);
        t_carry *= group_prod<kLongG>(al.q[0] * al.q[1] * al.q[2] * al.q[3]);
      }
      __syncwarp();
      // phase B: chunks in reverse; r_carry = sum of G_k w_k over all later chunks
      float r_carry = 0.f;
      for (int c = n_chunks - 1; c >= 0; --c) {
        const int c0 = c * 4 * kLongG, s0 = c0 + 4 * lane;
        Lane4 v;
        load_lane_stage(st, S, warp, s0, ray_ok, true, v);
        float gw_in[4] = {0.f, 0.f, 0.f, 0.f};
        if (has_gw && ray_ok && s0 < S) {
          const float4 t = *reinterpret_cast<const float4 *>(st + kLongRays * S * 5 + warp * S + s0);
          gw_in[0] = t.x; gw_in[1] = t.y; gw_in[2] = t.z; gw_in[3] = t.w;
        }
        const float zn = next_z_stage(st, v, S, warp, c0, lane, ray_ok);
        Alpha4 al;
        alpha_lane(v, zn, s0, S, ray_ok, dnorm, al);
        const float p1 = al.q[0], p2 = p1 * al.q[1], p3 = p2 * al.q[2], p4 = p3 * al.q[3];
        float chunk_all;
        const float tb = my_t[c] * group_excl_prod<kLongG>(p4, lane, chunk_all);
        const float T[4] = {tb, tb * p1, tb * p2, tb * p3};
        float w[4], Gi[4], gwk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[j] = al.alpha[j] * T[j];
          Gi[j] = gr * v.col[3 * j] + gg * v.col[3 * j + 1] + gb * v.col[3 * j + 2] + gd * v.z[j] + gw_in[j] - g_bg;
          gwk[j] = Gi[j] * w[j];
        }
        const float e3 = 0.f, e2 = gwk[3], e1 = e2 + gwk[2], e0 = e1 + gwk[1];
        float chunk_sum;
        const float rb = r_carry + group_excl_suffix_sum<kLongG>(e0 + gwk[0], lane, chunk_sum);
        const float R[4] = {rb + e0, rb + e1, rb + e2, rb + e3};
        r_carry += chunk_sum;
        float ds[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dalpha = Gi[j] * T[j] - NFS_K1_DIV(R[j], al.q[j]);
          ds[j] = (v.sg[j] > 0.f) ? dalpha * al.dist[j] * al.e[j] : 0.f;
        }
        if (ray_ok && s0 < S) {
          const long long base = ray * (long long)S + s0;
          float *op = a.d_rgb + base * 3;
          stg_stream4(op,     make_float4(w[0] * gr, w[0] * gg, w[0] * gb, w[1] * gr));
          stg_stream4(op + 4, make_float4(w[1] * gg, w[1] * gb, w[2] * gr, w[2] * gg));
          stg_stream4(op + 8, make_float4(w[2] * gb, w[3] * gr, w[3] * gg, w[3] * gb));
          stg_stream4(a.d_density + base, make_float4(ds[0], ds[1], ds[2], ds[3]));
        }
      }
    }
    __syncthreads();                                                // everyone has read this stage
    if (threadIdx.x == 0 && tile + kBwdStages * step < n_tiles) issue(tile + kBwdStages * step, stage);
  }
}

static size_t long_staged_smem(const CompositeArgs &a, bool fwd) {
  const int per_sample = fwd ? 5 : (a.g_w ? 6 : 5);
  return sizeof(float) * (size_t)kBwdStages * kLongRays * a.S * per_sample + 8 * kBwdStages +
         sizeof(float) * kLongRays * kLongMaxChunks + 64;
}

template <bool FWD>
int launch_long_staged(const CompositeArgs &a, cudaStream_t st) {
  const char *fn = FWD ? "nfs_composite_fwd" : "nfs_composite_bwd";
  const long long n_tiles = (a.n_rays + kLongRays - 1) / kLongRays;
  const size_t smem = long_staged_smem(a, FWD);
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(composite_long_staged_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long max_blocks = (long long)sms * 3;
  const unsigned grid = (unsigned)(n_tiles < max_blocks ? n_tiles : max_blocks);
  composite_long_staged_kernel<FWD><<<grid, kBlock, smem, st>>>(a, n_tiles);
  return check_launch(fn);
}

// ------------------------------- dispatch -----------------------------------
int pick_group(int S) { return S <= 32 ? 8 : (S <= 64 ? 16 : 32); }

template <int G, bool ALIGNED, bool PACKED>
int launch_fwd(const CompositeArgs &a, cudaStream_t st) {
  constexpr int kRaysPerBlock = (kBlock / 32) * (32 / G);
  const long long blocks = (a.n_rays + kRaysPerBlock - 1) / kRaysPerBlock;
  if (blocks > 0x7fffffffLL) return fail_arg("nfs_composite_fwd", NFS_E_TOOLARGE, "too many rays for one launch");
  composite_fwd_kernel<G, ALIGNED, PACKED><<<(unsigned)blocks, kBlock, 0, st>>>(a);
  return check_launch("nfs_composite_fwd");
}

template <int G, bool ALIGNED, bool PACKED>
int launch_bwd(const CompositeArgs &a, cudaStream_t st) {
  constexpr int kRaysPerBlock = (kBlock / 32) * (32 / G);
  const long long blocks = (a.n_rays + kRaysPerBlock - 1) / kRaysPerBlock;
  if (blocks > 0x7fffffffLL) return fail_arg("nfs_composite_bwd", NFS_E_TOOLARGE, "too many rays for one launch");
  const int n_chunks = (a.S + 4 * G - 1) / (4 * G);
  const size_t smem = n_chunks > 1 ? sizeof(float) * (size_t)n_chunks * kRaysPerBlock : 0;
  if (smem > 48 * 1024) return fail_arg("nfs_composite_bwd", NFS_E_TOOLARGE, "n_samples too large");
  composite_bwd_kernel<G, ALIGNED, PACKED><<<(unsigned)blocks, kBlock, smem, st>>>(a, n_chunks);
  return check_launch("nfs_composite_bwd");
}

template <bool FWD>
int dispatch(const CompositeArgs &a, bool aligned, bool packed, cudaStream_t st) {
  const int G = pick_group(a.S);
  // long rays (the 192-sample fine pass): both phases of the backward / the chunked forward read a staged tile
  if (aligned && !packed && a.noise == nullptr && a.S > 4 * kLongG && a.S <= 4 * kLongG * kLongMaxChunks &&
      long_staged_smem(a, FWD) <= 96 * 1024 && getenv("NFS_K1_LONG_UNSTAGED") == nullptr) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if ((a.n_rays + kLongRays - 1) / kLongRays >= 2LL * sms * 3) return launch_long_staged<FWD>(a, st);
  }
  if (!FWD && aligned && !packed && a.noise == nullptr && a.S <= 4 * G &&
      sizeof(float) * (size_t)kBwdStages * (kBlock / 32) * (32 / G) * a.S * (a.g_w ? 6 : 5) + 64 <= 96 * 1024) {
    if (G == 8) return launch_bwd_staged<8>(a, st);
    if (G == 16) return launch_bwd_staged<16>(a, st);
    return launch_bwd_staged<32>(a, st);
  }
  // forward: the staged pipeline pays once every block has at least two tiles to overlap
  if (FWD && aligned && !packed && a.noise == nullptr && a.S <= 4 * G && getenv("NFS_FWD_UNSTAGED") == nullptr &&
      sizeof(float) * (size_t)kBwdStages * (kBlock / 32) * (32 / G) * a.S * 5 + 64 <= 96 * 1024) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = (a.n_rays + (kBlock / 32) * (32 / G) - 1) / ((kBlock / 32) * (32 / G));
    if (n_tiles >= 2LL * sms * 4) {
      if (G == 8) return launch_fwd_staged<8>(a, st);
      if (G == 16) return launch_fwd_staged<16>(a, st);
      return launch_fwd_staged<32>(a, st);
    }
  }
#define NFS_CASE(GV, AL, PK)                                                         \
  if (G == GV && aligned == AL && packed == PK)                                      \
    return FWD ? launch_fwd<GV, AL, PK>(a, st) : launch_bwd<GV, AL, PK>(a, st);
  NFS_CASE(8, true, false)  NFS_CASE(8, true, true)  NFS_CASE(8, false, false)  NFS_CASE(8, false, true)
  NFS_CASE(16, true, false) NFS_CASE(16, true, true) NFS_CASE(16, false, false) NFS_CASE(16, false, true)
  NFS_CASE(32, true, false) NFS_CASE(32, true, true) NFS_CASE(32, false, false) NFS_CASE(32, false, true)
#undef NFS_CASE
  return fail_arg("nfs_composite", NFS_E_UNSUPPORTED, "no kernel variant");
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_composite_fwd(const float *rgb, const float *density, const float *z_vals,
                                 const float *rays_d, const float *noise, float noise_std,
                                 int64_t n_rays, int32_t n_samples, int32_t white_bkgd, int32_t packed,
                                 float *out_rgb, float *out_depth, float *out_weights, void *stream) {
  const char *fn = "nfs_composite_fwd";
  if (n_rays < 0 || n_samples <= 0) return fail_arg(fn, NFS_E_BADARG, "n_rays < 0 or n_samples <= 0");
  if (n_rays == 0) return 0;
  if (!rgb || !z_vals || !rays_d || !out_rgb || (!packed && !density))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  CompositeArgs a{};
  a.rgb = rgb; a.density = density; a.z = z_vals; a.rays_d = rays_d; a.noise = noise; a.noise_std = noise_std;
  a.n_rays = n_rays; a.S = n_samples; a.white = white_bkgd;
  a.out_rgb = out_rgb; a.out_depth = out_depth; a.out_w = out_weights;
  const bool aligned = (n_samples % 4 == 0) && aligned16(rgb) && aligned16(z_vals) &&
                       (packed || aligned16(density)) && (!noise || aligned16(noise)) &&
                       (!out_weights || aligned16(out_weights));
  return dispatch<true>(a, aligned, packed != 0, (cudaStream_t)stream);
}

extern "C" int nfs_composite_loss_fwd(const float *rgb, const float *density, const float *z_vals,
                                      const float *rays_d, const float *noise, float noise_std,
                                      const float *target_rgb, const float *target_depth,
                                      float rgb_weight, float depth_weight,
                                      int64_t n_rays, int32_t n_samples, int32_t white_bkgd, int32_t packed,
                                      float *out_rgb, float *out_depth, float *out_weights,
                                      float *g_rgb, float *g_depth, double *loss_sums, void *stream) {
  const char *fn = "nfs_composite_loss_fwd";
  if (n_rays < 0 || n_samples <= 0) return fail_arg(fn, NFS_E_BADARG, "n_rays < 0 or n_samples <= 0");
  if (n_rays == 0) return 0;
  if (!rgb || !z_vals || !rays_d || !out_rgb || (!packed && !density) || !target_rgb || !g_rgb || !loss_sums)
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (target_depth != nullptr && (!g_depth || !out_depth))
    return fail_arg(fn, NFS_E_BADARG, "a depth target needs the g_depth and out_depth outputs");
  CompositeArgs a{};
  a.rgb = rgb; a.density = density; a.z = z_vals; a.rays_d = rays_d; a.noise = noise; a.noise_std = noise_std;
  a.n_rays = n_rays; a.S = n_samples; a.white = white_bkgd;
  a.out_rgb = out_rgb; a.out_depth = out_depth; a.out_w = out_weights;
  a.target = target_rgb; a.target_depth = target_depth;
  a.rgb_coef = (float)(2.0 * (double)rgb_weight / (3.0 * (double)n_rays));
  a.depth_coef = (float)((double)depth_weight / (double)n_rays);
  a.g_rgb_out = g_rgb; a.g_depth_out = g_depth; a.loss_sums = loss_sums;
  const bool aligned = (n_samples % 4 == 0) && aligned16(rgb) && aligned16(z_vals) &&
                       (packed || aligned16(density)) && (!noise || aligned16(noise)) &&
                       (!out_weights || aligned16(out_weights));
  return dispatch<true>(a, aligned, packed != 0, (cudaStream_t)stream);
}

extern "C" int nfs_composite_bwd(const float *rgb, const float *density, const float *z_vals,
                                 const float *rays_d, const float *noise, float noise_std,
                                 const float *g_rgb, const float *g_depth, const float *g_weights,
                                 int64_t n_rays, int32_t n_samples, int32_t white_bkgd, int32_t packed,
                                 float *d_rgb, float *d_density, void *stream) {
  const char *fn = "nfs_composite_bwd";
  if (n_rays < 0 || n_samples <= 0) return fail_arg(fn, NFS_E_BADARG, "n_rays < 0 or n_samples <= 0");
  if (n_rays == 0) return 0;
  if (!rgb || !z_vals || !rays_d || !g_rgb || !d_rgb || (!packed && (!density || !d_density)))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  CompositeArgs a{};
  a.rgb = rgb; a.density = density; a.z = z_vals; a.rays_d = rays_d; a.noise = noise; a.noise_std = noise_std;
  a.n_rays = n_rays; a.S = n_samples; a.white = white_bkgd;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_w = g_weights; a.d_rgb = d_rgb; a.d_density = d_density;
  const bool aligned = (n_samples % 4 == 0) && aligned16(rgb) && aligned16(z_vals) &&
                       (packed || (aligned16(density) && aligned16(d_density))) &&
                       (!noise || aligned16(noise)) && (!g_weights || aligned16(g_weights)) && aligned16(d_rgb);
  return dispatch<false>(a, aligned, packed != 0, (cudaStream_t)stream);
}

extern "C" int nfs_composite_bwd_dy(const float *rgb_sigma, const float *z_vals, const float *rays_d, const float *g_rgb,
                                    const float *g_depth, const float *g_weights, int64_t n_rays, int32_t n_samples,
                                    int32_t white_bkgd, void *dy_bf16, int64_t dy_pitch, void *stream) {
  const char *fn = "nfs_composite_bwd_dy";
  if (n_rays < 0 || n_samples <= 0) return fail_arg(fn, NFS_E_BADARG, "n_rays < 0 or n_samples <= 0");
  if (n_rays == 0) return 0;
  if (!rgb_sigma || !z_vals || !rays_d || !g_rgb || !dy_bf16) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (dy_pitch < 4 || (dy_pitch & 3) || (reinterpret_cast<uintptr_t>(dy_bf16) & 7u))
    return fail_arg(fn, NFS_E_ALIGN, "dy rows must start on 8-byte boundaries (pitch % 4 == 0)");
  CompositeArgs a{};
  a.rgb = rgb_sigma; a.z = z_vals; a.rays_d = rays_d; a.n_rays = n_rays; a.S = n_samples; a.white = white_bkgd;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_w = g_weights;
  a.dy = static_cast<__nv_bfloat16 *>(dy_bf16); a.dy_pitch = dy_pitch;
  const bool aligned = (n_samples % 4 == 0) && aligned16(rgb_sigma) && aligned16(z_vals) && (!g_weights || aligned16(g_weights));
  return dispatch<false>(a, aligned, true, (cudaStream_t)stream);
}
