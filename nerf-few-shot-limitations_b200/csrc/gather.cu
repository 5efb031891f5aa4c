// K5 — feature-conditioning gather: world points -> image coordinates -> bilinear feature lookup.
//
// Replaces (the step immediately before the conditioned MLP):
//   utils.ray_utils.project_points_to_image                 /root/reference/src/utils/ray_utils.py:176-210
//   SpatialDINOFeatures.sample_features_at_points           /root/reference/src/models/dino_feature_model.py:114-148
//     (F.grid_sample, mode='bilinear', padding_mode='zeros', align_corners=False on a (1, Hp, Wp, C) map)
// One warp per point: the projection is a dozen flops (every lane computes it), the lanes then split
// the C channels of the four bilinear taps; the feature map (9 x 9 x 64 floats = 20 KB in the reference's
// configuration) stays in L1/L2.  HBM-bound: 12 B in, 4 C + 13 B out per point.
#include <cuda_bf16.h>

#include "nfs_common.cuh"

namespace nfs {
namespace {

struct GatherArgs {
  const float *points, *pose_inv, *features;
  float focal;
  int H, W, Hp, Wp, C;
  long long P;
  float *points_2d, *depths, *sampled;
  unsigned char *valid;
};

__global__ void __launch_bounds__(256) project_gather_kernel(const GatherArgs a) {
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool direct = a.pose_inv == nullptr;      // points are already normalised image coordinates (P,2)
  float m[12] = {};
  if (!direct) {
#pragma unroll
    for (int i = 0; i < 12; ++i) m[i] = __ldg(a.pose_inv + i);      // rows 0..2 of the inverse pose
  }
  for (long long p = warp; p < a.P; p += n_warps) {
    float xn, yn;
    if (direct) {
      xn = __ldg(a.points + p * 2); yn = __ldg(a.points + p * 2 + 1);
    } else {
    const float px = __ldg(a.points + p * 3), py = __ldg(a.points + p * 3 + 1), pz = __ldg(a.points + p * 3 + 2);
    // points_cam = [p, 1] @ pose_inv^T                                   ray_utils.py:193-195
    const float cx = px * m[0] + py * m[1] + pz * m[2] + m[3];
    const float cy = px * m[4] + py * m[5] + pz * m[6] + m[7];
    const float cz = px * m[8] + py * m[9] + pz * m[10] + m[11];
    const float den = cz + 1e-8f;
    const float x = __fadd_rn(__fmul_rn(__fdiv_rn(cx, den), a.focal), (float)a.W / 2);   // :201
    const float y = __fadd_rn(__fmul_rn(__fdiv_rn(cy, den), a.focal), (float)a.H / 2);   // :202
    xn = __fadd_rn(__fmul_rn(__fdiv_rn(x, (float)a.W), 2.f), -1.f);                      // :205
    yn = __fadd_rn(__fmul_rn(__fdiv_rn(y, (float)a.H), 2.f), -1.f);                      // :206
    if (lane == 0) {
      if (a.points_2d) { a.points_2d[p * 2] = xn; a.points_2d[p * 2 + 1] = yn; }
      if (a.depths) a.depths[p] = cz;
      if (a.valid) a.valid[p] = cz > 0.f ? 1 : 0;                                          // :198
    }
    }
    if (a.sampled != nullptr) {
      // grid_sample, align_corners=False: pixel = ((coord + 1) * size - 1) / 2; zeros outside
      const float ix = ((xn + 1.f) * (float)a.Wp - 1.f) * 0.5f, iy = ((yn + 1.f) * (float)a.Hp - 1.f) * 0.5f;
      const float fx = floorf(ix), fy = floorf(iy);
      const float tx = ix - fx, ty = iy - fy;
      const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
      const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
      const bool in_x0 = x0 >= 0 && x0 < a.Wp, in_x1 = x1 >= 0 && x1 < a.Wp;
      const bool in_y0 = y0 >= 0 && y0 < a.Hp, in_y1 = y1 >= 0 && y1 < a.Hp;
      const bool finite = ix == ix && iy == iy && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
      for (int c = lane; c < a.C; c += 32) {
        float v = 0.f;
        if (finite) {
          if (in_y0 && in_x0) v += w00 * __ldg(a.features + ((long long)y0 * a.Wp + x0) * a.C + c);
          if (in_y0 && in_x1) v += w01 * __ldg(a.features + ((long long)y0 * a.Wp + x1) * a.C + c);
          if (in_y1 && in_x0) v += w10 * __ldg(a.features + ((long long)y1 * a.Wp + x0) * a.C + c);
          if (in_y1 && in_x1) v += w11 * __ldg(a.features + ((long long)y1 * a.Wp + x1) * a.C + c);
        }
        a.sampled[p * a.C + c] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ K5 + K2 -> the first dense layer's operand
// The conditioned model's first layer multiplies c = [enc(x) | f(x)] (dino_feature_model.py:182): this kernel is the
// producer of that bf16 operand row by row - projection, bilinear feature lookup and positional encoding of one
// point per warp, staged in shared memory and written as whole 16-byte chunks - so neither the (P,C) fp32 features nor
// a separate encoding pass exist (K5 as the producer of G3's layer-0 operand).  Same arithmetic
// as project_gather_kernel and posenc_bf16_kernel (octave recurrence for f_k = f_0 2^k, sincosf per band otherwise).
struct OperandArgs {
  GatherArgs g;
  const float *freqs;
  int L, pow2, k_pad;
  long long out_pitch;
  __nv_bfloat16 *out;
};

constexpr int kOpWarps = 8;
constexpr int kOpMaxK = 320;

__global__ void __launch_bounds__(32 * kOpWarps) g3_operand_kernel(const OperandArgs a) {
  __shared__ __align__(16) __nv_bfloat16 s_rows[kOpWarps][kOpMaxK];
  pdl_trigger();
  pdl_wait();                                  // (nfs_common.cuh)
  const GatherArgs &g = a.g;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = (long long)blockIdx.x * kOpWarps + w, n_warps = (long long)gridDim.x * kOpWarps;
  __nv_bfloat16 *row = s_rows[w];
  const int enc_w = 3 * (2 * a.L + 1), chunks = a.k_pad >> 3;
  float m[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) m[i] = __ldg(g.pose_inv + i);
  for (long long p = warp; p < g.P; p += n_warps) {
    for (int c = lane; c < chunks; c += 32) reinterpret_cast<uint4 *>(row)[c] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const float px = __ldg(g.points + p * 3), py = __ldg(g.points + p * 3 + 1), pz = __ldg(g.points + p * 3 + 2);
    // ---- encoding: lane d < 3 owns coordinate d
    if (lane < 3) {
      const float xv = lane == 0 ? px : (lane == 1 ? py : pz);
      row[lane] = __float2bfloat16_rn(xv);
      if (a.pow2) {
        float s, c;
        sincosf(__fmul_rn(xv, __ldg(a.freqs)), &s, &c);
        for (int k = 0; k < a.L; ++k) {
          row[3 + 6 * k + lane] = __float2bfloat16_rn(s);
          row[3 + 6 * k + 3 + lane] = __float2bfloat16_rn(c);
          const float s2 = 2.f * s * c, c2 = 1.f - 2.f * s * s;
          s = s2; c = c2;
        }
      } else {
        for (int k = 0; k < a.L; ++k) {
          float s, c;
          sincosf(__fmul_rn(xv, __ldg(a.freqs + k)), &s, &c);
          row[3 + 6 * k + lane] = __float2bfloat16_rn(s);
          row[3 + 6 * k + 3 + lane] = __float2bfloat16_rn(c);
        }
      }
    }
    // ---- projection + bilinear lookup (project_gather_kernel's arithmetic)
    const float cx = px * m[0] + py * m[1] + pz * m[2] + m[3];
    const float cy = px * m[4] + py * m[5] + pz * m[6] + m[7];
    const float cz = px * m[8] + py * m[9] + pz * m[10] + m[11];
    const float den = cz + 1e-8f;
    const float x = __fadd_rn(__fmul_rn(__fdiv_rn(cx, den), g.focal), (float)g.W / 2);
    const float y = __fadd_rn(__fmul_rn(__fdiv_rn(cy, den), g.focal), (float)g.H / 2);
    const float xn = __fadd_rn(__fmul_rn(__fdiv_rn(x, (float)g.W), 2.f), -1.f);
    const float yn = __fadd_rn(__fmul_rn(__fdiv_rn(y, (float)g.H), 2.f), -1.f);
    const float ix = ((xn + 1.f) * (float)g.Wp - 1.f) * 0.5f, iy = ((yn + 1.f) * (float)g.Hp - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const float tx = ix - fx, ty = iy - fy;
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
    const bool in_x0 = x0 >= 0 && x0 < g.Wp, in_x1 = x1 >= 0 && x1 < g.Wp;
    const bool in_y0 = y0 >= 0 && y0 < g.Hp, in_y1 = y1 >= 0 && y1 < g.Hp;
    const bool finite = ix == ix && iy == iy && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
    for (int c = lane; c < g.C; c += 32) {
      float v = 0.f;
      if (finite) {
        if (in_y0 && in_x0) v += w00 * __ldg(g.features + ((long long)y0 * g.Wp + x0) * g.C + c);
        if (in_y0 && in_x1) v += w01 * __ldg(g.features + ((long long)y0 * g.Wp + x1) * g.C + c);
        if (in_y1 && in_x0) v += w10 * __ldg(g.features + ((long long)y1 * g.Wp + x0) * g.C + c);
        if (in_y1 && in_x1) v += w11 * __ldg(g.features + ((long long)y1 * g.Wp + x1) * g.C + c);
      }
      row[enc_w + c] = __float2bfloat16_rn(v);
    }
    __syncwarp();
    uint4 *dst = reinterpret_cast<uint4 *>(a.out + p * a.out_pitch);
    for (int c = lane; c < chunks; c += 32) dst[c] = reinterpret_cast<const uint4 *>(row)[c];
    __syncwarp();
  }
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_g3_operand(const float *points, const float *pose_inv, float focal, int32_t H, int32_t W,
                              const float *features, int32_t Hp, int32_t Wp, int32_t C, const float *freqs,
                              int32_t n_freqs, int32_t pow2_bands, int64_t n_points, int32_t k_pad, int64_t out_pitch,
                              void *out_bf16, void *stream) {
  const char *fn = "nfs_g3_operand";
  if (n_points < 0 || H <= 0 || W <= 0 || Hp <= 0 || Wp <= 0 || C <= 0 || n_freqs < 0 || k_pad <= 0 || (k_pad & 7) ||
      k_pad > kOpMaxK || 3 * (2 * n_freqs + 1) + C > k_pad)
    return fail_arg(fn, NFS_E_BADARG, "bad sizes (k_pad % 8 == 0, <= 320, >= encoding + feature width)");
  if (n_points == 0) return 0;
  if (!points || !pose_inv || !features || !out_bf16 || (n_freqs > 0 && !freqs))
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (out_pitch == 0) out_pitch = k_pad;
  if (out_pitch < k_pad || (out_pitch & 7) || !aligned16(out_bf16))
    return fail_arg(fn, NFS_E_ALIGN, "out_pitch must be >= k_pad, a multiple of 8, and out 16-byte aligned");
  OperandArgs a{};
  a.g.points = points; a.g.pose_inv = pose_inv; a.g.features = features; a.g.focal = focal;
  a.g.H = H; a.g.W = W; a.g.Hp = Hp; a.g.Wp = Wp; a.g.C = C; a.g.P = n_points;
  a.freqs = freqs; a.L = n_freqs; a.pow2 = (pow2_bands != 0 && n_freqs > 1) ? 1 : 0; a.k_pad = k_pad;
  a.out_pitch = out_pitch; a.out = (__nv_bfloat16 *)out_bf16;
  long long blocks = (n_points + kOpWarps - 1) / kOpWarps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  return launch_dep(fn, g3_operand_kernel, dim3((unsigned)blocks), dim3(32 * kOpWarps), 0, (cudaStream_t)stream, a);
}

extern "C" int nfs_project_gather(const float *points, const float *pose_inv, float focal, int32_t H, int32_t W,
                                  const float *features, int32_t Hp, int32_t Wp, int32_t C, int64_t n_points,
                                  float *points_2d, float *depths, unsigned char *valid, float *sampled, void *stream) {
  const char *fn = "nfs_project_gather";
  if (n_points < 0 || H <= 0 || W <= 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 0;
  if (!points) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (!pose_inv && (points_2d || depths || valid || !sampled))
    return fail_arg(fn, NFS_E_BADARG, "pose_inv == NULL (points are image coordinates) only produces `sampled`");
  if (sampled && (!features || Hp <= 0 || Wp <= 0 || C <= 0)) return fail_arg(fn, NFS_E_BADARG, "sampled needs a feature map");
  if (!points_2d && !depths && !valid && !sampled) return fail_arg(fn, NFS_E_BADARG, "no output requested");
  GatherArgs a{};
  a.points = points; a.pose_inv = pose_inv; a.features = features; a.focal = focal;
  a.H = H; a.W = W; a.Hp = Hp; a.Wp = Wp; a.C = C; a.P = n_points;
  a.points_2d = points_2d; a.depths = depths; a.valid = valid; a.sampled = sampled;
  long long blocks = (n_points + 7) / 8;                 // 8 warps per block, one point per warp per pass
  if (blocks > 148 * 16) blocks = 148 * 16;
  return launch_dep(fn, project_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, a);
}
