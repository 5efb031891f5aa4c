// K5 — feature-conditioning gather: world points -> image coordinates -> bilinear feature lookup.
//
// Replaces (the step immediately before the conditioned MLP):
//   utils.ray_utils.project_points_to_image                 /root/reference/src/utils/ray_utils.py:176-210
//   SpatialDINOFeatures.sample_features_at_points           /root/reference/src/models/dino_feature_model.py:114-148
//     (F.grid_sample, mode='bilinear', padding_mode='zeros', align_corners=False on a (1, Hp, Wp, C) map)
// One warp per point: the projection is a dozen flops (every lane computes it), the lanes then split
// the C channels of the four bilinear taps; the feature map (9 x 9 x 64 floats = 20 KB in the reference's
// configuration) stays in L1/L2.  HBM-bound: 12 B in, 4 C + 13 B out per point.
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

}  // namespace
}  // namespace nfs

using namespace nfs;

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
