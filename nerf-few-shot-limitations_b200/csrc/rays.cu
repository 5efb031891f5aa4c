// K6 — ray generation + batch assembly (the step in front of the samplers).
//
// Replaces, in one launch and without the (H*W,3) round trips through HBM:
//   models.ray_sampler.get_rays / utils.ray_utils.get_rays     /root/reference/src/models/ray_sampler.py:4-30,
//                                                              /root/reference/src/utils/ray_utils.py:4-37
//   the randperm gather of a training batch                    /root/reference/src/training/train.py:272-278
//     ray_batch_o = rays_o_full.view(-1,3)[idx]; ray_batch_d = ...; target_batch = target_rgb_full.view(-1,3)[idx]
// Given pixel indices idx (row-major, p = j*W + i) the kernel evaluates the pinhole formula for exactly those
// pixels and gathers their target colours; with idx == NULL it produces every pixel in order (= get_rays).
//
// Bit-exact with the reference's ATen CPU arithmetic (pinned by tests/golden/rays.pt): the division by the focal
// length is a true division (torch's CUDA kernels multiply by 1/focal instead and differ in ~25 % of the
// elements by one ulp), the three products of `sum(dirs[..., None, :] * c2w[:3,:3], -1)` are rounded
// individually and added left to right, no FMA contraction.
// One thread per output element (ray, component): stores are fully coalesced; 12 B (+8 B index, +12 B colour)
// read and 24 B (+12 B) written per ray - HBM-bound, a few microseconds per batch.
#include "nfs_common.cuh"

namespace nfs {
namespace {

struct RayArgs {
  int H, W;
  float focal, w_half, h_half;
  const long long *idx;
  long long n;
  const float *image;
  float *rays_o, *rays_d, *target;
};

// The pose lives in device memory (a torch tensor) and the ABI does not synchronise: the kernel reads it from
// there.  Twelve scalar loads per thread, all hitting the same L1 lines.
__global__ void __launch_bounds__(256) rays_kernel(RayArgs a, const float *c2w, int row_stride) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n * 3) return;
  const long long r = e / 3;
  const int k = (int)(e - r * 3);
  long long p = r;
  if (a.idx != nullptr) p = a.idx[r];
  const long long n_pix = (long long)a.H * a.W;
  if (p < 0 || p >= n_pix) {                        // torch indexing raises here; a kernel cannot: poison the ray
    const float nan = __int_as_float(0x7fc00000);
    if (a.rays_o) a.rays_o[e] = nan;
    if (a.rays_d) a.rays_d[e] = nan;
    if (a.target) a.target[e] = nan;
    return;
  }
  if (a.rays_d != nullptr) {
    const float i = (float)(int)(p % a.W), j = (float)(int)(p / a.W);
    const float dx = __fdiv_rn(__fsub_rn(i, a.w_half), a.focal);               // (i - W*0.5) / focal   ray_sampler.py:24
    const float dy = -__fdiv_rn(__fsub_rn(j, a.h_half), a.focal);              // -(j - H*0.5) / focal
    const float dz = -1.0f;
    const float r0 = __ldg(c2w + k * row_stride), r1 = __ldg(c2w + k * row_stride + 1), r2 = __ldg(c2w + k * row_stride + 2);
    a.rays_d[e] = __fadd_rn(__fadd_rn(__fmul_rn(dx, r0), __fmul_rn(dy, r1)), __fmul_rn(dz, r2));   // :27
  }
  if (a.rays_o != nullptr) a.rays_o[e] = __ldg(c2w + k * row_stride + 3);           // :28
  if (a.target != nullptr) a.target[e] = __ldg(a.image + p * 3 + k);                // train.py:278
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_rays_generate(int32_t height, int32_t width, float focal, const float *c2w, int32_t c2w_row_stride,
                                 const int64_t *pix_idx, int64_t n_rays, const float *image,
                                 float *rays_o, float *rays_d, float *target, void *stream) {
  const char *fn = "nfs_rays_generate";
  if (height <= 0 || width <= 0 || n_rays < 0 || c2w_row_stride < 4) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (pix_idx == nullptr && n_rays != (int64_t)height * width)
    return fail_arg(fn, NFS_E_BADARG, "without pixel indices n_rays must be height * width");
  if (n_rays == 0) return 0;
  if (!c2w || (!rays_o && !rays_d && !target)) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if ((target != nullptr) != (image != nullptr)) return fail_arg(fn, NFS_E_BADARG, "target and image go together");
  RayArgs a{};
  a.H = height; a.W = width; a.focal = focal;
  a.w_half = (float)(width * 0.5); a.h_half = (float)(height * 0.5);     // python floats W*0.5, H*0.5 -> fp32 scalars
  a.idx = reinterpret_cast<const long long *>(pix_idx); a.n = n_rays; a.image = image;
  a.rays_o = rays_o; a.rays_d = rays_d; a.target = target;
  const long long blocks = (n_rays * 3 + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail_arg(fn, NFS_E_TOOLARGE, "too many rays for one launch");
  rays_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, c2w, c2w_row_stride);
  return check_launch(fn);
}
