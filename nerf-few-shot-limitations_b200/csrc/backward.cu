// K3 (merged backward) — the dgrad chain and every weight-gradient GEMM of a training step in ONE persistent launch,
// with dY handed from the chain to the weight-gradient CTAs through L2 instead of through a kernel boundary.
//
// Replaces autograd's backward of
//   nerf_model.NeRFMLP.forward    /root/reference/src/models/nerf_model.py:16-24
// (dX_l = dY_l W_l masked by ReLU', dW_l = dY_l^T X_l, db_l = sum_p dY_l) for all the points of a step.
//
// The two halves have opposite orders: the dgrad chain is tile-major (a 128-point tile runs through all layers on chip,
// fused_mlp_body.cuh), a weight gradient is layer-major (one [256 x 256] fp32 accumulator = a whole SM's tensor memory,
// summed over ALL points, wgrad_body.cuh).  As separate launches every dY element is written to HBM by the chain and read
// back by nine weight-gradient launches that each start only when the chain has finished.  Here the CTA pairs of one
// grid are split by role:
//   * producers  (the first `producer_ctas` CTAs, whole pairs): the dgrad chain over quads of four tiles, unchanged,
//     plus a release: when an epilogue warp's TMA stores of a quad have completed it bumps quad_done[quad];
//   * consumers  (the remaining CTAs, divided among the jobs by operand bytes): the weight-gradient body; the TMA warp
//     acquires quad_done[slab / 8] before it loads a slab the chain produces (jobs whose operands exist before the
//     launch - the head's - do not wait).
// Producers never wait for consumers, so the kernel cannot deadlock whatever the block scheduler does (all waits are
// bounded and trap); a consumer that falls behind simply finds its operand in HBM instead of L2.  The dgrad chain is
// bound by the SM (shared-memory / store bandwidth), the weight gradients by DRAM: run side by side they overlap, and a
// dY tile is read back while it is still resident in the 126 MB L2.
#include "fused_mlp_body.cuh"
#include "wgrad_body.cuh"

#include <cstdlib>

namespace nfs {
namespace {

constexpr int kBwMaxJobs = 12;

struct alignas(64) BackwardJobs {
  CUtensorMap tu[kBwMaxJobs], tv[kBwMaxJobs];
  WgradArgs a[kBwMaxJobs];
  unsigned cta0[kBwMaxJobs + 1];        // job j owns consumer CTAs cta0[j] .. cta0[j+1]
  int waits[kBwMaxJobs];                // 1: the job's operands are produced by the chain of this launch
  int n_jobs;
  unsigned producer_ctas;
  unsigned int *quad_done;
  unsigned quad_target;                 // arrivals per quad: 16 epilogue warps x 2 CTAs
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFmThreads, 1)
backward_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_save, const __grid_constant__ CUtensorMap tmap_b,
                      const FusedArgs a, const __grid_constant__ BackwardJobs jobs) {
  if (blockIdx.x < jobs.producer_ctas) {
    chain_body<true, false>(&tmap_x, &tmap_w, &tmap_save, &tmap_b, a, blockIdx.x >> 1, jobs.producer_ctas >> 1,
                            jobs.quad_done);
  } else {
    const unsigned c = blockIdx.x - jobs.producer_ctas;
    if (c >= jobs.cta0[jobs.n_jobs]) return;          // the CTA that rounds the grid up to whole clusters
    int j = 0;
    while (j + 1 < jobs.n_jobs && c >= jobs.cta0[j + 1]) ++j;
    wgrad_body(&jobs.tu[j], &jobs.tv[j], jobs.a[j], c - jobs.cta0[j], jobs.cta0[j + 1] - jobs.cta0[j],
               jobs.waits[j] ? jobs.quad_done : nullptr, jobs.quad_target);
  }
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" int nfs_mlp_backward_fused(const void *dy_bf16, int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                                      const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                                      const void *wt_stack_bf16, int32_t w_rows, const void *relu_bits_in,
                                      int64_t bits_rows_per_layer, const int32_t *mask_idx, void *dys_bf16,
                                      int64_t save_rows_per_layer, const nfs_wgrad_job *jobs, int32_t n_jobs,
                                      const int32_t *job_waits, uint32_t *quad_flags, int32_t producer_pairs,
                                      void *stream) {
  const char *fn = "nfs_mlp_backward_fused";
  if (n_jobs < 1 || n_jobs > kBwMaxJobs || !jobs || !job_waits || !quad_flags || !relu_bits_in || !dys_bf16)
    return fail_arg(fn, NFS_E_BADARG, "need 1..12 weight-gradient jobs, their wait list, the quad counters and a dgrad chain");
  FusedArgs a{};
  CUtensorMap tx{}, tw{}, ts{}, tb{};
  int rc = chain_prepare(fn, dy_bf16, nullptr, 0.f, 0, n_points, n_layers, k_dims, n_dims, acts, row0, wt_stack_bf16, w_rows,
                         nullptr, relu_bits_in, bits_rows_per_layer, mask_idx, dys_bf16, nullptr, save_rows_per_layer,
                         nullptr, 0, &a, &tx, &tw, &ts, &tb);
  if (rc == 1) return 0;
  if (rc) return rc;

  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = sms / 2;
  const long long n_quads = ((n_points + 127) / 128 + 3) / 4;

  BackwardJobs m{};
  double bytes[kBwMaxJobs], total = 0.0;
  long long slabs[kBwMaxJobs];
  size_t smem = kChainSmemBytes;
  int k = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const nfs_wgrad_job &j = jobs[i];
    if (j.n_points > n_points) return fail_arg(fn, NFS_E_BADARG, "a weight-gradient job covers more rows than the chain");
    size_t sm = 0;
    rc = wgrad_prepare_job(fn, j.u_bf16, j.u_pitch, j.v_bf16, j.v_pitch, j.n_points, j.m_dim, j.n_dim, j.m_valid, j.n_valid,
                           j.dw, j.ld_m, j.ld_n, j.colsum, j.colsum_of_v, &m.tu[k], &m.tv[k], &m.a[k], &sm);
    if (rc == 1) continue;
    if (rc) return rc;
    if (sm > smem) smem = sm;
    m.waits[k] = job_waits[i] != 0;
    slabs[k] = ((j.n_points + kSlabP - 1) / kSlabP + 7) / 8;           // scheduling units: quads of 8 slabs
    bytes[k] = (double)j.n_points * (j.m_dim + j.n_dim) * 2.0 + 8e5;    // + the fixed cost of a CTA, in byte-equivalents
    total += bytes[k];
    ++k;
  }
  if (k == 0) return fail_arg(fn, NFS_E_BADARG, "no non-empty weight-gradient job");
  // Split of the CTA pairs between the chain (producers) and the weight gradients (consumers).  Default from the
  // measured balance on B200 (scripts/dev/ab_backward.py); NFS_BWD_PRODUCERS overrides it for experiments.
  int prod = producer_pairs > 0 ? producer_pairs : 40;
  if (const char *e = getenv("NFS_BWD_PRODUCERS")) { if (atoi(e) > 0) prod = atoi(e); }
  if ((long long)prod > n_quads) prod = (int)n_quads;
  int min_cons_pairs = (k + 1) / 2;
  if (prod > pairs - min_cons_pairs) prod = pairs - min_cons_pairs;
  if (prod < 1) return fail_arg(fn, NFS_E_UNSUPPORTED, "too few SMs for the producer / consumer split");
  const unsigned consumers = (unsigned)(2 * (pairs - prod));
  unsigned used = 0;
  for (int i = 0; i < k; ++i) {
    long long c = (long long)(consumers * bytes[i] / total);
    if (c < 1) c = 1;
    if (c > slabs[i]) c = slabs[i];
    m.cta0[i] = used;
    used += (unsigned)c;
  }
  m.cta0[k] = used;
  if (used < consumers) {
    // distribute the remainder: rebuild the ranges with +1 for the first (consumers - used) jobs that can take it
    unsigned extra = consumers - used, counts[kBwMaxJobs];
    for (int i = 0; i < k; ++i) counts[i] = m.cta0[i + 1] - m.cta0[i];
    for (int pass = 0; pass < 8 && extra > 0; ++pass)
      for (int i = 0; i < k && extra > 0; ++i)
        if ((long long)counts[i] < slabs[i]) { ++counts[i]; --extra; }
    used = 0;
    for (int i = 0; i < k; ++i) { m.cta0[i] = used; used += counts[i]; }
    m.cta0[k] = used;
  }
  m.n_jobs = k;
  m.producer_ctas = (unsigned)(2 * prod);
  m.quad_done = quad_flags;
  m.quad_target = 32;
  unsigned grid = m.producer_ctas + used;
  grid += grid & 1u;                                  // whole clusters (the surplus CTA returns at once)

  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(backward_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  cudaError_t e = cudaMemsetAsync(quad_flags, 0, (size_t)n_quads * sizeof(uint32_t), (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(fn, e);
  backward_fused_kernel<<<grid, kFmThreads, smem, (cudaStream_t)stream>>>(tx, tw, ts, tb, a, m);
  return check_launch(fn);
}
