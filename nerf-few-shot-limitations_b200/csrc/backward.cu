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
//   * producers  (the first `producer_ctas` CTAs, whole pairs; 38 of 74 pairs on B200): the dgrad chain over quads of
//     four tiles, unchanged, plus a release: when an epilogue warp's TMA stores of a quad have COMPLETED it bumps
//     quad_done[quad].  The stores carry an L2 evict_last hint;
//   * consumers  (the remaining pairs, divided among the jobs by measured cost): the weight-gradient body on CTA pairs
//     (wgrad_pair_body.cuh; the single-CTA body for shapes it does not cover); the TMA thread acquires
//     quad_done[slab / 8] before it loads a slab the chain produces (jobs whose operands exist before the launch - the
//     head's - do not wait), loads everything evict_first, and once the slab's MMAs have completed the idle drain warps
//     discard the dY lines the CTA loaded (discard.global.L2: each plane of dY has exactly one reader), so that dY is
//     neither read from nor written back to DRAM.
// Producers never depend on consumers for progress - the back-pressure that keeps them at most 5/4 of a round ahead (so
// that a quad's dY is still in L2 when it is read) gives up after a bounded spin - so the kernel cannot deadlock whatever
// the block scheduler does; every other wait is bounded and traps.  Measured on the cfg 3 step (profiles/r02f_*): 2.0 ms
// against 2.37 ms for the dgrad chains + nine weight-gradient launches, 6.3 GB of DRAM traffic against 13.1 GB; what
// bounds it is SM time on both sides, not DRAM.
#include "fused_mlp_body.cuh"
#include "wgrad_pair_body.cuh"

#include <cstdlib>

namespace nfs {
namespace {

constexpr int kBwMaxJobs = 12;

struct alignas(64) BackwardJobs {
  CUtensorMap tu[kBwMaxJobs], tv[kBwMaxJobs];
  WgradArgs a[kBwMaxJobs];
  unsigned cta0[kBwMaxJobs + 1];        // job j owns consumer CTAs cta0[j] .. cta0[j+1]
  int waits[kBwMaxJobs];                // 1: the job's operands are produced by the chain of this launch
  int paired[kBwMaxJobs];               // 1: the job's CTAs work as pairs (wgrad_pair_body), 0: one by one (wgrad_body)
  int n_jobs;
  unsigned producer_ctas;
  unsigned int *quad_done;
  unsigned quad_target;                 // arrivals per quad: 16 epilogue warps x 2 CTAs
  unsigned int *quad_consumed;          // back-pressure: jobs that have loaded the quad (NULL = off)
  unsigned consumed_target;             // number of jobs that wait for the chain
  long long window;                     // producers stay at most this many quads ahead of the consumers
  int print_times;                      // developer switch NFS_BWD_TIMES: every CTA prints its role and start / end time
};

__device__ __forceinline__ unsigned long long bw_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFmThreads, 1)
backward_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_save, const __grid_constant__ CUtensorMap tmap_b,
                      const FusedArgs a, const __grid_constant__ BackwardJobs jobs) {
  const unsigned long long t_start = jobs.print_times ? bw_now() : 0ull;
  if (blockIdx.x < jobs.producer_ctas) {
    chain_body<true, false>(&tmap_x, &tmap_w, &tmap_save, &tmap_b, a, blockIdx.x >> 1, jobs.producer_ctas >> 1,
                            jobs.quad_done, jobs.quad_consumed, jobs.consumed_target, jobs.window);
  } else {
    const unsigned c = blockIdx.x - jobs.producer_ctas;
    if (c >= jobs.cta0[jobs.n_jobs]) return;          // the CTA that rounds the grid up to whole clusters
    int j = 0;
    while (j + 1 < jobs.n_jobs && c >= jobs.cta0[j + 1]) ++j;
    if (jobs.paired[j])
      wgrad_pair_body(&jobs.tu[j], &jobs.tv[j], jobs.a[j], (c - jobs.cta0[j]) >> 1, (jobs.cta0[j + 1] - jobs.cta0[j]) >> 1,
                      jobs.waits[j] ? jobs.quad_done : nullptr, jobs.quad_target,
                      jobs.waits[j] ? jobs.quad_consumed : nullptr);
    else
      wgrad_body(&jobs.tu[j], &jobs.tv[j], jobs.a[j], c - jobs.cta0[j], jobs.cta0[j + 1] - jobs.cta0[j],
                 jobs.waits[j] ? jobs.quad_done : nullptr, jobs.quad_target, jobs.waits[j] ? jobs.quad_consumed : nullptr);
    if (jobs.print_times && threadIdx.x == 0)
      printf("bwtimes consumer cta %u job %d of %u start %llu end %llu\n", c, j, jobs.cta0[j + 1] - jobs.cta0[j], t_start, bw_now());
    return;
  }
  if (jobs.print_times && threadIdx.x == 0) printf("bwtimes producer cta %u start %llu end %llu\n", blockIdx.x, t_start, bw_now());
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
    m.waits[k] = job_waits[i] != 0 && getenv("NFS_BWD_NOWAIT") == nullptr;   // (developer switch: timing only, wrong results)
    if (m.waits[k] && getenv("NFS_BWD_NODISCARD") == nullptr) {
      // which operand lies in the chain's output tensor (256-wide rows: whole 128-byte lines per CTA of a pair)?
      const uint8_t *lo = reinterpret_cast<const uint8_t *>(dys_bf16);
      const uint8_t *hi = lo + (size_t)n_layers * (size_t)save_rows_per_layer * (size_t)n_dims[0] * 2;
      if (m.a[k].u_ptr >= lo && m.a[k].u_ptr < hi && j.m_dim == 256 && j.u_pitch == 256) m.a[k].discard = 1;
      else if (m.a[k].v_ptr >= lo && m.a[k].v_ptr < hi && j.n_dim == 256 && j.v_pitch == 256) m.a[k].discard = 2;
    }
    slabs[k] = ((j.n_points + kSlabP - 1) / kSlabP + 7) / 8;           // scheduling units: quads of 8 slabs
    ++k;
  }
  if (k == 0) return fail_arg(fn, NFS_E_BADARG, "no non-empty weight-gradient job");
  // Split of the CTA pairs between the chain (producers) and the weight gradients (consumers).  Default from the
  // measured balance on B200 (scripts/dev/ab_backward.py); NFS_BWD_PRODUCERS overrides it for experiments.
  int prod = producer_pairs > 0 ? producer_pairs : (pairs + 2) / 2;
  if (const char *e = getenv("NFS_BWD_PRODUCERS")) { if (atoi(e) > 0) prod = atoi(e); }
  if ((long long)prod > n_quads) prod = (int)n_quads;
  int min_cons_pairs = k;
  if (prod > pairs - min_cons_pairs) prod = pairs - min_cons_pairs;
  if (prod < 1) return fail_arg(fn, NFS_E_UNSUPPORTED, "too few SMs for the producer / consumer split");
  // The consumer PAIRS are divided among the jobs in proportion to their cost (largest remainder, at least one pair
  // each).  Cost per point, measured on B200 with the jobs confined to 16-36 SMs (scripts/dev/wgrad_pair.py): a pair
  // takes a 64-point slab of a 256 x 256 job in 390 ns, of the first layer's job (N = 64, column sums on the tensor
  // core) in 260 ns, of the head's (N = 64, column sums from shared memory) in 380 ns; jobs the pair body does not
  // cover run CTA by CTA (wgrad_body) at about twice that.  Inside the merged kernel, with the chain running beside them,
  // the three take ~470 / 370 / 515 ns (scripts/dev/bwd_times.py) - the weights below.
  const int cons_pairs = pairs - prod;
  double small_w = 0.8, head_w = 0.97;
  if (const char *e = getenv("NFS_BWD_SMALL_W")) small_w = atof(e);
  if (const char *e = getenv("NFS_BWD_HEAD_W")) head_w = atof(e);
  double cost[kBwMaxJobs], cost_total = 0.0;
  for (int i = 0; i < k; ++i) {
    m.paired[i] = wgrad_pair_ok(m.a[i]) && getenv("NFS_BWD_NOPAIR") == nullptr;
    if (m.paired[i] && kWpSmemBytes > smem) smem = kWpSmemBytes;
    double w = 2.0 * (m.a[i].M + m.a[i].N) / 512.0;
    if (m.paired[i]) w = m.a[i].N == 256 ? 1.0 : (m.a[i].colsum != nullptr && m.a[i].colsum_of_v ? head_w : small_w);
    cost[i] = (double)m.a[i].P * w + 4e3;
    cost_total += cost[i];
  }
  int counts[kBwMaxJobs], given = 0;
  double frac[kBwMaxJobs];
  for (int i = 0; i < k; ++i) {
    const double want = cons_pairs * cost[i] / cost_total;
    counts[i] = (int)want < 1 ? 1 : (int)want;
    frac[i] = want - counts[i];
    given += counts[i];
  }
  while (given < cons_pairs) {
    int best = 0;
    for (int i = 1; i < k; ++i) if (frac[i] > frac[best]) best = i;
    ++counts[best]; frac[best] -= 1.0; ++given;
  }
  while (given > cons_pairs) {
    int worst = -1;
    for (int i = 0; i < k; ++i) if (counts[i] > 1 && (worst < 0 || frac[i] < frac[worst])) worst = i;
    if (worst < 0) return fail_arg(fn, NFS_E_UNSUPPORTED, "more weight-gradient jobs than consumer pairs");
    --counts[worst]; frac[worst] += 1.0; --given;
  }
  unsigned used = 0;
  for (int i = 0; i < k; ++i) {
    const long long units = slabs[i];                              // no more pairs / CTAs than scheduling units (quads)
    long long c = 2LL * counts[i];
    if (!m.paired[i] && c > units + (units & 1)) c = units + (units & 1);
    if (m.paired[i] && c > 2 * units) c = 2 * units;
    m.cta0[i] = used;
    used += (unsigned)c;
  }
  m.cta0[k] = used;
  m.n_jobs = k;
  m.producer_ctas = (unsigned)(2 * prod);
  m.quad_done = quad_flags;
  m.quad_target = 32;
  // back-pressure window (quads): producers stay at most this far ahead of the slowest consumer, so that the dY of a
  // quad (1.8 MB for 7 x 256-wide layers, stored with an L2 evict_last hint) is still in L2 when its consumers load it.
  // Measured with ncu on the cfg 3 step (38 producer pairs, gpurun_out/r2_handoff_matrix.log): window 76 -> the kernel
  // reads 7.7 GB from DRAM, 57 -> 7.0 GB, 48 -> 6.4 GB, at the same duration (2.24-2.29 ms); without the hint 8.9 GB.
  // NFS_BWD_WINDOW=0 switches it off.
  long long window = 5LL * prod / 4;
  if (const char *e = getenv("NFS_BWD_WINDOW")) window = atoll(e);
  int n_wait = 0;
  for (int i = 0; i < k; ++i) n_wait += m.waits[i];
  m.quad_consumed = window > 0 ? quad_flags + n_quads : nullptr;
  m.consumed_target = (unsigned)n_wait;
  m.window = window;
  m.print_times = getenv("NFS_BWD_TIMES") != nullptr;
  unsigned grid = m.producer_ctas + used;
  grid += grid & 1u;                                  // whole clusters (the surplus CTA returns at once)

  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(backward_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  cudaError_t e = cudaMemsetAsync(quad_flags, 0, (size_t)2 * n_quads * sizeof(uint32_t), (cudaStream_t)stream);
  if (e != cudaSuccess) return fail_cuda(fn, e);
  backward_fused_kernel<<<grid, kFmThreads, smem, (cudaStream_t)stream>>>(tx, tw, ts, tb, a, m);
  return check_launch(fn);
}
