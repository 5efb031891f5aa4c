// K3 (weight gradients): kernel instantiations + C ABI; the body lives in wgrad_body.cuh (it is also the consumer
// half of the merged backward kernel, backward.cu).
#include "wgrad_pair_body.cuh"

namespace nfs {
namespace {

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_v, const WgradArgs a) {
  wgrad_body(&tmap_u, &tmap_v, a, blockIdx.x, gridDim.x);
}

// CTA-pair form (wgrad_pair_body.cuh): the consumer layout of the merged backward kernel, launchable on its own for
// parity tests and per-SM throughput measurements (NFS_WGRAD_PAIR=1, developer switch).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgThreads, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_v, const WgradArgs a) {
  wgrad_pair_body(&tmap_u, &tmap_v, a, blockIdx.x >> 1, gridDim.x >> 1);
}

// Several independent weight-gradient jobs in ONE launch (the ~21 layers of NeRFWithDINO at a few ten thousand
// points each: a launch per layer is ~25 us of mostly fixed cost).  The CTAs are divided among the jobs in
// proportion to their operand bytes; each CTA then runs exactly the single-job body on its job.
constexpr int kWgMaxJobs = 24;
struct alignas(64) WgradMulti {
  CUtensorMap tu[kWgMaxJobs], tv[kWgMaxJobs];
  WgradArgs a[kWgMaxJobs];
  unsigned cta0[kWgMaxJobs + 1];        // job j owns CTAs cta0[j] .. cta0[j+1] (whole pairs)
  int paired[kWgMaxJobs];               // 1: the job's CTAs work as pairs (wgrad_pair_body), 0: one by one (wgrad_body)
  int n_jobs;
  int print_times;                      // developer switch NFS_WGRAD_TIMES: every CTA prints its job and start / end time
};

__device__ __forceinline__ unsigned long long wg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kWgThreads, 1)
wgrad_multi_kernel(const __grid_constant__ WgradMulti m) {
  if (blockIdx.x >= m.cta0[m.n_jobs]) return;
  int j = 0;
  while (j + 1 < m.n_jobs && blockIdx.x >= m.cta0[j + 1]) ++j;
  const unsigned c = blockIdx.x - m.cta0[j], n = m.cta0[j + 1] - m.cta0[j];
  const unsigned long long t0 = m.print_times ? wg_now() : 0ull;
  if (m.paired[j]) wgrad_pair_body(&m.tu[j], &m.tv[j], m.a[j], c >> 1, n >> 1);
  else wgrad_body(&m.tu[j], &m.tv[j], m.a[j], c, n);
  if (m.print_times && threadIdx.x == 0 && c == 0)
    printf("wgtimes job %d paired %d ctas %u M %d N %d P %lld start %llu end %llu\n", j, m.paired[j], n, m.a[j].M, m.a[j].N,
           m.a[j].P, t0, wg_now());
}

}  // namespace
}  // namespace nfs

using namespace nfs;

static int wgrad_attrs(const char *fn) {
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(fn, e);
    attr_once.mark(attr_dev);
  }
  return 0;
}

extern "C" int nfs_wgrad_bf16(const void *u_bf16, int64_t u_pitch, const void *v_bf16, int64_t v_pitch,
                              int64_t n_points, int32_t m_dim, int32_t n_dim, int32_t m_valid, int32_t n_valid,
                              float *dw, int64_t ld_m, int64_t ld_n, float *colsum, int32_t colsum_of_v, void *stream) {
  const char *fn = "nfs_wgrad_bf16";
  CUtensorMap tu, tv;
  WgradArgs a{};
  size_t smem = 0;
  int rc = wgrad_prepare_job(fn, u_bf16, u_pitch, v_bf16, v_pitch, n_points, m_dim, n_dim, m_valid, n_valid, dw, ld_m, ld_n,
                       colsum, colsum_of_v, &tu, &tv, &a, &smem);
  if (rc == 1) return 0;
  if (rc) return rc;
  if ((rc = wgrad_attrs(fn))) return rc;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_slabs = (n_points + kSlabP - 1) / kSlabP;
  // Every CTA pays a fixed ~12 us (setup, first slab, ~19k cycles of accumulator drain) whatever its share of the
  // points, so small launches want ALL SMs to keep the streaming part short (scripts/dev/wgrad_trace.py).
  long long g = n_slabs < sms ? n_slabs : sms;
  if (const char *force = getenv("NFS_WGRAD_GRID")) g = atoll(force) > 0 && atoll(force) < g ? atoll(force) : g;   // developer switch
  if (getenv("NFS_WGRAD_PAIR") != nullptr && wgrad_pair_ok(a)) {       // developer switch: the CTA-pair body on its own
    long long gp = n_slabs < sms / 2 ? n_slabs : sms / 2;
    if (const char *force = getenv("NFS_WGRAD_GRID")) gp = atoll(force) / 2 > 0 && atoll(force) / 2 < gp ? atoll(force) / 2 : gp;
    wgrad_pair_kernel<<<(unsigned)(2 * gp), kWgThreads, kWpSmemBytes, (cudaStream_t)stream>>>(tu, tv, a);
    return check_launch(fn);
  }
  return launch_dep(fn, wgrad_kernel, dim3((unsigned)g), dim3(kWgThreads), smem, (cudaStream_t)stream, tu, tv, a);
}

extern "C" int nfs_wgrad_multi_bf16(const nfs_wgrad_job *jobs, int32_t n_jobs, void *stream) {
  const char *fn = "nfs_wgrad_multi_bf16";
  if (n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail_arg(fn, NFS_E_BADARG, "bad job list");
  int rc = wgrad_attrs(fn);
  if (rc) return rc;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int first = 0; first < n_jobs;) {
    WgradMulti m{};
    double bytes[kWgMaxJobs];
    long long slabs[kWgMaxJobs];
    double total = 0.0;
    size_t smem = 0;
    int k = 0;
    for (; first < n_jobs && k < kWgMaxJobs && k < sms / 2; ++first) {
      const nfs_wgrad_job &j = jobs[first];
      size_t sm = 0;
      rc = wgrad_prepare_job(fn, j.u_bf16, j.u_pitch, j.v_bf16, j.v_pitch, j.n_points, j.m_dim, j.n_dim, j.m_valid, j.n_valid,
                       j.dw, j.ld_m, j.ld_n, j.colsum, j.colsum_of_v, &m.tu[k], &m.tv[k], &m.a[k], &sm);
      if (rc == 1) continue;
      if (rc) return rc;
      if (sm > smem) smem = sm;
      slabs[k] = (j.n_points + kSlabP - 1) / kSlabP;
      // cost in "points of a 256 x 256 job on a CTA pair" (390 ns per 64-point slab, scripts/dev/wgrad_pair.py), plus
      // the fixed ~12 us a CTA pays whatever its share (setup, first slab, accumulator drain)
      m.paired[k] = wgrad_pair_ok(m.a[k]) && getenv("NFS_WGRAD_NOPAIR") == nullptr;
      if (m.paired[k] && kWpSmemBytes > smem) smem = kWpSmemBytes;
      // CTA by CTA, measured inside this launch at cfg 4 (scripts/dev/wgrad_multi_times.py), in units of the pair body's
      // ~1000 CTA-ns per slab: 128 x 64 0.95, 128 x 128 0.98, 256 x 128 1.27, 256 x 192 1.6
      double w = 0.5 + 1.25 * (j.m_dim + j.n_dim) / 512.0;
      if (m.paired[k]) w = j.n_dim == 256 ? 1.0 : (j.colsum != nullptr && j.colsum_of_v ? 0.97 : 0.67);
      bytes[k] = (double)j.n_points * w + 2000.0;
      total += bytes[k];
      ++k;
    }
    if (k == 0) continue;
    // CTA pairs per job: proportional to its share of the work (largest remainder), at least 1, at most its slab count
    const int pairs = sms / 2;
    int counts[kWgMaxJobs], given = 0;
    double frac[kWgMaxJobs];
    for (int i = 0; i < k; ++i) {
      const double want = pairs * bytes[i] / total;
      counts[i] = (int)want < 1 ? 1 : (int)want;
      frac[i] = want - counts[i];
      given += counts[i];
    }
    while (given < pairs) {
      int best = 0;
      for (int i = 1; i < k; ++i) if (frac[i] > frac[best]) best = i;
      ++counts[best]; frac[best] -= 1.0; ++given;
    }
    while (given > pairs) {
      int worst = -1;
      for (int i = 0; i < k; ++i) if (counts[i] > 1 && (worst < 0 || frac[i] < frac[worst])) worst = i;
      if (worst < 0) break;                       // (more jobs than pairs cannot happen: k <= sms / 2 per launch)
      --counts[worst]; frac[worst] += 1.0; --given;
    }
    unsigned used = 0;
    for (int i = 0; i < k; ++i) {
      long long c = 2LL * counts[i];
      const long long cap = m.paired[i] ? 2 * slabs[i] : slabs[i] + (slabs[i] & 1);     // whole pairs either way
      if (c > cap) c = cap;
      m.cta0[i] = used;
      used += (unsigned)c;
    }
    m.cta0[k] = used;
    m.n_jobs = k;
    m.print_times = getenv("NFS_WGRAD_TIMES") != nullptr;
    rc = launch_dep(fn, wgrad_multi_kernel, dim3(used), dim3(kWgThreads), smem, (cudaStream_t)stream, m);
    if (rc) return rc;
  }
  return 0;
}
