// K3 (weight gradients) — D[M,N] += U[P,M]^T . V[P,N] on tcgen05 tensor cores, with the
// bias gradient (column sums of U or V) folded into the same pass.
//
// Replaces the wgrad half of autograd's Linear backward for
//   nerf_model.NeRFMLP            /root/reference/src/models/nerf_model.py:16-24
//   NeRFWithDINO / NeRFDINOFusion /root/reference/src/models/nerf_mlp.py:134-158,
//                                 /root/reference/src/models/dino_feature_model.py:175-197
// (dW[n,k] = sum_p dY[p,n] X[p,k];  db[n] = sum_p dY[p,n]).
//
// The reduction runs over the POINT index, which is the row index of both row-major operands,
// so both are MN-major UMMA operands: a TMA box of [64 points x 64 columns] lands in shared
// memory as 64 rows of 128 B (SWIZZLE_128B), which read as "K rows x 64 MN elements" is the
// canonical MN-major layout (LBO = distance between 64-column blocks, SBO = 1024 B between
// groups of 8 points).  Persistent CTAs take 64-point slabs round-robin and keep the whole
// [M x N] fp32 partial in TMEM (M/128 accumulators of N columns, <= 512 columns); while the
// MMA thread works, the four epilogue warps add up the bias columns straight from the staged
// slab; at the end they drain TMEM with tcgen05.ld and reduce into global memory with fp32
// red.add (lanes own consecutive m, so m should be the contiguous index of the destination).
#include "tc_common.cuh"
#include <cmath>
#include <cstdlib>

namespace nfs {
namespace {

using namespace tc;

constexpr int kWgThreads = 192;
constexpr int kSlabP = 64;
constexpr int kBlockBytes = kSlabP * 128;   // one [64 points x 64 cols] box

struct WgradArgs {
  long long P;
  int M, N;
  float *dw;
  long long ld_m, ld_n;
  int m_valid, n_valid;   // only m < m_valid, n < n_valid are written (operands are zero-padded)
  float *colsum;      // bias gradient destination | NULL
  int colsum_of_v;    // 1: columns of V (N of them), 0: columns of U (M of them)
  int n_stages, tmem_cols;
  int bulk_drain;     // destination is a contiguous n-major [n_valid x M] block: drain through smem + TMA bulk reduce
};

// One CTA's share of one weight-gradient job: CTA `cta` of `n_cta` takes the slabs cta, cta + n_cta, ...
__device__ __forceinline__ void wgrad_body(const CUtensorMap *tmap_u_p, const CUtensorMap *tmap_v_p, const WgradArgs &a,
                                           const unsigned cta, const unsigned n_cta) {
  const CUtensorMap &tmap_u = *tmap_u_p, &tmap_v = *tmap_v_p;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int mb = a.M >> 6, nb = a.N >> 6, S = a.n_stages;
  const int stage_bytes = (mb + nb) * kBlockBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * stage_bytes);
  uint64_t *empty = full + S;
  uint64_t *acc_full = empty + S;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_slabs = (a.P + kSlabP - 1) / kSlabP;
  const bool do_colsum = a.colsum != nullptr;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, do_colsum ? 5 : 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_u);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (long long slab = cta; slab < n_slabs; slab += n_cta, ++it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(empty + stage, ph ^ 1);
        mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
        uint8_t *us = smem + stage * stage_bytes, *vs = us + mb * kBlockBytes;
        const int row = (int)(slab * kSlabP);
        for (int b = 0; b < mb; ++b) tma_load_2d(us + b * kBlockBytes, &tmap_u, full + stage, b * 64, row);
        for (int b = 0; b < nb; ++b) tma_load_2d(vs + b * kBlockBytes, &tmap_v, full + stage, b * 64, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, a.N, 1, 1);   // both operands MN-major
      uint32_t it = 0;
      for (long long slab = cta; slab < n_slabs; slab += n_cta, ++it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(full + stage, ph);
        tc_fence_after();
        const uint32_t ua = smem_u32(smem + stage * stage_bytes);
        const uint32_t va = ua + mb * kBlockBytes;
#pragma unroll
        for (int k = 0; k < kSlabP / 16; ++k) {                 // 16 points per MMA = two 8-row groups
          const uint64_t bd = umma_desc_sw128(va + k * 2048, kBlockBytes, 1024);
          for (int h = 0; h < (a.M >> 7); ++h) {
            const uint64_t ad = umma_desc_sw128(ua + h * 2 * kBlockBytes + k * 2048, kBlockBytes, 1024);
            umma_bf16(tmem_base + (uint32_t)(h * a.N), ad, bd, idesc, (uint32_t)((it | (uint32_t)k) != 0));
          }
        }
        umma_commit(empty + stage);
      }
      umma_commit(acc_full);
    }
  } else {
    const int et = threadIdx.x - 64;          // 0..127
    if (do_colsum) {
      // thread owns columns 2*et, 2*et+1 of the summed operand (<= 256 columns)
      const int ncols = a.colsum_of_v ? a.N : a.M;
      const int col = 2 * et;
      const bool active = col < ncols;
      const int blk = col >> 6, cc = col & 63;
      float s0 = 0.f, s1 = 0.f;
      uint32_t it = 0;
      for (long long slab = cta; slab < n_slabs; slab += n_cta, ++it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(full + stage, ph);
        if (active) {
          const uint8_t *base = smem + stage * stage_bytes + (a.colsum_of_v ? mb * kBlockBytes : 0) + blk * kBlockBytes;
#pragma unroll 8
          for (int p = 0; p < kSlabP; ++p) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(base + p * 128 + ((((cc >> 3) ^ (p & 7))) << 4) + (cc & 7) * 2);
            s0 += __uint_as_float(w << 16);
            s1 += __uint_as_float(w & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
      }
      const int cvalid = a.colsum_of_v ? a.n_valid : a.m_valid;
      if (active && col < cvalid) atomicAdd(a.colsum + col, s0);
      if (active && col + 1 < cvalid) atomicAdd(a.colsum + col + 1, s1);
    }
    // drain the accumulators
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3;
    if (a.bulk_drain) {
      // The destination is one contiguous fp32 block dw[n * M + m]: transpose the accumulator through shared
      // memory (lanes own consecutive m: conflict-free 128-byte rows) and let the TMA reduce it into global memory
      // (cp.reduce.async.bulk ... add.f32).  The per-thread fp32 atomics of the fallback below are 2048 warp
      // instructions per CTA, ~16k cycles on top of this path for every launch whatever its size
      // (scripts/dev/wgrad_sizes.py, wgrad_trace.py).
      // Rounds of 64 n-rows through two buffers: staging round r + 1 overlaps the TMA's read of round r (the
      // bulk reduction drains shared memory at ~24 B/cycle: 2 650 cycles per 64 KB round, staging takes 1 400).
      const int M = a.M;
      float *stage0 = reinterpret_cast<float *>(smem);       // the operand ring is idle: every MMA has completed
      asm volatile("bar.sync 1, 128;" ::: "memory");         // every drain warp has finished its column sums (they read the ring)
      int round = 0;
      for (int n0 = 0; n0 < a.n_valid; n0 += 64, ++round) {
        float *stage = stage0 + (round & 1) * 64 * M;
        if (round >= 2) {                                    // the reduction issued two rounds ago has read this buffer
          if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const int rows = min(64, a.n_valid - n0);
        for (int h = 0; h < (M >> 7); ++h) {
          const int m = h * 128 + q * 32 + lane;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * a.N + n0);
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[(c0 + j) * M + m] = v[j];
          }
        }
        fence_proxy_async();                                 // generic-proxy writes -> visible to the bulk copy
        asm volatile("bar.sync 1, 128;" ::: "memory");       // the four drain warps
        if (threadIdx.x == 64 && n_slabs > (long long)cta) {
          const uint32_t bytes = (uint32_t)rows * (uint32_t)M * 4u;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                       ::"l"(a.dw + (long long)n0 * M), "r"(smem_u32(stage)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (n_slabs > (long long)cta) {
      for (int h = 0; h < (a.M >> 7); ++h) {
        const long long m = h * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * a.N);
        for (int c0 = 0; c0 < a.N; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          float *dst = a.dw + m * a.ld_m + (long long)c0 * a.ld_n;
          if (m < a.m_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < a.n_valid) atomicAdd(dst + j * a.ld_n, v[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmap_u, const __grid_constant__ CUtensorMap tmap_v, const WgradArgs a) {
  wgrad_body(&tmap_u, &tmap_v, a, blockIdx.x, gridDim.x);
}

// Several independent weight-gradient jobs in ONE launch (the ~21 layers of NeRFWithDINO at a few ten thousand
// points each: a launch per layer is ~25 us of mostly fixed cost).  The CTAs are divided among the jobs in
// proportion to their operand bytes; each CTA then runs exactly the single-job body on its job.
constexpr int kWgMaxJobs = 24;
struct alignas(64) WgradMulti {
  CUtensorMap tu[kWgMaxJobs], tv[kWgMaxJobs];
  WgradArgs a[kWgMaxJobs];
  unsigned cta0[kWgMaxJobs + 1];        // job j owns CTAs cta0[j] .. cta0[j+1]
  int n_jobs;
};

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_multi_kernel(const __grid_constant__ WgradMulti m) {
  int j = 0;
  while (j + 1 < m.n_jobs && blockIdx.x >= m.cta0[j + 1]) ++j;
  wgrad_body(&m.tu[j], &m.tv[j], m.a[j], blockIdx.x - m.cta0[j], m.cta0[j + 1] - m.cta0[j]);
}

}  // namespace
}  // namespace nfs

using namespace nfs;

// Validates one job and fills its tensor maps / arguments; returns 0, or 1 for an empty job, or an error (< 0 / cudaError).
static int prepare_job(const char *fn, const void *u_bf16, int64_t u_pitch, const void *v_bf16, int64_t v_pitch,
                       int64_t n_points, int32_t m_dim, int32_t n_dim, int32_t m_valid, int32_t n_valid, float *dw,
                       int64_t ld_m, int64_t ld_n, float *colsum, int32_t colsum_of_v, CUtensorMap *tu, CUtensorMap *tv,
                       WgradArgs *out, size_t *smem) {
  if (n_points < 0 || m_dim <= 0 || n_dim <= 0) return fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 1;
  if (!u_bf16 || !v_bf16 || !dw) return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (m_dim % 128 != 0 || m_dim > 256 || n_dim % 64 != 0 || n_dim > 256 || (m_dim / 128) * n_dim > 512)
    return fail_arg(fn, NFS_E_UNSUPPORTED, "need M in {128,256}, N % 64 == 0, N <= 256");
  if (u_pitch < m_dim || v_pitch < n_dim) return fail_arg(fn, NFS_E_BADARG, "row pitch smaller than the row");
  int rc = tc::make_tmap_bf16(tu, u_bf16, (uint64_t)n_points, (uint64_t)m_dim, (uint64_t)u_pitch, kSlabP, fn);
  if (rc) return rc;
  rc = tc::make_tmap_bf16(tv, v_bf16, (uint64_t)n_points, (uint64_t)n_dim, (uint64_t)v_pitch, kSlabP, fn);
  if (rc) return rc;
  WgradArgs a{};
  a.P = n_points; a.M = m_dim; a.N = n_dim; a.dw = dw; a.ld_m = ld_m; a.ld_n = ld_n;
  a.colsum = colsum; a.colsum_of_v = colsum_of_v;
  a.m_valid = (m_valid <= 0 || m_valid > m_dim) ? m_dim : m_valid;
  a.n_valid = (n_valid <= 0 || n_valid > n_dim) ? n_dim : n_valid;
  const int stage_bytes = ((m_dim + n_dim) / 64) * kBlockBytes;
  int stages = (227 * 1024 - 1024 - 256) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return fail_arg(fn, NFS_E_TOOLARGE, "operands too wide for the staging ring");
  a.n_stages = stages;
  const size_t ring = (size_t)stages * stage_bytes;
  const size_t round_bytes = (size_t)2 * 64 * m_dim * 4;          // two 64-row staging buffers
  a.bulk_drain = ld_m == 1 && ld_n == m_dim && a.m_valid == m_dim && round_bytes <= ring &&
                 (reinterpret_cast<uintptr_t>(dw) & 15u) == 0 && getenv("NFS_WGRAD_ATOMIC_DRAIN") == nullptr;
  int cols = 32;
  while (cols < (m_dim / 128) * n_dim) cols <<= 1;
  a.tmem_cols = cols;
  *smem = 1024 + (size_t)stages * stage_bytes + 256;
  *out = a;
  return 0;
}

static int wgrad_attrs(const char *fn) {
  static PerDeviceOnce attr_once;
  int attr_dev = 0;
  if (attr_once.need(&attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
  int rc = prepare_job(fn, u_bf16, u_pitch, v_bf16, v_pitch, n_points, m_dim, n_dim, m_valid, n_valid, dw, ld_m, ld_n,
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
  wgrad_kernel<<<(unsigned)g, kWgThreads, smem, (cudaStream_t)stream>>>(tu, tv, a);
  return check_launch(fn);
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
    for (; first < n_jobs && k < kWgMaxJobs && k < sms; ++first) {
      const nfs_wgrad_job &j = jobs[first];
      size_t sm = 0;
      rc = prepare_job(fn, j.u_bf16, j.u_pitch, j.v_bf16, j.v_pitch, j.n_points, j.m_dim, j.n_dim, j.m_valid, j.n_valid,
                       j.dw, j.ld_m, j.ld_n, j.colsum, j.colsum_of_v, &m.tu[k], &m.tv[k], &m.a[k], &sm);
      if (rc == 1) continue;
      if (rc) return rc;
      if (sm > smem) smem = sm;
      slabs[k] = (j.n_points + kSlabP - 1) / kSlabP;
      bytes[k] = (double)j.n_points * (j.m_dim + j.n_dim) * 2.0 + 8e5;    // + the fixed cost of a CTA, in byte-equivalents
      total += bytes[k];
      ++k;
    }
    if (k == 0) continue;
    // CTAs per job: proportional to its share of the work, at least 1, at most its slab count
    unsigned used = 0;
    for (int i = 0; i < k; ++i) {
      long long c = (long long)(sms * bytes[i] / total);
      if (c < 1) c = 1;
      if (c > slabs[i]) c = slabs[i];
      m.cta0[i] = used;
      used += (unsigned)c;
    }
    m.cta0[k] = used;
    m.n_jobs = k;
    wgrad_multi_kernel<<<used, kWgThreads, smem, (cudaStream_t)stream>>>(m);
    rc = check_launch(fn);
    if (rc) return rc;
  }
  return 0;
}
