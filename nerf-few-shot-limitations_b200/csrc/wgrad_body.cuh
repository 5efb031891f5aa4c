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
#pragma once
#include "tc_common.cuh"
#include "../../include/nfs_b200.h"
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
  int dbg;            // developer bisection switches (NFS_WGRAD_DBG): 1 = no MMAs, 2 = no column sums, 4 = one TMA per operand
  int bulk_drain;     // destination is a contiguous n-major [n_valid x M] block: drain through smem + TMA bulk reduce
  // merged backward kernel (CTA-pair body): operands that the dgrad chain of the same launch produced are dead once their
  // slab has been loaded - their L2 lines are discarded so that they are never written back to DRAM
  const uint8_t *u_ptr, *v_ptr;       // the operands themselves (row-major bf16) and their row pitches in bytes
  long long u_pitch_b, v_pitch_b;
  int discard;                        // bit 0: U, bit 1: V
};

// Slabs of a CTA, in order: units (runs of `U` consecutive slabs) cta, cta + n_cta, ...   U = 1 in the stand-alone
// kernels (finest balance); U = 8 = one 512-row quad in the merged backward kernel, so that a consumer acquires each
// hand-off counter once per eight slabs instead of once per slab.
#define NFS_WG_FOR_SLABS(slab, it)                                                           \
  for (long long unit_ = cta; unit_ * U < n_slabs; unit_ += n_cta)                           \
    for (long long slab = unit_ * U, end_ = (slab + U < n_slabs ? slab + U : n_slabs); slab < end_; ++slab, ++it)

// One CTA's share of one weight-gradient job: CTA `cta` of `n_cta` takes the slab units cta, cta + n_cta, ...
// quad_done (NULL in the stand-alone kernels): the merged backward kernel's hand-off counters - the operands of slab s
// (rows 64 s ...) may be loaded once quad_done[s / 8] has reached `quad_target` (the dgrad chain's epilogue warps have
// completed their stores of that 512-row quad).
__device__ __forceinline__ void wgrad_body(const CUtensorMap *tmap_u_p, const CUtensorMap *tmap_v_p, const WgradArgs &a,
                                           const unsigned cta, const unsigned n_cta,
                                           const unsigned int *quad_done = nullptr, const unsigned quad_target = 0,
                                           unsigned int *quad_consumed = nullptr) {
  const CUtensorMap &tmap_u = *tmap_u_p, &tmap_v = *tmap_v_p;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int mb = a.M >> 6, nb = a.N >> 6, S = a.n_stages;
  const int stage_bytes = (mb + nb) * kBlockBytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * stage_bytes);
  uint64_t *empty = full + S;
  uint64_t *acc_full = empty + S;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long n_slabs = (a.P + kSlabP - 1) / kSlabP;
  const long long U = quad_done != nullptr ? 8 : 1;
  const bool do_colsum = a.colsum != nullptr;
  // column-sum warps: 4 in the stand-alone kernels (6 warps per CTA), 16 in the merged backward kernel (CTAs of 18 warps),
  // where a job confined to a few SMs must take a slab every ~1000 cycles: group g sums the rows [g, g + 1) * 64 / groups
  const int cs_groups = (int)(blockDim.x >> 5) >= 18 ? 4 : 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, do_colsum ? 1 + 4 * cs_groups : 1); }
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
  pdl_trigger();                                   // (nfs_common.cuh: programmatic dependent launch)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      long long quad_seen = -1;
      const uint64_t load_policy = quad_done != nullptr ? l2_policy_evict_first() : 0ull;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        if (quad_done != nullptr && (slab >> 3) != quad_seen) {
          if (quad_consumed != nullptr && quad_seen >= 0) atomicAdd(quad_consumed + quad_seen, 1u);   // back-pressure credit
          // hand-off from the dgrad chain: acquire the quad's counter, then order the TMA (async proxy) loads after it
          const unsigned int *flag = quad_done + (slab >> 3);
          for (uint32_t spin = 0;; ++spin) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v >= quad_target) break;
            __nanosleep(100);
            if (spin > (1u << 22)) {
              printf("nfs_b200: weight-gradient consumer timed out waiting for quad %lld (block %d, have %u of %u)\n",
                     (long long)(slab >> 3), (int)blockIdx.x, v, quad_target);
              __trap();
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");
          quad_seen = slab >> 3;
        }
        mbar_wait(empty + stage, ph ^ 1);
        mbar_expect_tx(full + stage, (uint32_t)stage_bytes);
        uint8_t *us = smem + stage * stage_bytes, *vs = us + mb * kBlockBytes;
        const int row = (int)(slab * kSlabP);
#ifdef NFS_NO_LOAD_HINT
        if (false) {
#else
        if (quad_done != nullptr) {          // merged backward kernel: every operand row is read exactly once
#endif
          for (int b = 0; b < mb; ++b) tma_load_2d_hint(us + b * kBlockBytes, &tmap_u, full + stage, b * 64, row, load_policy);
          for (int b = 0; b < nb; ++b) tma_load_2d_hint(vs + b * kBlockBytes, &tmap_v, full + stage, b * 64, row, load_policy);
        } else {
          for (int b = 0; b < mb; ++b) tma_load_2d(us + b * kBlockBytes, &tmap_u, full + stage, b * 64, row);
          for (int b = 0; b < nb; ++b) tma_load_2d(vs + b * kBlockBytes, &tmap_v, full + stage, b * 64, row);
        }
      }
      if (quad_consumed != nullptr && quad_seen >= 0) atomicAdd(quad_consumed + quad_seen, 1u);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, a.N, 1, 1);   // both operands MN-major
      uint32_t it = 0;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(full + stage, ph);
        tc_fence_after();
        const uint32_t ua = smem_u32(smem + stage * stage_bytes);
        const uint32_t va = ua + mb * kBlockBytes;
        // accumulator-major: the slab's four K steps run back to back on one accumulator (switching the D operand
        // between consecutive MMAs costs ~150 cycles, scripts/ubench/mma_modes.cu)
        for (int h = 0; h < (a.M >> 7); ++h) {
#pragma unroll
          for (int k = 0; k < kSlabP / 16; ++k) {               // 16 points per MMA = two 8-row groups
            const uint64_t bd = umma_desc_sw128(va + k * 2048, kBlockBytes, 1024);
            const uint64_t ad = umma_desc_sw128(ua + h * 2 * kBlockBytes + k * 2048, kBlockBytes, 1024);
            if (!(a.dbg & 1)) umma_bf16(tmem_base + (uint32_t)(h * a.N), ad, bd, idesc, (uint32_t)((it | (uint32_t)k) != 0));
          }
        }
        umma_commit(empty + stage);
      }
      umma_commit(acc_full);
    }
  } else if (warp < 2 + 4 * cs_groups) {
    const int et = (threadIdx.x - 64) & 127;  // 0..127
    const int grp = (warp - 2) >> 2, rows_per = kSlabP / cs_groups;
    if (do_colsum) {
      // thread owns columns 2*et, 2*et+1 of the summed operand (<= 256 columns)
      const int ncols = a.colsum_of_v ? a.N : a.M;
      const int col = 2 * et;
      const bool active = col < ncols;
      const int blk = col >> 6, cc = col & 63;
      float s0 = 0.f, s1 = 0.f;
      uint32_t it = 0;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(full + stage, ph);
        if (active && !(a.dbg & 2)) {
          const uint8_t *base = smem + stage * stage_bytes + (a.colsum_of_v ? mb * kBlockBytes : 0) + blk * kBlockBytes;
#pragma unroll 8
          for (int p = grp * rows_per; p < (grp + 1) * rows_per; ++p) {
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
      if (cs_groups > 1) asm volatile("bar.sync 2, 512;" ::: "memory");   // all sixteen have finished reading the ring
    }
    if (warp >= 6) goto wg_done;              // the four drain warps (TMEM lane quadrants) are warps 2..5
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
        if (threadIdx.x == 64 && n_slabs > (long long)cta * U) {
          const uint32_t bytes = (uint32_t)rows * (uint32_t)M * 4u;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                       ::"l"(a.dw + (long long)n0 * M), "r"(smem_u32(stage)), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (threadIdx.x == 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (n_slabs > (long long)cta * U) {
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

wg_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

}  // namespace
}  // namespace nfs

// Validates one job and fills its tensor maps / arguments; returns 0, or 1 for an empty job, or an error (< 0 / cudaError).
static inline int wgrad_prepare_job(const char *fn, const void *u_bf16, int64_t u_pitch, const void *v_bf16, int64_t v_pitch,
                       int64_t n_points, int32_t m_dim, int32_t n_dim, int32_t m_valid, int32_t n_valid, float *dw,
                       int64_t ld_m, int64_t ld_n, float *colsum, int32_t colsum_of_v, CUtensorMap *tu, CUtensorMap *tv,
                       nfs::WgradArgs *out, size_t *smem) {
  if (n_points < 0 || m_dim <= 0 || n_dim <= 0) return nfs::fail_arg(fn, NFS_E_BADARG, "bad sizes");
  if (n_points == 0) return 1;
  if (!u_bf16 || !v_bf16 || !dw) return nfs::fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (m_dim % 128 != 0 || m_dim > 256 || n_dim % 64 != 0 || n_dim > 256 || (m_dim / 128) * n_dim > 512)
    return nfs::fail_arg(fn, NFS_E_UNSUPPORTED, "need M in {128,256}, N % 64 == 0, N <= 256");
  if (u_pitch < m_dim || v_pitch < n_dim) return nfs::fail_arg(fn, NFS_E_BADARG, "row pitch smaller than the row");
  int rc = nfs::tc::make_tmap_bf16(tu, u_bf16, (uint64_t)n_points, (uint64_t)m_dim, (uint64_t)u_pitch, nfs::kSlabP, fn);
  if (rc) return rc;
  rc = nfs::tc::make_tmap_bf16(tv, v_bf16, (uint64_t)n_points, (uint64_t)n_dim, (uint64_t)v_pitch, nfs::kSlabP, fn);
  if (rc) return rc;
  nfs::WgradArgs a{};
  a.P = n_points; a.M = m_dim; a.N = n_dim; a.dw = dw; a.ld_m = ld_m; a.ld_n = ld_n;
  a.colsum = colsum; a.colsum_of_v = colsum_of_v;
  a.u_ptr = reinterpret_cast<const uint8_t *>(u_bf16); a.v_ptr = reinterpret_cast<const uint8_t *>(v_bf16);
  a.u_pitch_b = u_pitch * 2; a.v_pitch_b = v_pitch * 2; a.discard = 0;
  a.dbg = getenv("NFS_WGRAD_DBG") ? atoi(getenv("NFS_WGRAD_DBG")) : 0;
  a.m_valid = (m_valid <= 0 || m_valid > m_dim) ? m_dim : m_valid;
  a.n_valid = (n_valid <= 0 || n_valid > n_dim) ? n_dim : n_valid;
  const int stage_bytes = ((m_dim + n_dim) / 64) * nfs::kBlockBytes;
  int stages = (227 * 1024 - 1024 - 256) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return nfs::fail_arg(fn, NFS_E_TOOLARGE, "operands too wide for the staging ring");
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

