// K3 (weight gradients on CTA pairs) — D[256,256] += U[P,256]^T . V[P,256] with tcgen05 cta_group::2, the bias gradient
// on the tensor core as well.
//
// Same contraction as wgrad_body.cuh (dW[n,k] = sum_p dY[p,n] X[p,k], db[n] = sum_p dY[p,n]: the wgrad half of autograd's
// Linear backward, /root/reference/src/models/nerf_model.py:16-24), laid out for a job that is confined to a FEW SMs -
// the consumer side of the merged backward kernel (backward.cu), where a layer's weight gradient gets 6-8 SMs and each
// of them has to take a 64-point slab every ~1000 cycles.  The single-CTA body is bound by shared-memory bandwidth
// there: per slab the TMA writes 64 KB, the MMAs read 96 KB (V twice, once per 128-row accumulator) and the column-sum
// warps another 32 KB while the tensor pipe starves their LDS (measured 57 GB/s per SM against the 121 GB/s the MMAs
// would allow).  Here the pair shares one slab:
//   * each CTA loads HALF of the columns of both operands (U[:, 128 r ..], V[:, 128 r ..]: 2 + 2 boxes of 8 KB), the
//     MMA 256 x 256 x 16 takes the A rows (U^T) of both CTAs and each CTA's half of B (V): 32 KB written and 32 KB
//     read per CTA and slab, 512 cycles of tensor time per slab for the pair;
//   * the column sums are one more MMA per 16 points on the spare TMEM columns: A = the operand whose columns are
//     summed, read MN-major exactly as the main MMA reads it, B = a constant block of ones (N = 16; every core matrix
//     aliases the same 128 bytes) -> D2[m, 0] = sum_p X[p, m].  No LDS at all while the ring is running;
//   * each CTA drains its 128 accumulator rows through shared memory with the TMA bulk reduction, as the single-CTA
//     body does.
// N = 64 (the head's and the first layer's jobs) runs as N = 128: CTA 0 loads the 64 real columns of V, CTA 1's box lies
// outside the tensor and arrives as zeros (no memory traffic); column sums of such a V (64 columns cannot be an M = 256
// operand) are taken by CTA 0's drain warps from shared memory - 8 KB per slab.
// Restrictions (everything else takes the single-CTA body): M = 256, N = 256 or 64, contiguous n-major destination
// (ld_m = 1, ld_n = 256), m_valid = 256.
#pragma once
#include "wgrad_body.cuh"

namespace nfs {
namespace {

constexpr int kWpStageBytes = 4 * kBlockBytes;     // this CTA's halves of U and V for one 64-point slab
constexpr int kWpStages = 6;
constexpr size_t kWpSmemBytes = 1024 + (size_t)kWpStages * kWpStageBytes + 512;

__host__ __device__ inline bool wgrad_pair_ok(const WgradArgs &a) {
  return a.M == 256 && (a.N == 256 || a.N == 64) && a.bulk_drain && a.m_valid == 256 && !(a.dbg & 8);
}

// `pair` of `n_pairs`: this cluster's share of the job (slab units pair, pair + n_pairs, ...).  quad_done /
// quad_target / quad_consumed as in wgrad_body.
__device__ __forceinline__ void wgrad_pair_body(const CUtensorMap *tmap_u_p, const CUtensorMap *tmap_v_p, const WgradArgs &a,
                                                const unsigned pair, const unsigned n_pairs,
                                                const unsigned int *quad_done = nullptr, const unsigned quad_target = 0,
                                                unsigned int *quad_consumed = nullptr) {
  const CUtensorMap &tmap_u = *tmap_u_p, &tmap_v = *tmap_v_p;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int S = kWpStages;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * kWpStageBytes);   // leader: both CTAs' loads of the stage landed
  uint64_t *empty = full + S;                                                 // both: the stage's MMAs have completed
  uint64_t *acc_full = empty + S;                                             // both: every MMA has completed
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
  uint8_t *s_ones = reinterpret_cast<uint8_t *>(full) + 256;                  // 128 B: one core matrix of bf16 ones

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_slabs = (a.P + kSlabP - 1) / kSlabP;
  const long long U = quad_done != nullptr ? 8 : 1;
  const unsigned cta = pair, n_cta = n_pairs;                                 // (names used by NFS_WG_FOR_SLABS)
  const bool do_colsum = a.colsum != nullptr;
  const bool has_work = n_slabs > (long long)pair * U;
  const bool narrow = a.N == 64;                          // V: one box per CTA (CTA 1's is out of bounds = zeros)
  const bool cs_lds = do_colsum && narrow && a.colsum_of_v;   // column sums by CTA 0's drain warps instead of the MMA
  const uint32_t stage_tx = narrow ? 3u * kBlockBytes : 4u * kBlockBytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, cs_lds && rank == 0 ? 5 : 1); }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_u);
    tma_prefetch_desc(&tmap_v);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 72) {
    *reinterpret_cast<uint4 *>(s_ones + (threadIdx.x - 64) * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
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
          if (quad_consumed != nullptr && quad_seen >= 0 && rank == 0) atomicAdd(quad_consumed + quad_seen, 1u);
          const unsigned int *flag = quad_done + (slab >> 3);
          for (uint32_t spin = 0;; ++spin) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v >= quad_target) break;
            __nanosleep(100);
            if (spin > (1u << 22)) {
              printf("nfs_b200: weight-gradient consumer pair timed out waiting for quad %lld (block %d, have %u of %u)\n",
                     (long long)(slab >> 3), (int)blockIdx.x, v, quad_target);
              __trap();
            }
          }
          asm volatile("fence.proxy.async;" ::: "memory");
          quad_seen = slab >> 3;
        }
        mbar_wait(empty + stage, ph ^ 1);
        if (rank == 0) mbar_expect_tx(full + stage, 2u * stage_tx);
        uint8_t *us = smem + stage * kWpStageBytes, *vs = us + 2 * kBlockBytes;
        const int row = (int)(slab * kSlabP), col = (int)rank * 128;
#ifdef NFS_NO_LOAD_HINT
        if (false) {
#else
        if (quad_done != nullptr) {          // merged backward kernel: every operand row is read exactly once
#endif
          tma_load_2d_pair_hint(us, &tmap_u, full + stage, col, row, load_policy);
          tma_load_2d_pair_hint(us + kBlockBytes, &tmap_u, full + stage, col + 64, row, load_policy);
          if (narrow) {
            tma_load_2d_pair_hint(vs, &tmap_v, full + stage, (int)rank * 64, row, load_policy);
          } else {
            tma_load_2d_pair_hint(vs, &tmap_v, full + stage, col, row, load_policy);
            tma_load_2d_pair_hint(vs + kBlockBytes, &tmap_v, full + stage, col + 64, row, load_policy);
          }
        } else {
          tma_load_2d_pair(us, &tmap_u, full + stage, col, row);
          tma_load_2d_pair(us + kBlockBytes, &tmap_u, full + stage, col + 64, row);
          if (narrow) {
            tma_load_2d_pair(vs, &tmap_v, full + stage, (int)rank * 64, row);
          } else {
            tma_load_2d_pair(vs, &tmap_v, full + stage, col, row);
            tma_load_2d_pair(vs + kBlockBytes, &tmap_v, full + stage, col + 64, row);
          }
        }
      }
      if (quad_consumed != nullptr && quad_seen >= 0 && rank == 0) atomicAdd(quad_consumed + quad_seen, 1u);
    }
  } else if (warp == 1) {
    if (rank == 0) {                 // all 32 lanes walk the (warp-uniform) schedule; one elected lane issues
      const uint32_t idesc = umma_idesc_bf16(256, narrow ? 128 : 256, 1, 1);        // both operands MN-major
      const uint32_t idesc_cs = umma_idesc_bf16(256, 16, 1, 0);      // column sums: A MN-major, B = ones (K-major, no swizzle)
      const uint64_t ones_desc = umma_desc_noswizzle(smem_u32(s_ones), 0, 0);
      const uint32_t d_main = tmem_base, d_cs = tmem_base + 256u;
      const uint32_t cs_off = a.colsum_of_v ? 2 * kBlockBytes : 0;
      uint32_t it = 0;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait(full + stage, ph);
        tc_fence_after();
        const uint32_t ua = smem_u32(smem + stage * kWpStageBytes);
        const uint32_t va = ua + 2 * kBlockBytes;
        if (elect_one()) {
          if (!(a.dbg & 1)) {
#pragma unroll
            for (int k = 0; k < kSlabP / 16; ++k)                    // 16 points per MMA = two 8-row groups
              umma_bf16_pair(d_main, umma_desc_sw128(ua + k * 2048, kBlockBytes, 1024),
                             umma_desc_sw128(va + k * 2048, kBlockBytes, 1024), idesc, (uint32_t)((it | (uint32_t)k) != 0));
          }
          if (do_colsum && !cs_lds && !(a.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < kSlabP / 16; ++k)
              umma_bf16_pair(d_cs, umma_desc_sw128(ua + cs_off + k * 2048, kBlockBytes, 1024), ones_desc, idesc_cs,
                             (uint32_t)((it | (uint32_t)k) != 0));
          }
          umma_commit_pair(empty + stage);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit_pair(acc_full);
      __syncwarp();
    }
  } else if (warp < 6) {
    // drain: this CTA's 128 accumulator rows m = 128 rank + (TMEM lane) -> dw[n * 256 + m]
    const int q = warp & 3, ml = q * 32 + lane;
    float cs_sum = 0.f;
    if (cs_lds && rank == 0) {
      // thread = (column et & 63 of V, half et >> 6 of the slab's rows); the full barrier lives in this (leader) CTA
      const int et = threadIdx.x - 64, c = et & 63, p0 = (et >> 6) * 32;
      uint32_t it = 0;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait_relaxed(full + stage, ph);
        if (!(a.dbg & 2)) {
          const uint8_t *base = smem + stage * kWpStageBytes + 2 * kBlockBytes + (c & 7) * 2;
#pragma unroll 8
          for (int p = p0; p < p0 + 32; ++p) {
            const uint32_t w = *reinterpret_cast<const uint16_t *>(base + p * 128 + (((c >> 3) ^ (p & 7)) << 4));
            cs_sum += __uint_as_float(w << 16);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
      }
      if (c < a.n_valid) atomicAdd(a.colsum + c, cs_sum);
    }
    if (a.discard != 0 && quad_done != nullptr && !(cs_lds && rank == 0)) {
      // dY rows of the chain are dead once this pair has loaded them (one job reads each plane): discard their L2 lines
      // as soon as the slab's MMAs have completed (`empty`: the loads have landed long before), so that the dirty lines
      // are dropped instead of written back - the 4.2 GB per step of DRAM writes that remained of the dY round trip.
      // One 128-byte line per thread and slab: this CTA loaded columns [128 rank, 128 rank + 128) = 2 lines of 64 rows.
      const int et = threadIdx.x - 64, drow = et >> 1, dline = et & 1;
      const uint8_t *base = (a.discard & 1) ? a.u_ptr : a.v_ptr;
      const long long pitch = (a.discard & 1) ? a.u_pitch_b : a.v_pitch_b;
      uint32_t it = 0;
      NFS_WG_FOR_SLABS(slab, it) {
        const uint32_t stage = it % S, ph = (it / S) & 1;
        mbar_wait_relaxed(empty + stage, ph);
        const long long row = slab * kSlabP + drow;
        if (row < a.P)
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + row * pitch + (long long)rank * 256 + dline * 128) : "memory");
      }
    }
    mbar_wait_relaxed(acc_full, 0);
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");            // every drain warp has finished its column sums: they read the ring
    if (has_work) {
      float *stage0 = reinterpret_cast<float *>(smem);       // the operand ring is idle: every MMA has completed
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      int round = 0;
      for (int n0 = 0; n0 < a.n_valid; n0 += 64, ++round) {
        float *stage = stage0 + (round & 1) * 64 * 128;
        if (round >= 2) {                                    // the reductions issued two rounds ago have read this buffer
          if (warp == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const int rows = min(64, a.n_valid - n0);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + (uint32_t)(n0 + c0), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) stage[(c0 + j) * 128 + ml] = v[j];
        }
        fence_proxy_async();                                 // generic-proxy writes -> visible to the bulk copy
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2) {                                     // 64 rows of 512 B: two per lane, one bulk group per lane and round
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = lane + 32 * i;
            if (r < rows)
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                           ::"l"(a.dw + (long long)(n0 + r) * 256 + (long long)rank * 128), "r"(smem_u32(stage + r * 128)),
                             "r"(512u) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (warp == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (do_colsum && !cs_lds) {
        float v[16];
        tmem_ld16(taddr + 256u, v);
        const int idx = (int)rank * 128 + ml;
        const int cvalid = a.colsum_of_v ? a.n_valid : a.m_valid;
        if (idx < cvalid) atomicAdd(a.colsum + idx, v[0]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the leader's MMAs read the peer's operands / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace
}  // namespace nfs
