// K3 (fused multi-layer MLP, forward and dgrad) — a whole chain of dense layers per launch, activations
// kept in shared memory / TMEM, weights streamed by TMA, every layer on tcgen05 tensor cores.
//
// Replaces the Linear(+ReLU) chain + heads of
//   nerf_model.NeRFMLP.forward    /root/reference/src/models/nerf_model.py:16-24
// (and any sub-chain of nerf_mlp.NeRFWithDINO, nerf_mlp.py:134-158) for 256 points per CTA step.
//
// Persistent CTA PAIRS (thread-block clusters of 2 = two SMs, tcgen05 cta_group::2): each CTA keeps two
// 128-point tiles (A, B) in flight, so a pair works on four tiles.
//   * act[A], act[B]  : 2 x 64 KB shared memory per CTA, the bf16 activations of the current layer in
//                       the canonical K-major SWIZZLE_128B operand layout (4 slabs of [128 x 64]);
//                       the epilogue overwrites them IN PLACE with the next layer's input;
//   * weight ring     : 6 x 16 KB stages per CTA.  One tcgen05.mma.cta_group::2 multiplies the pair's
//                       256 points (128 from each CTA) by all N output columns, and each CTA supplies
//                       HALF of the weight rows: a stage is [N/2 x 64], a whole layer is 4 stages, so
//                       the layer's weights stay resident while tile A and then tile B use them and
//                       the next layer's first stages are prefetched.  Per-SM shared-memory traffic per
//                       MMA drops from 12 KB to 8 KB and the MMA runs at its 128-cycle floor
//                       (scripts/ubench/mma_2cta.cu; the single-CTA form is smem-bound at 161 cycles);
//   * TMEM            : two 128 x 256 fp32 accumulators (512 columns) per CTA, one per tile.
// The leader CTA's MMA thread runs tile-major - 16 MMAs on the pair's A tiles, 16 on the B tiles - so
// A's accumulator completes half a layer (2048 cycles) before B's: A's epilogue (tcgen05.ld -> +bias ->
// ReLU -> bf16 -> swizzled st.shared) overlaps B's MMAs and B's epilogue overlaps A's MMAs of the next
// layer.  Barriers that gate the MMA thread (weights landed, tile ready, accumulator drained) live in
// the leader CTA and count both CTAs (TMA .cta_group::2 completion, remote mbarrier arrives); barriers
// the MMA thread signals (accumulator full, weight stage free) are multicast to both CTAs by
// tcgen05.commit.
// Three instantiations: inference forward (optionally computing the positional encoding of the points
// in-kernel, K2 fused in), the forward of a training step (each epilogue warp TMA-stores the 32-row x
// 64-column boxes it has just written, for wgrad, and writes one ReLU sign bit per activation, for the
// backward), and the dgrad chain (ReLU backward from those sign bits).
// The last layer of a forward chain is a narrow head (N = 64 padded) whose first `out_cols` columns are
// written as fp32 with the reference's output activation.
// Warp roles per CTA: 0 = TMA producer, 1 = TMEM owner (+ MMA issuer in the leader), 2..17 = epilogue
// (four warps per TMEM lane quadrant, each draining a quarter of the columns of tile A, then of tile B).
#pragma once
#include "tc_common.cuh"

namespace nfs {
namespace {

using namespace tc;

constexpr int kFmThreads = 576;   // 18 warps
constexpr int kFmMaxLayers = 12;
constexpr int kActBytes = 128 * 256 * 2;
constexpr int kActSlab = 128 * 128;
constexpr int kWStage = 128 * 128;   // one weight stage: this CTA's half of the rows, [<= 128 x 64] bf16
constexpr int kWStages = 6;

struct FusedArgs {
  long long P;
  long long save_rows;                 // rows per layer in the saved-activation tensor (P rounded up to 128)
  int n_layers;
  int K[kFmMaxLayers], N[kFmMaxLayers], act[kFmMaxLayers], row0[kFmMaxLayers];
  int has_bias;                        // the biases arrive as a tensor-core operand (tmap_b), see below
  int x_save;                          // in-kernel encoding of a training forward: tmap_x stores the encoded operand
  float *out;                          // [P, out_cols] fp32
  int out_cols;
  int save;
  int head;                            // 1: last layer is the fp32 output head; 0: it is a regular (saved) layer
  // ReLU mask bits, 256 per row = 8 words [layer][row][8]; word w covers columns 32w .. 32w+31 as 16 packed pairs: bit
  // 15 - j (j < 16) is column 32w + 2j, bit 31 - j is column 32w + 2j + 1 (the two halves of packed pair j; 1 = the
  // pre-activation was positive), so (word >> (15 - j)) & 0x10001 times 0x3F80 is the pair's bf16x2 {1.0 | 0.0}
  // multiplier.  bits_out: written by layers with act 1 (the forward chain of a training step) - three instructions per
  // packed pair (push_mask); bits_in: read by layers with act 4 (the dgrad chain):
  // result *= bit(mask_idx[l], row, col).
  uint32_t *bits_out;
  const uint32_t *bits_in;
  long long bits_rows;                 // rows per layer of bits_in
  int mask_idx[kFmMaxLayers];
  const float *points;                 // in-kernel encoding mode: [P,3] fp32 sample positions (x_bf16 unused) | NULL
  // in-kernel SAMPLER mode (points then aliases rays_o and only marks the encoding mode): point p = sample p % ray_S of
  // ray p / ray_S, position = rays_o + rays_d * z_vals[p] evaluated here (un-contracted, as the sampler kernels do)
  const float *ray_z, *rays_o, *rays_d;
  int ray_S;
  float freq0;                         // first frequency band; band k = freq0 * 2^k
  int n_octaves;                       // number of bands (3 * (2 * n_octaves + 1) <= 64)
  int dbg;                             // developer bisection switches (nfs_set_debug_flags), 0 in production
  unsigned long long *trace;           // developer timeline of CTA 0 (nfs_set_debug_trace), NULL in production
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap *map, const void *src, int c_inner, int c_outer, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(map), "r"(smem_u32(src)), "r"(c_inner), "r"(c_outer), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float fm_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// bf16x2 pack with the ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// ReLU backward on packed pair j (0..15) of a 32-column word of mask bits (layout in FusedArgs)
__device__ __forceinline__ uint32_t relu_bits_bf16x2(uint32_t v, uint32_t word, int j) {
  const uint32_t m = ((word >> (15 - j)) & 0x00010001u) * 0x3F80u;     // bf16x2 {1.0 | 0.0, 1.0 | 0.0}
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(m));
  return r;
}
// Mask bits of a post-ReLU packed pair shifted into `acc` from the right, both 16-bit lanes at once: one bf16x2 compare
// (0xFFFF per half that is > 0 - exactly torch's relu'(x) = [x > 0], also for +-0), one shift, one LOP3.  After the 16
// pairs of a word, pair j's bits sit at positions 15 - j (even column) and 31 - j (odd column).
__device__ __forceinline__ uint32_t push_mask(uint32_t acc, uint32_t pk) {
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
  const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&pk), zero);
  return (acc << 1) | (m & 0x00010001u);
}

// TMEM -> registers, 16 consecutive fp32 columns of this thread's lane, WITHOUT waiting: the
// registers are valid after the next tmem_ld_wait().  Lets the load of chunk i+1 fly while chunk i
// is being processed.
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One epilogue chunk: 16 accumulator columns of one row (the bias is already in the accumulator, see the bias
// MMA) -> ReLU | ReLU-backward mask | none
// -> bf16 -> two 16-byte chunks of the row in the SWIZZLE_128B operand layout.
template <bool kMasked, bool kTrain>
__device__ __forceinline__ void epi_chunk(uint32_t (&r)[16], int act, uint32_t bits_word, int j0,
                                          uint32_t &bits_acc, uint8_t *srow, int ch, int r7, int dbg) {
  uint32_t pk[8];
  if (act == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
  }
  if (kMasked && act == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) pk[j] = relu_bits_bf16x2(pk[j], bits_word, j0 + j);
  }
  if (kTrain && act == 1) {              // mask bits for the backward pass
    // pair j0 + j of the word goes to bits 15 - (j0 + j) (even column) and 31 - (j0 + j) (odd column): the compare
    // yields 0xFFFF per positive half, so ONE LOP3 per pair (mask & positioned constant, OR-ed in) places both bits,
    // on two independent accumulators.  (The earlier shift-and-insert form was a chain of 16 dependent instructions
    // per chunk in a step that is on the layer's critical path: scripts/trace_fused.py, 1808 vs 988 cycles.)
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
    uint32_t part0 = 0u, part1 = 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&pk[j]), zero);
      const uint32_t c = 0x00010001u << (15 - (j0 + j));
      if (j & 1) part1 |= m & c; else part0 |= m & c;
    }
    bits_acc |= part0 | part1;
  }
  if (dbg & 32) {                       // bisection: keep the values alive without the shared-memory stores
    uint32_t x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) x ^= pk[j];
    if (x == 0x9e3779b9u) *reinterpret_cast<uint32_t *>(srow) = x;
    return;
  }
  *reinterpret_cast<uint4 *>(srow + ((ch ^ r7) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4 *>(srow + (((ch + 1) ^ r7) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
}

// Developer tools (bisection switches + timeline) are compiled in only with -DNFS_DEVTOOLS
// (NFS_DEVTOOLS=1 python -m nfs_b200.build): they cost registers in a kernel that is short of them.
#ifndef NFS_DEVTOOLS
#define NFS_TRACE(code, l, t) do { } while (0)
#define NFS_DBG(a) 0
#else
#define NFS_DBG(a) ((a).dbg)
// Developer timeline: CTAs 0 and 1 (one pair) append (time << 16 | code << 12 | layer << 4 | tile) per warp; time =
// clock64 of the SM, or (debug flag 64) the global nanosecond timer, which both CTAs of the pair share.
__device__ __forceinline__ unsigned long long fm_trace_time(int dbg) {
  unsigned long long t;
  if (dbg & 64) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  else t = (unsigned long long)clock64();
  return t;
}
#define NFS_TRACE(code, l, t)                                                                       \
  do {                                                                                              \
    if (a.trace != nullptr && blockIdx.x < 2 && lane == 0 && trace_n < 1023) {                      \
      unsigned long long *tb_ = a.trace + blockIdx.x * 18 * 1024 + warp * 1024;                     \
      tb_[1 + trace_n++] = (fm_trace_time(a.dbg) << 16) | ((unsigned long long)(code) << 12) | ((l) << 4) | (t); \
      tb_[0] = trace_n;                                                                             \
    }                                                                                               \
  } while (0)
#endif

// kMasked: dgrad chain (ReLU backward from sign bits).  kTrain: forward chain of a training step (saves activations
// and sign bits).  Neither: inference forward (optionally with the positional encoding computed in-kernel).
// pair0 / pair_step: this CTA pair works on quads pair0, pair0 + pair_step, ...  quad_done (NULL in the stand-alone
// chain kernels): one counter per quad, incremented by every epilogue warp of the pair (32 arrivals) once ALL of the
// quad's TMA stores have completed - the hand-off to the weight-gradient consumers of the merged backward kernel.
// (Two other ways of getting the saved rows out were built and measured on a same-box A/B and dropped: ONE store thread per
// CTA issuing all TMA stores, the epilogue warps only arriving on / waiting for mbarriers: training forward 1.06-1.11 vs
// 0.81 ms, dgrad 1.0 vs 0.78 ms - a tile's 64 KB then has to be READ by the TMA before the tile's next layer may
// overwrite it, which takes longer than the 2000 cycles between them, while per-warp 4 KB stores overlap it
// (gpurun_out/r2_ab_libs_storethread.log); and st.global straight from the epilogue's registers for half / all of the
// warps: 1.24 / 1.88 vs 0.83 ms (gpurun_out/r2_ab_direct_store.log).)
// (Issuing the stores from the elected lane with uniform operands removes the R2UR waterfall loop the compiler builds around
// the UTMASTG inside `if (lane == 0)` - its exit branch shows 6 % of the training forward's stall samples - but the other
// lanes only wait there for lane 0's arrive + store: no change on a same-box A/B, gpurun_out/r2_ab_libs_elect.log.)
// (A finer hand-off - one counter per quad AND layer, published two layers behind by one thread per CTA - was built and
// measured: same duration, DRAM reads 5.1 instead of 6.4 GB per step, but it failed the step's gradient-equality test
// in one configuration and was dropped; gpurun_out/r2_handoff_matrix.log.)
template <bool kMasked, bool kTrain>
__device__ __forceinline__ void chain_body(const CUtensorMap *tmap_x_p, const CUtensorMap *tmap_w_p,
                                           const CUtensorMap *tmap_save_p, const CUtensorMap *tmap_b_p,
                                           const FusedArgs &a, const long long quad0, const long long quad_step,
                                           unsigned int *quad_done, const unsigned int *quad_consumed = nullptr,
                                           const unsigned consumed_target = 0, const long long window = 0) {
  const CUtensorMap &tmap_x = *tmap_x_p, &tmap_w = *tmap_w_p, &tmap_save = *tmap_save_p, &tmap_b = *tmap_b_p;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = smem_raw;                      // SWIZZLE_128B operands need 1024-byte alignment (checked below)
  uint8_t *act[2] = {smem, smem + kActBytes};
  uint8_t *wring = smem + 2 * kActBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(wring + kWStages * kWStage);
  uint64_t *w_full = bars, *w_empty = bars + kWStages;
  uint64_t *in_full = bars + 2 * kWStages;       // [2] input operand of the chain landed (TMA)
  uint64_t *act_free = in_full + 2;              // [2] last layer's MMAs have read act[t]
  uint64_t *act_ready = act_free + 2;            // [2] epilogue wrote next layer's operand into act[t]
  uint64_t *acc_full = act_ready + 2;            // [2] accumulator of tile t complete
  uint64_t *head_done = acc_full + 2;            // [2] this CTA's epilogue is done with act[t] / accumulator t (producer)
  uint64_t *acc_free = head_done + 2;            // [2] leader: both CTAs' epilogues drained accumulator t (MMA thread)
  uint64_t *in_ready = acc_free + 2;             // [2] leader: both CTAs' encoding warps wrote the chain input of tile t
  uint64_t *bias_full = in_ready + 2;            // [1] leader: both CTAs' bias operands of the current layer landed (TMA)
  uint64_t *bias_empty = bias_full + 1;          // [1] both bias MMAs of the layer have read the operand (commit, multicast)
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bias_empty + 1);
  // The bias enters the accumulator through the tensor core: one extra K = 16 MMA per tile and layer multiplies a
  // constant "ones" operand (rows of [1,1,1,0,...]) by the layer's bias operand (row n = [hi, mid, lo, 0, ...], the
  // fp32 bias split into three bf16 terms, exact to ~2^-24).  Both are un-swizzled K-major core matrices; every 8-row
  // group of the ones operand aliases the same 128 bytes (SBO = 0) and the bias operand's second K core matrix
  // aliases its first (LBO = 0: it only meets the zero half of the ones rows).  This takes 16 shared-memory loads and
  // 32 adds per thread and tile out of the epilogue, which is the critical path (scripts/ubench/mma_2cta.cu).
  uint8_t *s_ones = reinterpret_cast<uint8_t *>(bars + 32);          // 256 B
  uint8_t *s_bias = s_ones + 256;                                    // [N/2 <= 128 rows x 16 B] of this CTA

  // warp index through a shuffle: the compiler then treats it (and every branch on it) as warp-uniform and
  // keeps the MMA / TMA operands in uniform registers instead of R2UR "waterfall" loops per instruction
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  unsigned trace_n = 0;
  (void)trace_n;
  const int L = a.n_layers;
  const long long n_tiles = (a.P + 127) / 128;
  const long long n_quads = (n_tiles + 3) / 4;          // a CTA pair works on 4 tiles: tile = 4*quad + 2*rank + t
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { printf("nfs_b200: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(in_full + t, 1); mbar_init(act_free + t, 1); mbar_init(act_ready + t, 32);
      mbar_init(acc_full + t, 1); mbar_init(head_done + t, 16); mbar_init(acc_free + t, 32);
      mbar_init(in_ready + t, 8);
    }
    mbar_init(bias_full, 1); mbar_init(bias_empty, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    if (a.save) tma_prefetch_desc(&tmap_save);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 80) {      // the "ones" operand: core matrix 0 = rows of [1,1,1,0,...], core matrix 1 = 0
    const int i = threadIdx.x - 64;
    const uint32_t one2 = 0x3F803F80u, one1 = 0x00003F80u;           // bf16 {1,1}, {1,0}
    *reinterpret_cast<uint4 *>(s_ones + i * 16) = i < 8 ? make_uint4(one2, one1, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (warp == 1) {                                 // the pair allocates together: warp 1 of both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();                                   // (nfs_common.cuh: programmatic dependent launch)
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t wit = 0, iter = 0;
      const int ks0 = a.K[0] >> 6;
      for (long long quad = quad0; quad < n_quads; quad += quad_step, ++iter) {
        if (quad_consumed != nullptr && quad >= window) {
          // back-pressure of the merged backward kernel (performance only, never correctness): stay at most `window`
          // quads ahead of the weight-gradient consumers so that what this pair stores is still in L2 when they read
          // it.  Gives up after ~2 ms: producers must never depend on consumers for progress.
          const unsigned int *c = quad_consumed + (quad - window);
          for (uint32_t spin = 0; spin < (1u << 12); ++spin) {
            unsigned v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if (v >= consumed_target) break;
            __nanosleep(100);
          }
        }
        for (int t = 0; t < 2 && a.points == nullptr; ++t) {
          mbar_wait_relaxed(act_free + t, (iter & 1) ^ 1);
          mbar_wait_relaxed(head_done + t, (iter & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(in_full + t, (uint32_t)(2 * ks0 * kActSlab));     // both CTAs' tiles
          for (int s = 0; s < ks0; ++s)
            tma_load_2d_pair(act[t] + s * kActSlab, &tmap_x, in_full + t, s * 64, (int)((4 * quad + 2 * rank + t) * 128));
        }
        for (int l = 0; l < L; ++l) {
          const int ks = a.K[l] >> 6, nh = a.N[l] >> 1;           // this CTA's half of the output rows
          const int nb = (nh + 63) >> 6;                          // 64-row TMA boxes (a 32-row half loads a full box)
          for (int s = 0; s < ks; ++s, ++wit) {
            const uint32_t stage = wit % kWStages, ph = (wit / kWStages) & 1;
            if ((NFS_DBG(a) & 2) && wit >= kWStages) continue;
            mbar_wait_relaxed(w_empty + stage, ph ^ 1);
            if (rank == 0) mbar_expect_tx(w_full + stage, (uint32_t)(2 * nb * 8192));
            for (int b = 0; b < nb; ++b)
              tma_load_2d_pair(wring + stage * kWStage + b * 8192, &tmap_w, w_full + stage, s * 64,
                               a.row0[l] + (int)rank * nh + b * 64);
            // The bias operand (single buffer) is released when the previous layer's B pass STARTS (its bias MMA
            // is issued first), about when that layer's first weight stages come free: loading it here, between
            // the weight stages, keeps the weight prefetch running ahead.
            if (s == (ks > 1 ? 1 : 0) && a.has_bias && !(NFS_DBG(a) & 8)) {
              const uint32_t gl = iter * (uint32_t)L + (uint32_t)l;
              mbar_wait_relaxed(bias_empty, (gl & 1) ^ 1);
              if (rank == 0) mbar_expect_tx(bias_full, 2u * 2048u);
              tma_load_2d_pair(s_bias, &tmap_b, bias_full, 0, a.row0[l] + (int)rank * nh);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {                 // all 32 lanes walk the (warp-uniform) schedule; one elected lane issues
      uint32_t wit = 0, iter = 0, n_ready[2] = {0, 0};
      const uint64_t b_desc0 = umma_desc_sw128(smem_u32(wring), 16, 1024);
      const bool no_mma = (NFS_DBG(a) & 4) != 0;
      const bool use_bias = a.has_bias && !(NFS_DBG(a) & 8);
      const uint64_t ones_desc = umma_desc_noswizzle(smem_u32(s_ones), 128, 0);
      const uint64_t bias_desc = umma_desc_noswizzle(smem_u32(s_bias), 0, 128);
      // the barriers a tile step (layer l, tile t) needs before its first MMA.  (Acquiring them one slab EARLY, while the
      // previous step's last MMAs are still queued, was tried and is slower - 0.61 vs 0.56 ms for the inference chain: the
      // tile's epilogue + hand-off completes just about when the other tile's MMAs have been issued, so an early wait
      // holds that last slab back.  scripts/trace_fused.py, gpurun_out/r2_ab_lookahead.log)
      auto wait_input = [&](int l, int t) {
        if (l == 0) {
          mbar_wait_cluster(acc_free + t, (iter & 1) ^ 1);   // accumulator t drained by the previous quad's last layer
          mbar_wait_cluster((a.points != nullptr ? in_ready : in_full) + t, iter & 1);
        } else {
          mbar_wait_cluster(act_ready + t, n_ready[t] & 1);
          ++n_ready[t];
        }
      };
      for (long long quad = quad0; quad < n_quads; quad += quad_step, ++iter) {
        // Issue order per layer: tile major - 4*ks MMAs on the pair's A tiles, then 4*ks on the B tiles.
        // Long runs on one accumulator (switching the D operand between consecutive MMAs costs
        // ~150-270 cycles, scripts/ubench/mma_modes.cu); the layer's weights (ks stages of [N/2 x 64]
        // per CTA) stay resident for both passes.
        for (int l = 0; l < L; ++l) {
          const int ks = a.K[l] >> 6;
          const uint32_t idesc = umma_idesc_bf16(256, a.N[l], 0, 0);
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            wait_input(l, t);
            NFS_TRACE(14, l, t);
            NFS_TRACE(15, l, t);
            if (!(NFS_DBG(a) & 128)) tc_fence_after();
            NFS_TRACE(1, l, t);
            const uint32_t d_tmem = tmem_base + (uint32_t)(t * 256);
            // descriptors differ only in their 14-bit start-address field: +2 per 32-byte K step
            const uint64_t a_desc0 = umma_desc_sw128(smem_u32(act[0]) + (uint32_t)t * kActBytes, 16, 1024);
            if (t == 1 && use_bias) {           // tile B: bias first, which frees the operand for the next layer
              if (elect_one()) {
                umma_bf16_pair(d_tmem, ones_desc, bias_desc, idesc, 0u);
                umma_commit_pair(bias_empty);
              }
              __syncwarp();
            }
            for (int s = 0; s < ks; ++s) {
              const uint32_t w = wit + s, stage = w % kWStages, wph = (w / kWStages) & 1;
              if (t == 0) {
                if (!((NFS_DBG(a) & 2) && w >= kWStages)) mbar_wait_cluster(w_full + stage, wph);
                tc_fence_after();
                NFS_TRACE(10 + s, l, t);
              }
              const uint64_t ad = a_desc0 + (uint64_t)((s * kActSlab) >> 4);
              const uint64_t bd = b_desc0 + (uint64_t)((stage * kWStage) >> 4);
              if (elect_one()) {
                if (!no_mma) {
                  umma_bf16_pair(d_tmem, ad, bd, idesc, (uint32_t)(s != 0 || (t == 1 && use_bias)));
                  umma_bf16_pair(d_tmem, ad + 2, bd + 2, idesc, 1u);
                  umma_bf16_pair(d_tmem, ad + 4, bd + 4, idesc, 1u);
                  umma_bf16_pair(d_tmem, ad + 6, bd + 6, idesc, 1u);
                }
                if (t == 1) umma_commit_pair(w_empty + stage);
              }
              __syncwarp();
              NFS_TRACE(6 + s, l, t);
            }
            if (t == 0 && use_bias) {           // tile A: bias last (its operand has had the whole pass to land)
              mbar_wait_cluster(bias_full, (iter * (uint32_t)L + (uint32_t)l) & 1);
              tc_fence_after();
              if (elect_one()) umma_bf16_pair(d_tmem, ones_desc, bias_desc, idesc, 1u);
              __syncwarp();
            }
            NFS_TRACE(2, l, t);
            if (elect_one()) {
              umma_commit_pair(acc_full + t);
              if (l == L - 1) umma_commit_pair(act_free + t);
            }
            __syncwarp();
            NFS_TRACE(0, l, t);
          }
          wit += ks;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    // Each warp drains one (TMEM lane quadrant q) x (quarter of the columns) block of the accumulator
    // with ONE tcgen05.ld (x64 for 256-wide layers), first for tile A then for tile B.
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int r_in = q * 32 + lane;
    uint32_t n_full[2] = {0, 0}, gl = 0;
    const uint64_t store_policy = quad_done != nullptr ? l2_policy_evict_last() : 0ull;
#ifdef NFS_DEVTOOLS
    uint32_t n_arr[2] = {0, 0};          // tracer probe: phases of act_ready this warp has arrived on
#endif
    uint2 b_n1 = make_uint2(0u, 0u), b_n2 = make_uint2(0u, 0u);   // sign bits of the next two epilogue steps
    bool store_pending = false;
    const uint32_t ready_bar[2] = {map_to_cta(act_ready, 0), map_to_cta(act_ready + 1, 0)};   // in the leader CTA
    const uint32_t free_bar[2] = {map_to_cta(acc_free, 0), map_to_cta(acc_free + 1, 0)};
    const uint32_t in_bar[2] = {map_to_cta(in_ready, 0), map_to_cta(in_ready + 1, 0)};
    uint32_t iter = 0;
    for (long long quad = quad0; quad < n_quads; quad += quad_step, ++iter) {
      if (!kMasked && a.points != nullptr && cq < 2) {
        // K2 fused in: the warps with cq == t write the positional encoding of tile t (thread = point) straight
        // into slab 0 of act[t] - [x | sin(x f_0) | cos(x f_0) | sin(x f_1) | ...], one accurate sincosf per
        // coordinate and the double-angle recurrence for the higher octaves (identical arithmetic to
        // posenc_bf16_kernel's octave path, so the result is bit-identical to the two-kernel route)
        const int t = cq;
        mbar_wait_relaxed(act_free + t, (iter & 1) ^ 1);       // the previous quad's last MMAs have read act[t]
        mbar_wait_relaxed(head_done + t, (iter & 1) ^ 1);      // and every epilogue warp is done with it
        const long long row = (4 * quad + 2 * rank + t) * 128 + r_in;
        float x[3] = {0.f, 0.f, 0.f}, sn[3], cs[3];
        if (row < a.P) {
          if (a.ray_z != nullptr) {             // sampler fused in: the (P,3) positions never exist in HBM
            const long long ray = row / a.ray_S;
            const float zz = __ldg(a.ray_z + row);
#pragma unroll
            for (int d = 0; d < 3; ++d)
              x[d] = __fadd_rn(__ldg(a.rays_o + ray * 3 + d), __fmul_rn(__ldg(a.rays_d + ray * 3 + d), zz));   // ray_utils.py:82
          } else {
            x[0] = __ldg(a.points + row * 3); x[1] = __ldg(a.points + row * 3 + 1); x[2] = __ldg(a.points + row * 3 + 2);
          }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) sincosf(__fmul_rn(x[d], a.freq0), &sn[d], &cs[d]);
        uint8_t *srow = smem + t * kActBytes + r_in * 128;
        float buf[8];
        int nb = 0, chunk = 0;
        auto push = [&](float v) {
          buf[nb++] = v;
          if (nb == 8) {
            *reinterpret_cast<uint4 *>(srow + ((chunk ^ (r_in & 7)) << 4)) =
                make_uint4(pack_bf16x2(buf[0], buf[1]), pack_bf16x2(buf[2], buf[3]), pack_bf16x2(buf[4], buf[5]),
                           pack_bf16x2(buf[6], buf[7]));
            nb = 0; ++chunk;
          }
        };
        push(x[0]); push(x[1]); push(x[2]);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
          if (k < a.n_octaves) {
            push(sn[0]); push(sn[1]); push(sn[2]); push(cs[0]); push(cs[1]); push(cs[2]);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float s2 = 2.f * sn[d] * cs[d], c2 = 1.f - 2.f * sn[d] * sn[d];
              sn[d] = s2; cs[d] = c2;
            }
          } else {
            push(0.f); push(0.f); push(0.f); push(0.f); push(0.f); push(0.f);
          }
        }
        push(0.f);                                            // column 63: zero padding
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (kTrain && a.x_save) {
          // training: the first layer's weight gradient needs this operand - store the warp's 32 rows (4 KB) and
          // wait until the TMA has read them (another warp overwrites them in layer 0's epilogue)
          if (lane == 0) {
            tma_store_2d(&tmap_x, smem + t * kActBytes + q * 32 * 128, 0, (int)((4 * quad + 2 * rank + t) * 128 + q * 32));
            bulk_commit();
            bulk_wait_read0();
          }
          __syncwarp();
        }
        if (lane == 0) mbar_arrive_cluster(in_bar[t]);
      }
      for (int l = 0; l < L; ++l, ++gl) {
        const bool last = (l == L - 1);
        const bool is_head = last && a.head != 0;
        const int Nl = a.N[l], quarter = Nl >> 2, c0 = cq * quarter;
        const int act_l = a.act[l];
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const long long tile = 4 * quad + 2 * rank + t;
          const long long row = tile * 128 + r_in;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);
          uint8_t *act_t = smem + t * kActBytes;
          // ReLU-backward mask row (straight from HBM): issue the loads before blocking on the accumulator
          // ReLU-backward sign bits of this thread's `quarter` columns (8 or 4 bytes per tile step): loaded two epilogue
          // steps ahead (same tile, previous layer's step), so their latency never shows
          const bool use_bits = kMasked && act_l == 4 && tile < n_tiles;
          const int words = quarter >> 5;                       // 2 (256-wide layers) or 1 (128-wide)
          auto load_bits = [&](int layer) -> uint2 {           // widths may differ from layer to layer: the words this
            const int lq = a.N[layer] >> 2;                    // thread owns follow the width of the layer they mask
            const uint32_t *bp = a.bits_in + ((long long)a.mask_idx[layer] * a.bits_rows + row) * 8 + ((cq * lq) >> 5);
            uint2 v = make_uint2(0u, 0u);
            if (lq == 64) v = __ldg(reinterpret_cast<const uint2 *>(bp));
            else v.x = __ldg(bp);
            return v;
          };
          uint2 b_cur = make_uint2(0u, 0u);
          if (kMasked) {
            if (use_bits) b_cur = (l == 0) ? load_bits(0) : b_n1;
            b_n1 = b_n2;
            if (l + 1 < L && a.act[l + 1] == 4 && tile < n_tiles) b_n2 = load_bits(l + 1);
          }
          uint32_t bits_acc[2] = {0u, 0u};
          NFS_TRACE(15, l, t);
          mbar_wait_relaxed(acc_full + t, n_full[t] & 1);
          ++n_full[t];
          tc_fence_after();
          NFS_TRACE(3, l, t);
          if (!is_head) {
            // the TMA store this warp issued from act[t] one layer ago has finished READING the block that is
            // overwritten below (the store issued for the other tile a moment ago may still be in flight)
            // 128-wide layers: two warps (cq, cq^1) share one 64-column TMA-store box -> pair barriers
            const bool do_save = (kTrain || kMasked) && a.save != 0;
            const bool paired = do_save && quarter == 32;
            const uint32_t pair_bar = 1u + (uint32_t)(q * 2 + (cq >> 1));
            // Width change (256 -> 128 or back): the columns this warp writes now were stored by ANOTHER warp of the
            // same 32 rows one layer ago - every warp waits for its own store, then the four warps of the row quarter
            // meet (without this, one tile in a few hundred kept stale columns: scripts/dev/chain_mixed.py).
            const bool changed = l > 0 && a.N[l - 1] != Nl;
            if (store_pending) {
              if (lane == 0 && (changed || !paired || (cq & 1) == 0)) bulk_wait_read1();
              __syncwarp();
              if (changed) asm volatile("bar.sync %0, 128;" ::"r"(9u + (uint32_t)q) : "memory");
              else if (paired) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            }
            NFS_TRACE(14, l, t);
            uint8_t *srow = act_t + (c0 >> 6) * kActSlab + r_in * 128;
            const int ch0 = (c0 & 63) >> 3;
            if (NFS_DBG(a) & 1) {
              // bisection: no TMEM drain, no math, no smem writes
            } else {
              // `quarter` (64 or 32) columns in chunks of 16 (small chunks leave registers for a whole chunk of
              // bias values to be in flight at once; the other three warps of the scheduler hide the latencies).
              // Software-pipelining the chunks over two register buffers (chunk c + 1 in flight from TMEM while chunk c
              // is converted) was measured on one box against this form: inference chain 0.495 vs 0.485 ms, dgrad chain
              // 0.89 vs 0.77 ms - slower (96 registers), gpurun_out/r2_ab_libs1.log.
              const int n_chunks = quarter >> 4;
              uint32_t va[16];
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < n_chunks) {
                  tmem_ld16_async(taddr + c0 + 16 * c, va);
                  tmem_ld_wait();
                  epi_chunk<kMasked, kTrain>(va, act_l, (c >> 1) ? b_cur.y : b_cur.x, 8 * (c & 1),
                                     bits_acc[c >> 1], srow, ch0 + 2 * c, r_in & 7, NFS_DBG(a));
                }
              }
            }
            if (kTrain && a.bits_out != nullptr && act_l == 1 && tile < n_tiles) {
              uint32_t *bp = a.bits_out + ((long long)l * a.save_rows + row) * 8 + (c0 >> 5);
              if (words == 2) *reinterpret_cast<uint2 *>(bp) = make_uint2(bits_acc[0], bits_acc[1]);
              else *bp = bits_acc[0];
            }
            NFS_TRACE(4, l, t);
            tc_fence_before();
            if (!(NFS_DBG(a) & 16)) fence_proxy_async();   // generic-proxy smem writes -> visible to UMMA / TMA
            __syncwarp();
            NFS_TRACE(5, l, t);
            if (paired) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            if (lane == 0) {
              if (!last) mbar_arrive_cluster(ready_bar[t]);
#ifdef NFS_DEVTOOLS
              if (!last) {
                // tracer probe (leader CTA, one warp): when does the act_ready phase this warp just arrived on complete?
                if (a.trace != nullptr && blockIdx.x == 0 && warp == 2) {
                  for (int spin = 0; spin < 100000 && !mbar_try_wait(act_ready + t, n_arr[t] & 1); ++spin) { }
                  NFS_TRACE(13, l, t);
                }
                ++n_arr[t];
              }
#endif
              if (do_save && tile < n_tiles && (!paired || (cq & 1) == 0)) {
                // merged backward kernel: the rows are read back by the weight-gradient consumers within a few tens of
                // microseconds - ask L2 to keep them (a plain bulk store is not retained: the consumers then read
                // 8.9 GB per step from DRAM instead of 6.4, gpurun_out/r2c_bwd_dram*.csv, r2_handoff_matrix.log)
#ifndef NFS_NO_STORE_HINT
                if (quad_done != nullptr)
                  tma_store_2d_hint(&tmap_save, act_t + (c0 >> 6) * kActSlab + q * 32 * 128, c0 & ~63,
                                    (int)(l * a.save_rows + tile * 128 + q * 32), store_policy);
                else
#endif
                  tma_store_2d(&tmap_save, act_t + (c0 >> 6) * kActSlab + q * 32 * 128, c0 & ~63,
                               (int)(l * a.save_rows + tile * 128 + q * 32));
                bulk_commit();
              }
            }
            store_pending = do_save;
            if (last) {                          // chain ends in a regular layer: tile t is finished once the
              if (lane == 0 && store_pending) bulk_wait_read0();   // stores have read act[t] (it is reloaded next)
              __syncwarp();
              if (lane == 0) { mbar_arrive(head_done + t); mbar_arrive_cluster(free_bar[t]); }
              if (quad_done != nullptr && t == 1 && lane == 0) {
                // hand-off: every store this warp issued for the quad is complete (not just read) -> publish.
                // The pipeline has already been released above, so this wait overlaps the next quad's first MMAs.
                bulk_wait0();
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                atomicAdd(quad_done + quad, 1u);
              }
            }
          } else {
            // head: first out_cols (<= 16) columns, fp32, reference output activation; quarter 0 only
            if (cq == 0) {
              float v[16];
              tmem_ld16(taddr, v);
              if (row < a.P) {
                if (act_l == 5) {                  // 2-way softmax gate (dino_feature_model.py:169,188), as nfs_linear_bf16
                  const float mx = fmaxf(v[0], v[1]);
                  const float e0 = expf(v[0] - mx), e1 = expf(v[1] - mx);
                  const float inv = 1.0f / (e0 + e1);
                  v[0] = e0 * inv; v[1] = e1 * inv;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (j < a.out_cols) {
                    float x = v[j];
                    if (act_l == 1) x = fmaxf(x, 0.f);
                    else if (act_l == 3 || (act_l == 2 && j < 3)) x = fm_sigmoid(x);
                    v[j] = x;
                  }
                }
                if (a.out_cols == 4) {
                  *reinterpret_cast<float4 *>(a.out + row * 4) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (j < a.out_cols) a.out[row * a.out_cols + j] = v[j];
                }
              }
            }
            tc_fence_before();
            if (store_pending && lane == 0) bulk_wait_read0();
            __syncwarp();
            if (lane == 0) { mbar_arrive(head_done + t); mbar_arrive_cluster(free_bar[t]); }
          }
        }
        if (last) store_pending = false;
      }
    }
    if (lane == 0) bulk_wait0();               // all saved activations are in global memory
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be signalling this CTA's barriers / reading its operands
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace
}  // namespace nfs

// ---------------------------------------------------------------- host: argument checks + tensor maps of a chain launch
// returns 0, or 1 for an empty launch, or an error (< 0 / cudaError)
// Boxes of the stacked weight tensor and of the saved activations differ from make_tmap_bf16's
// default only in their row count (64 resp. 32).
// points != NULL: in-kernel encoding; x_bf16 is then NULL (inference) or the [rows128, 64] bf16 buffer that RECEIVES the
// encoded operand (training forward).
static inline int chain_prepare(const char *fn, const void *x_bf16, const float *points, float freq0, int n_octaves,
                        int64_t n_points, int32_t n_layers, const int32_t *k_dims,
                        const int32_t *n_dims, const int32_t *acts, const int32_t *row0,
                        const void *w_stack_bf16, int32_t w_rows, const void *bias_terms_bf16,
                        const void *relu_bits_in, int64_t bits_rows_per_layer, const int32_t *mask_idx,
                        void *save_bf16, void *relu_bits_out, int64_t save_rows_per_layer, float *out_f32,
                        int32_t out_cols, nfs::FusedArgs *a_out, CUtensorMap *tx_out, CUtensorMap *tw_out,
                        CUtensorMap *ts_out, CUtensorMap *tb_out) {
  using namespace nfs;
  if (n_points < 0 || n_layers < 2 || n_layers > kFmMaxLayers) return fail_arg(fn, NFS_E_BADARG, "need 2..12 layers");
  if (n_points == 0) return 1;          // nothing to do
  if ((!x_bf16 && !points) || !k_dims || !n_dims || !acts || !row0 || !w_stack_bf16 || (!out_f32 && !save_bf16))
    return fail_arg(fn, NFS_E_BADARG, "null pointer");
  FusedArgs a{};
  a.P = n_points; a.n_layers = n_layers; a.has_bias = bias_terms_bf16 != nullptr; a.out = out_f32; a.out_cols = out_cols;
  a.save = save_bf16 != nullptr; a.save_rows = save_rows_per_layer;
  a.head = out_f32 != nullptr;
  a.points = points; a.freq0 = freq0; a.n_octaves = n_octaves;
  a.dbg = 0;
  a.trace = nullptr;
  a.bits_in = (const uint32_t *)relu_bits_in; a.bits_rows = bits_rows_per_layer;
  a.bits_out = (uint32_t *)relu_bits_out;
  for (int l = 0; l < n_layers; ++l) {
    a.K[l] = k_dims[l]; a.N[l] = n_dims[l]; a.act[l] = acts[l]; a.row0[l] = row0[l];
    a.mask_idx[l] = mask_idx ? mask_idx[l] : 0;
    if (a.act[l] == 4 && (!relu_bits_in || !mask_idx || a.mask_idx[l] < 0))
      return fail_arg(fn, NFS_E_BADARG, "act 4 (ReLU backward) needs relu_bits_in and mask_idx");
    if ((a.act[l] == 4 || (a.act[l] == 1 && relu_bits_out)) && a.N[l] < 128 && !(out_f32 && l == n_layers - 1))
      return fail_arg(fn, NFS_E_UNSUPPORTED, "ReLU sign bits need layers at least 128 wide");
    if (a.act[l] < 0 || a.act[l] > 5) return fail_arg(fn, NFS_E_BADARG, "act must be 0..5");
    if (a.act[l] == 5 && !(out_f32 && l == n_layers - 1 && out_cols == 2))
      return fail_arg(fn, NFS_E_BADARG, "act 5 (2-way softmax) belongs to an output head with out_cols == 2");
    if (a.K[l] % 64 || a.K[l] <= 0 || a.K[l] > 256 || a.N[l] % 64 || a.N[l] <= 0 || a.N[l] > 256 ||
        a.row0[l] < 0 || a.row0[l] + a.N[l] > w_rows)
      return fail_arg(fn, NFS_E_UNSUPPORTED, "layer dims must be multiples of 64 in [64,256] and fit the weight stack");
    if (l > 0 && a.K[l] != a.N[l - 1]) return fail_arg(fn, NFS_E_BADARG, "layer l input width != layer l-1 output width");
  }
  if (a.head && (out_cols <= 0 || out_cols > a.N[n_layers - 1])) return fail_arg(fn, NFS_E_BADARG, "out_cols out of range");
  const long long rows128 = ((n_points + 127) / 128) * 128;
  if (a.save && save_rows_per_layer < rows128)
    return fail_arg(fn, NFS_E_BADARG, "save_rows_per_layer must be >= n_points rounded up to 128");
  if (relu_bits_in && bits_rows_per_layer < rows128)
    return fail_arg(fn, NFS_E_BADARG, "bits_rows_per_layer must be >= n_points rounded up to 128");
  if (relu_bits_out && (relu_bits_in || save_rows_per_layer < rows128))
    return fail_arg(fn, NFS_E_BADARG, "relu_bits_out belongs to a forward chain (no relu_bits_in) with save_rows_per_layer >= rows");
  const int n_saved = a.head ? n_layers - 1 : n_layers;
  // saved rows are as wide as the widest saved layer; a narrower layer fills the first N_l columns of its rows
  int save_w = 0;
  if (a.save)
    for (int l = 0; l < n_saved; ++l) {
      if (a.N[l] < 128) return fail_arg(fn, NFS_E_UNSUPPORTED, "saved activations need layers at least 128 wide");
      save_w = a.N[l] > save_w ? a.N[l] : save_w;
    }

  CUtensorMap tx{}, tw{}, ts{}, tb{};
  int rc = 0;
  if (points != nullptr) {
    if (a.K[0] != 64 || n_octaves < 1 || n_octaves > 10)
      return fail_arg(fn, NFS_E_UNSUPPORTED, "in-kernel encoding needs a 64-wide first layer and 1..10 octaves");
    if (x_bf16 != nullptr) {
      if (!a.save) return fail_arg(fn, NFS_E_BADARG, "the encoded operand is only stored by a training forward (save_bf16)");
      a.x_save = 1;
      rc = tc::make_tmap_bf16(&tx, x_bf16, (uint64_t)rows128, 64, 64, 32, fn);
      if (rc) return rc;
    }
  } else {
    rc = tc::make_tmap_bf16(&tx, x_bf16, (uint64_t)n_points, (uint64_t)a.K[0], (uint64_t)a.K[0], 128, fn);
    if (rc) return rc;
  }
  rc = tc::make_tmap_bf16(&tw, w_stack_bf16, (uint64_t)w_rows, 256, 256, 64, fn);
  if (rc) return rc;
  if (bias_terms_bf16 != nullptr) {
    rc = tc::make_tmap_rows16(&tb, bias_terms_bf16, (uint64_t)w_rows, 128, fn);
    if (rc) return rc;
  }
  if (a.save) {
    rc = tc::make_tmap_bf16(&ts, save_bf16, (uint64_t)(save_rows_per_layer * n_saved), (uint64_t)save_w,
                            (uint64_t)save_w, 32, fn);
    if (rc) return rc;
  }
  *a_out = a; *tx_out = tx; *tw_out = tw; *ts_out = ts; *tb_out = tb;
  return 0;
}

// dynamic shared memory of the chain body: tiles, weight ring, barriers, ones, bias operand
static constexpr size_t kChainSmemBytes = 2 * nfs::kActBytes + nfs::kWStages * nfs::kWStage + 256 + 256 + 128 * 16;
