// Microbenchmark + semantics check: tcgen05.mma.cta_group::2 (CTA pair, M = 256 = 2 x 128 rows,
// N = 256 split as 128 B-operand rows per CTA), bf16 -> fp32, operands in shared memory.
// Verifies both CTAs' accumulators against a host reference and times the MMA stream.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ inline float a_val(int r, int k) { return (float)((r * 7 + k * 3) % 13 - 6); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k) % 11 - 5); }
__host__ __device__ inline float bias_val(int n) { return 0.1234f * (float)n - 7.7f + 1e-3f * (float)(n % 7); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;                       // layout_type 0 = no swizzle
}

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(576, 1)
mma2_bench(int N, int iters, float *d_out, unsigned long long *cycles, int rotate) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sa = smem, *sb = smem + 131072;          // A: [128 x 64], B half: [N/2 x 64]
  uint64_t *bar = (uint64_t *)(smem + 131072 + 81920);
  uint32_t *slot = (uint32_t *)(bar + 4);
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_rank();
  const int nh = N / 2;
  for (int i = tid; i < 128 * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(a_val(r + 128 * (int)rank, c * 8 + j));
    *(uint4 *)(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *(uint4 *)v;
  }
  for (int i = tid; i < nh * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(b_val(r + nh * (int)rank, c * 8 + j));
    *(uint4 *)(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *(uint4 *)v;
  }
  // bias operands (no swizzle, K-major core matrices of 8 rows x 16 bytes):
  //   ones: core matrix 0 = rows of [1,1,1,0,0,0,0,0], core matrix 1 (k = 8..15) = zeros; every 8-row group aliases it (SBO = 0)
  //   bias: row n = [hi, mid, lo, 0...] (three bf16 terms of the fp32 bias), groups of 8 rows 128 bytes apart
  uint8_t *s_ones = smem + 131072 + 81920 + 64, *s_bias = s_ones + 256;
  if (tid < 16) {
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((tid < 8 && j < 3) ? 1.f : 0.f);
    *(uint4 *)(s_ones + tid * 16) = *(uint4 *)v;
  }
  for (int i = tid; i < nh; i += blockDim.x) {
    const float b = bias_val(i + nh * (int)rank);
    const __nv_bfloat16 hi = __float2bfloat16(b);
    const float r1 = b - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16(r1);
    const __nv_bfloat16 lo = __float2bfloat16(r1 - __bfloat162float(mid));
    __nv_bfloat16 v[8] = {hi, mid, lo, __float2bfloat16(0.f), __float2bfloat16(0.f), __float2bfloat16(0.f), __float2bfloat16(0.f), __float2bfloat16(0.f)};
    *(uint4 *)(s_bias + i * 16) = *(uint4 *)v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 2)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 3)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  unsigned long long t0 = 0;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = idesc_bf16(256, N);
    t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int k = 0; k < 4; ++k) {
        const uint32_t acc = (uint32_t)((it | k) != 0);
        if (rotate == 3 && k == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (rotate == 4 && k == 0) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(smem_u32(bar + 3)), "r"(1) : "memory");
          if (!ok) __trap();
        }
        const uint32_t a_off = rotate == 1 ? (uint32_t)(it % 8) * 16384u : 0u;
        const uint32_t b_off = rotate == 1 ? (uint32_t)(it % 5) * 16384u : 0u;
        const uint32_t d_off = rotate == 1 ? (uint32_t)((it / 4) & 1) * 256u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tbase + d_off), "l"(desc_sw128(smem_u32(smem) + a_off + k * 32)), "l"(desc_sw128(smem_u32(smem) + 131072u + b_off + k * 32)), "r"(idesc), "r"(acc)
                     : "memory");
        if (rotate == 2 && k == 3)
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(smem_u32(bar + 2)), "h"((uint16_t)3) : "memory");
      }
    if (rotate == 7)      // D += ones . bias^T
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tbase), "l"(desc_noswz(smem_u32(s_ones), 128, 0)), "l"(desc_noswz(smem_u32(s_bias), 0, 128)), "r"(idesc), "r"(1u)
                   : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
  }
  if (rotate >= 5 && rotate < 7 && warp >= 4) {   // pollers: spin on the local barrier like the chain kernel's epilogue warps
    uint32_t ok = 0; unsigned spin = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
      if (rotate == 6 && !ok) __nanosleep(40);
      if (++spin > (1u << 26)) { __trap(); }
    }
  }
  if (tid == 0) {        // both CTAs: wait for the multicast commit on the local barrier
    uint32_t ok = 0; unsigned spin = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
      if (++spin > (1u << 24)) { printf("timeout rank %u block %d\n", rank, (int)blockIdx.x); __trap(); }
    }
    if (rank == 0) cycles[blockIdx.x >> 1] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (d_out != nullptr && blockIdx.x < 2 && warp < 4) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(lane_addr + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) d_out[(rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
  float *d_out; unsigned long long *cyc;
  cudaMalloc(&d_out, 256 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int smem = 1024 + 131072 + 81920 + 64 + 256 + 2048 + 64;
  cudaFuncSetAttribute(mma2_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float *h = (float *)malloc(256 * 256 * 4);
  for (int N : {256, 128, 64}) {
    cudaMemset(d_out, 0, 256 * 256 * 4);
    mma2_bench<<<2, 576, smem>>>(N, 1, d_out, cyc, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: error %s\n", N, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d_out, 256 * 256 * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int r = 0; r < 256; ++r)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 64; ++k) ref += a_val(r, k) * b_val(n, k);
        double err = fabs((double)h[r * 256 + n] - ref);
        if (err > maxerr) maxerr = err;
        if (err > 1e-3 && bad++ < 4) printf("   mismatch r %d n %d got %f ref %f\n", r, n, h[r * 256 + n], ref);
      }
    {   // bias MMA check
      cudaMemset(d_out, 0, 256 * 256 * 4);
      mma2_bench<<<2, 576, smem>>>(N, 1, d_out, cyc, 7);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (e2 != cudaSuccess) { printf("bias MMA N %d: error %s\n", N, cudaGetErrorString(e2)); return 1; }
      cudaMemcpy(h, d_out, 256 * 256 * 4, cudaMemcpyDeviceToHost);
      double me = 0; int nbad = 0;
      for (int r = 0; r < 256; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += a_val(r, k) * b_val(n, k);
          ref += bias_val(n);
          double err = fabs((double)h[r * 256 + n] - ref);
          if (err > me) me = err;
          if (err > 1e-3 && nbad++ < 4) printf("   bias mismatch r %d n %d got %f ref %f\n", r, n, h[r * 256 + n], ref);
        }
      printf("2-CTA N=%3d + bias MMA (ones x [hi mid lo]): max|D-ref| = %g (%d bad)\n", N, me, nbad);
    }
    const int iters = 4000;
  for (int rotate : {0, 5, 6}) {
    mma2_bench<<<148, 576, smem>>>(N, iters, nullptr, cyc, rotate);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N %d: error %s\n", N, cudaGetErrorString(e)); return 1; }
    unsigned long long hc[74];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int i = 0; i < 74; ++i) if (hc[i] > mx) mx = hc[i];
    printf("2-CTA SS N=%3d rotate=%d: max|D-ref| = %g (%d bad); %.1f cycles per MMA (M256 x N x K16), %.0f flop/cyc/SM\n", N, rotate, maxerr, bad,
           (double)mx / (iters * 4.0), 2.0 * 256 * N * 16 * iters * 4.0 / (double)mx / 2.0);
  }
  }
  return 0;
}
