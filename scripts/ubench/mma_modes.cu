// Microbenchmark + layout check: tcgen05.mma kind::f16 (bf16 -> fp32), M = 128, with the A operand
// in shared memory (SS) or in tensor memory (TS), N = 256 / 128.  Verifies D against a host
// reference (exact small integers) and reports cycles per MMA instruction on all SMs at once.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
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

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: SS, 1: TS.  K = 64 per "step" (4 MMAs of K = 16).
__global__ void __launch_bounds__(128, 1) mma_bench(int mode, int N, int iters, float *d_out, unsigned long long *cycles, int n_acc) {
  extern __shared__ uint8_t raw[];
  uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t *sa = smem, *sb = smem + 16384;
  uint64_t *bar = (uint64_t *)(sb + 32768);
  uint32_t *slot = (uint32_t *)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // A [128 x 64], B [256 x 64], K-major, 128-byte rows, 16-byte chunks XOR-swizzled with (row & 7)
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(a_val(r, c * 8 + j));
    *(uint4 *)(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *(uint4 *)v;
  }
  for (int i = tid; i < 256 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(b_val(r, c * 8 + j));
    *(uint4 *)(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *(uint4 *)v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *slot;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_tmem = tbase + 256;            // A operand: 128 lanes x 32 columns (64 bf16 per row)
  if (mode == 1) {
    // thread = row; 32-bit column j holds K elements (2j, 2j+1)
    uint32_t p[32];
    for (int j = 0; j < 32; ++j) {
      __nv_bfloat162 t = __floats2bfloat162_rn(a_val(tid, 2 * j), a_val(tid, 2 * j + 1));
      p[j] = *(uint32_t *)&t;
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(lane_addr + 256), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]),
          "r"(p[8]), "r"(p[9]), "r"(p[10]), "r"(p[11]), "r"(p[12]), "r"(p[13]), "r"(p[14]), "r"(p[15]), "r"(p[16]),
          "r"(p[17]), "r"(p[18]), "r"(p[19]), "r"(p[20]), "r"(p[21]), "r"(p[22]), "r"(p[23]), "r"(p[24]), "r"(p[25]),
          "r"(p[26]), "r"(p[27]), "r"(p[28]), "r"(p[29]), "r"(p[30]), "r"(p[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N);
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tbase + (uint32_t)((n_acc == 1 ? 0 : (it * 4 + k) % n_acc) * N);
        if (mode == 0) mma_ss(d, desc_sw128(smem_u32(sa) + k * 32), desc_sw128(smem_u32(sb) + k * 32), idesc, (uint32_t)((it | k) != 0));
        else mma_ts(d, a_tmem + k * 8, desc_sw128(smem_u32(sb) + k * 32), idesc, (uint32_t)((it | k) != 0));
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (d_out != nullptr && blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(lane_addr + c0) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) d_out[tid * 256 + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
  float *d_out; unsigned long long *cyc;
  cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
  const int smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float *h = (float *)malloc(128 * 256 * 4);
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {256, 128}) {
      cudaMemset(d_out, 0, 128 * 256 * 4);
      mma_bench<<<1, 128, smem>>>(mode, N, 1, d_out, cyc, 1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d N %d: error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, 128 * 256 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += a_val(r, k) * b_val(n, k);
          double err = fabs((double)h[r * 256 + n] - ref);
          if (err > maxerr) maxerr = err;
          if (err > 1e-3 && bad++ < 3) printf("   mismatch r %d n %d got %f ref %f\n", r, n, h[r * 256 + n], ref);
        }
      const int iters = 4000;
      printf("%s N=%3d: max|D-ref| = %g (%d bad)\n", mode ? "TS" : "SS", N, maxerr, bad);
      for (int n_acc : {1, 2}) {
        if (mode == 1 && n_acc * N > 256) continue;      // A operand lives at column 256
        mma_bench<<<148, 128, smem>>>(mode, N, iters, nullptr, cyc, n_acc);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d N %d: error %s\n", mode, N, cudaGetErrorString(e)); return 1; }
        unsigned long long hc[148];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < 148; ++i) if (hc[i] > mx) mx = hc[i];
        printf("    %d accumulator(s): %.1f cycles per MMA (M128 x N%d x K16), %.0f flop/cyc/SM\n", n_acc,
               (double)mx / (iters * 4.0), N, 2.0 * 128 * N * 16 * iters * 4.0 / (double)mx);
      }
    }
  return 0;
}
