// Microbenchmark: the tcgen05.mma stream of the weight-gradient kernel (wgrad_body.cuh) without its TMA loads - one
// 64-point slab resident in shared memory, [64 x 64] SWIZZLE_128B boxes read as MN-major operands, M = 2 x 128 output
// rows, N = 256, four K = 16 steps per slab.  Variants: operand major-ness (MN as in wgrad / K-major for reference) and
// issue order (accumulator alternating per MMA = round-1 order / runs of 4 per accumulator / one accumulator only).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nerf-few-shot-limitations_b200/csrc -o mma_wgrad mma_wgrad.cu \
//      nerf-few-shot-limitations_b200/csrc/tc_host.cu nerf-few-shot-limitations_b200/csrc/api.cu -lcuda
#include "tc_common.cuh"
#include <cstdio>
using namespace nfs;
using namespace nfs::tc;

__global__ void __launch_bounds__(128, 1) bench(int mn_major, int order, int iters, unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t *bar = (uint64_t *)(smem + 65536);
  uint32_t *slot = (uint32_t *)(bar + 1);
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3C003C00u + (uint32_t)(i * 2654435761u >> 28);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 256, mn_major, mn_major);
    const uint32_t ua = smem_u32(smem), va = ua + 4 * 8192;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (order == 0) {               // alternating accumulators (k outer, h inner)
        for (int k = 0; k < 4; ++k)
          for (int h = 0; h < 2; ++h) {
            const uint64_t ad = mn_major ? umma_desc_sw128(ua + h * 2 * 8192 + k * 2048, 8192, 1024) : umma_desc_sw128(ua + h * 16384 + k * 32, 16, 1024);
            const uint64_t bd = mn_major ? umma_desc_sw128(va + k * 2048, 8192, 1024) : umma_desc_sw128(va + k * 32, 16, 1024);
            umma_bf16(tbase + h * 256, ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
      } else {                        // runs of 4 per accumulator (order 1) / a single accumulator (order 2)
        for (int h = 0; h < 2; ++h)
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = mn_major ? umma_desc_sw128(ua + h * 2 * 8192 + k * 2048, 8192, 1024) : umma_desc_sw128(ua + h * 16384 + k * 32, 16, 1024);
            const uint64_t bd = mn_major ? umma_desc_sw128(va + k * 2048, 8192, 1024) : umma_desc_sw128(va + k * 32, 16, 1024);
            umma_bf16(tbase + (order == 2 ? 0 : h * 256), ad, bd, idesc, (uint32_t)((it | k) != 0));
          }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

int main() {
  unsigned long long *cyc;
  cudaMalloc(&cyc, 148 * 8);
  const int smem = 65536 + 64;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int grid : {148, 74})
    for (int mn = 1; mn >= 0; --mn)
      for (int order = 0; order < 3; ++order) {
        bench<<<grid, 128, smem>>>(mn, order, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long hc[148];
        cudaMemcpy(hc, cyc, grid * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx = 0; for (int i = 0; i < grid; ++i) if (hc[i] > mx) mx = hc[i];
        printf("grid %3d  %s-major  order %d (%s): %.1f cycles per MMA (128x256x16), %.0f cycles per 64-point slab\n", grid,
               mn ? "MN" : "K ", order, order == 0 ? "alternating" : order == 1 ? "runs of 4" : "one accumulator",
               (double)mx / (iters * 8.0), (double)mx / iters);
      }
  return 0;
}
