// Shared helpers for the sm_100a kernels behind include/nfs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/nfs_b200.h"

namespace nfs {

// ---- error plumbing (per host thread) --------------------------------------
void set_error(const char *where, const char *what);
int  fail_cuda(const char *where, cudaError_t e);
int  fail_arg(const char *where, int code, const char *what);
void count_launch(int n = 1);

// Checks the launch that was just issued on this thread.
static inline int check_launch(const char *where) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail_cuda(where, e); }
  count_launch();
  return 0;
}

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// A step is a few dozen dependent launches of 5-100 us each; every kernel's set-up (barrier init, TMEM allocation,
// tensor-map prefetch) and the previous kernel's tail (store drain, TMEM release, CTA exit) are pure latency.  Kernels
// launched through launch_dep() may start while their predecessor in the stream is still finishing: they run their
// set-up, then pdl_wait() blocks until the predecessor has completed and its writes are visible.  RULES for such a
// kernel: every thread executes pdl_wait() before its first access to global memory (reads AND writes - the
// predecessor may still be reading what this kernel overwrites); pdl_trigger() only after the CTA holds its TMEM
// allocation (a dependent CTA co-resident on the SM could otherwise take the columns and wait for this grid forever).
// Without the launch attribute both instructions do nothing.  Opt-in: NFS_PDL=1 (default: classic stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline int launch_dep(const char *where, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail_cuda(where, e); }
  return check_launch(where);
}

// cudaFuncSetAttribute is per device: remember per (call site, device) whether the opt-in shared-memory size has
// been set, so that a process driving several GPUs configures the kernel on each of them.
struct PerDeviceOnce {
  std::atomic<unsigned long long> done{0};          // bit d = done on device d (d < 64)
  bool need(int *dev_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    *dev_out = dev;
    return dev >= 64 || !((done.load(std::memory_order_acquire) >> dev) & 1ull);
  }
  void mark(int dev) { if (dev < 64) done.fetch_or(1ull << dev, std::memory_order_release); }
};

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned8(const void *p)  { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

constexpr unsigned kFullMask = 0xffffffffu;
__device__ __forceinline__ bool aligned16_dev(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- streaming loads/stores: every composite operand is touched exactly once,
// so keep it out of L1 (Guideline 13/14 of the Blackwell playbook).
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace nfs
