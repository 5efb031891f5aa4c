// Data-parallel weight update — the sum of the flat fp32 gradient over the GPUs of one NVLink / NVSwitch box FUSED into
// the Adam step: every rank reads its peers' gradient buffers directly over NVLink (peer memory mapped through CUDA IPC)
// while it updates its replica of the weights.  One kernel per step, capturable in the step's CUDA graph; no NCCL
// collective and no host synchronisation on the data path.
//
// The reference has no multi-GPU path ("new functionality"); what is replaced is
// optimizer.step() of src/training/train.py:286 for a ray-sharded batch whose loss mean spans all ranks.
//
// Protocol (epoch e = number of exchanges so far + 1; all flags live in the RECEIVING rank's buffer):
//   ready[r][s] = e   rank s's gradient of exchange e is complete         (written by s into r's buffer)
//   done [r][s] = e   rank s has finished reading r's gradient of exchange e
//   nfs_dp_adam_step: CTA 0 publishes ready to every peer; every CTA waits for all peers' ready, sums the W gradient
//                     buffers element-wise (own: local, peers: ld.relaxed.sys over NVLink), applies Adam; the last CTA
//                     to finish publishes done to every peer.
//   nfs_dp_wait_readers: first kernel of the NEXT step - waits until every peer's done has reached the last epoch
//                     before this rank's gradient buffer is zeroed and accumulated into again.
// Every wait is bounded (a lost peer becomes a CUDA error after ~2 s, never a hung GPU).
#include "nfs_common.cuh"

#include <cmath>

namespace nfs {
namespace {

constexpr int kDpMaxWorld = 8;

struct DpPeers {
  const float *grad[kDpMaxWorld];      // grad[r] = rank r's gradient buffer (grad[rank] is local)
  unsigned *ready[kDpMaxWorld];        // ready[r] = rank r's ready[world] array
  unsigned *done[kDpMaxWorld];         // done[r]  = rank r's done[world] array
  int world, rank;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_relaxed_sys4(const float *p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys1(const float *p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// thread r (< world, != rank) of the calling block spins until flags[r] >= epoch
__device__ __forceinline__ void wait_flags(const unsigned *flags, unsigned epoch, const DpPeers &p, const char *what) {
  const int r = threadIdx.x;
  if (r < p.world && r != p.rank) {
    for (unsigned spin = 0;; ++spin) {
      if (ld_acquire_sys(flags + r) >= epoch) break;
      __nanosleep(100);
      if (spin > (1u << 24)) {
        printf("nfs_b200: rank %d timed out waiting for rank %d (%s, epoch %u)\n", p.rank, r, what, epoch);
        __trap();
      }
    }
  }
}

__global__ void dp_wait_readers_kernel(const DpPeers p, const unsigned *epoch_dev) {
  wait_flags(p.done[p.rank], *epoch_dev, p, "gradient readers");
}

__device__ __forceinline__ float adam_one(float w, float grad, float &m, float &v, float lr, float b1, float b2, float eps,
                                          float wd, float bc1, float bc2_sqrt, int decoupled) {
  if (wd != 0.f) {
    if (decoupled) w *= 1.f - lr * wd;      // AdamW
    else grad += wd * w;                    // Adam (L2)
  }
  m = b1 * m + (1.f - b1) * grad;
  v = b2 * v + (1.f - b2) * grad * grad;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  return w - (lr / bc1) * (m / denom);
}

// n4 = n / 4 float4 elements + (n & 3) scalar tail; state = [1 - beta1^t, sqrt(1 - beta2^t), lr] (device, as adam_kernel)
__global__ void __launch_bounds__(256)
dp_adam_kernel(float *__restrict__ param, float *__restrict__ m, float *__restrict__ v, long long n, float b1, float b2,
               float eps, float wd, float gscale, int decoupled, const float *__restrict__ state, const DpPeers p,
               unsigned *epoch_dev, unsigned *cta_counter) {
  const unsigned epoch = *epoch_dev + 1;
  if (blockIdx.x == 0 && threadIdx.x < p.world && (int)threadIdx.x != p.rank) {
    __threadfence_system();                                 // this rank's gradient (earlier kernels) before the flag
    st_release_sys(p.ready[threadIdx.x] + p.rank, epoch);
  }
  wait_flags(p.ready[p.rank], epoch, p, "gradients");
  __syncthreads();
  const float bc1 = __ldg(state), bc2_sqrt = __ldg(state + 1), lr = __ldg(state + 2);
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kDpMaxWorld; ++r) {
      if (r < p.world) {
        const float4 t = (r == p.rank) ? reinterpret_cast<const float4 *>(p.grad[r])[i] : ld_relaxed_sys4(p.grad[r] + 4 * i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
    }
    float4 w = reinterpret_cast<float4 *>(param)[i], mi = reinterpret_cast<float4 *>(m)[i], vi = reinterpret_cast<float4 *>(v)[i];
    w.x = adam_one(w.x, g.x * gscale, mi.x, vi.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, decoupled);
    w.y = adam_one(w.y, g.y * gscale, mi.y, vi.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, decoupled);
    w.z = adam_one(w.z, g.z * gscale, mi.z, vi.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, decoupled);
    w.w = adam_one(w.w, g.w * gscale, mi.w, vi.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, decoupled);
    reinterpret_cast<float4 *>(param)[i] = w;
    reinterpret_cast<float4 *>(m)[i] = mi;
    reinterpret_cast<float4 *>(v)[i] = vi;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = 4 * n4 + threadIdx.x;
    float g = 0.f;
    for (int r = 0; r < p.world; ++r) g += (r == p.rank) ? p.grad[r][i] : ld_relaxed_sys1(p.grad[r] + i);
    float mi = m[i], vi = v[i];
    param[i] = adam_one(param[i], g * gscale, mi, vi, lr, b1, b2, eps, wd, bc1, bc2_sqrt, decoupled);
    m[i] = mi; v[i] = vi;
  }
  // the last CTA to finish tells every peer that this rank no longer reads their gradient of this epoch
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicAdd(cta_counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x < p.world && (int)threadIdx.x != p.rank) st_release_sys(p.done[threadIdx.x] + p.rank, epoch);
    if (threadIdx.x == 0) { *cta_counter = 0; *epoch_dev = epoch; }
  }
}

__global__ void dp_tick_kernel(int *step, float b1, float b2, float *state) {
  const int t = *step + 1;
  *step = t;
  state[0] = (float)(1.0 - pow((double)b1, (double)t));
  state[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
}

int fill_peers(const char *fn, void *const *peer_bases, int32_t world, int32_t rank, int64_t n, DpPeers *out) {
  if (world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world || !peer_bases)
    return fail_arg(fn, NFS_E_BADARG, "need 1..8 ranks and their exchange buffers");
  DpPeers p{};
  p.world = world; p.rank = rank;
  const size_t flags_off = nfs_dp_flags_offset(n);
  for (int r = 0; r < world; ++r) {
    if (!peer_bases[r]) return fail_arg(fn, NFS_E_BADARG, "null exchange buffer");
    char *b = static_cast<char *>(peer_bases[r]);
    p.grad[r] = reinterpret_cast<const float *>(b);
    p.ready[r] = reinterpret_cast<unsigned *>(b + flags_off);
    p.done[r] = p.ready[r] + kDpMaxWorld;
  }
  *out = p;
  return 0;
}

}  // namespace
}  // namespace nfs

using namespace nfs;

extern "C" uint64_t nfs_dp_flags_offset(int64_t n_grad) { return (((uint64_t)n_grad * 4u + 255u) / 256u) * 256u; }
extern "C" uint64_t nfs_dp_buffer_bytes(int64_t n_grad) { return nfs_dp_flags_offset(n_grad) + 256u; }

extern "C" int nfs_dp_alloc(int64_t n_grad, void **base_out) {
  const char *fn = "nfs_dp_alloc";
  if (n_grad <= 0 || !base_out) return fail_arg(fn, NFS_E_BADARG, "bad size / null pointer");
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, nfs_dp_buffer_bytes(n_grad));     // plain cudaMalloc: exportable through CUDA IPC
  if (e != cudaSuccess) return fail_cuda(fn, e);
  e = cudaMemset(p, 0, nfs_dp_buffer_bytes(n_grad));
  if (e != cudaSuccess) { cudaFree(p); return fail_cuda(fn, e); }
  *base_out = p;
  return 0;
}
extern "C" int nfs_dp_free(void *base) {
  cudaError_t e = cudaFree(base);
  return e == cudaSuccess ? 0 : fail_cuda("nfs_dp_free", e);
}
extern "C" int nfs_dp_ipc_export(void *base, void *handle64) {
  if (!base || !handle64) return fail_arg("nfs_dp_ipc_export", NFS_E_BADARG, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), base);
  return e == cudaSuccess ? 0 : fail_cuda("nfs_dp_ipc_export", e);
}
extern "C" int nfs_dp_ipc_open(const void *handle64, void **base_out) {
  if (!handle64 || !base_out) return fail_arg("nfs_dp_ipc_open", NFS_E_BADARG, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : fail_cuda("nfs_dp_ipc_open", e);
}
extern "C" int nfs_dp_ipc_close(void *base) {
  cudaError_t e = cudaIpcCloseMemHandle(base);
  return e == cudaSuccess ? 0 : fail_cuda("nfs_dp_ipc_close", e);
}

extern "C" int nfs_dp_wait_readers(void *const *peer_bases, int32_t world, int32_t rank, int64_t n_grad,
                                   const uint32_t *epoch_dev, void *stream) {
  const char *fn = "nfs_dp_wait_readers";
  DpPeers p{};
  int rc = fill_peers(fn, peer_bases, world, rank, n_grad, &p);
  if (rc) return rc;
  if (!epoch_dev) return fail_arg(fn, NFS_E_BADARG, "null epoch counter");
  dp_wait_readers_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p, epoch_dev);
  return check_launch(fn);
}

extern "C" int nfs_dp_adam_step(float *param, void *const *peer_bases, int32_t world, int32_t rank, float *exp_avg,
                                float *exp_avg_sq, int64_t n, float beta1, float beta2, float eps, float weight_decay,
                                int32_t *step_counter, float *state, float grad_scale, int32_t decoupled,
                                uint32_t *epoch_dev, uint32_t *cta_counter, void *stream) {
  const char *fn = "nfs_dp_adam_step";
  if (n <= 0) return fail_arg(fn, NFS_E_BADARG, "n <= 0");
  if (!param || !exp_avg || !exp_avg_sq || !step_counter || !state || !epoch_dev || !cta_counter)
    return fail_arg(fn, NFS_E_BADARG, "null tensor pointer");
  if (!aligned16(param) || !aligned16(exp_avg) || !aligned16(exp_avg_sq))
    return fail_arg(fn, NFS_E_ALIGN, "parameter / moment buffers must be 16-byte aligned");
  DpPeers p{};
  int rc = fill_peers(fn, peer_bases, world, rank, n, &p);
  if (rc) return rc;
  dp_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter, beta1, beta2, state);
  rc = check_launch(fn);
  if (rc) return rc;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > 2LL * sms) blocks = 2LL * sms;          // all CTAs are resident: none waits for a peer while another is queued
  if (blocks < 1) blocks = 1;
  dp_adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, exp_avg, exp_avg_sq, n, beta1, beta2, eps,
                                                                    weight_decay, grad_scale, decoupled, state, p, epoch_dev,
                                                                    cta_counter);
  return check_launch(fn);
}
