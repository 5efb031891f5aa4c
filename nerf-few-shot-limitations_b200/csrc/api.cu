// Error plumbing and bookkeeping shared by every entry point of include/nfs_b200.h.
#include <stdlib.h>
#include "nfs_common.cuh"

#include <string.h>

namespace nfs {

static thread_local char g_err[512] = "no error";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *where, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, what);
}

int fail_cuda(const char *where, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s): %s", where, (int)e, cudaGetErrorName(e),
           cudaGetErrorString(e));
  return (int)e;
}

int fail_arg(const char *where, int code, const char *what) {
  snprintf(g_err, sizeof(g_err), "%s: invalid argument (%d): %s", where, code, what);
  return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

bool pdl_enabled() {
  // opt-in (NFS_PDL=1): worth 2-4 % of the many-launch cfg 4 step and nothing on the three-kernel headline step
  static const bool on = [] { const char *v = getenv("NFS_PDL"); return v != nullptr && v[0] == '1'; }();
  return on;
}

}  // namespace nfs

extern "C" int nfs_abi_version(void) { return NFS_B200_ABI_VERSION; }
extern "C" const char *nfs_last_error_string(void) { return nfs::g_err; }
extern "C" uint64_t nfs_launch_count(void) { return nfs::g_launches.load(std::memory_order_relaxed); }
