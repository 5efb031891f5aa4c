// Host side of the TMA plumbing: tensor-map encoding through the driver entry point
// (resolved at run time so the library has no link-time dependency on libcuda and still
// loads on a machine without a driver).
#include "tc_common.cuh"

#include <mutex>

namespace nfs {
namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;

static void resolve() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) g_encode = (EncodeTiledFn)fn;
  else (void)cudaGetLastError();
}

int make_tmap_bf16(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                   uint32_t box_rows, const char *where) {
  std::call_once(g_once, resolve);
  if (!g_encode) return fail_arg(where, NFS_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((pitch_elems * 2) & 15u))
    return fail_arg(where, NFS_E_ALIGN, "bf16 operand needs a 16-byte aligned base and row pitch");
  if (box_rows == 0 || box_rows > 256) return fail_arg(where, NFS_E_TOOLARGE, "TMA box rows must be in [1,256]");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return fail_arg(where, NFS_E_UNSUPPORTED, msg);
  }
  return 0;
}

int make_tmap_rows16(CUtensorMap *map, const void *base, uint64_t rows, uint32_t box_rows, const char *where) {
  std::call_once(g_once, resolve);
  if (!g_encode) return fail_arg(where, NFS_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  if (reinterpret_cast<uintptr_t>(base) & 15u) return fail_arg(where, NFS_E_ALIGN, "operand needs a 16-byte aligned base");
  if (box_rows == 0 || box_rows > 256) return fail_arg(where, NFS_E_TOOLARGE, "TMA box rows must be in [1,256]");
  cuuint64_t gdim[2] = {8, rows};
  cuuint64_t gstride[1] = {16};
  cuuint32_t box[2] = {8, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return fail_arg(where, NFS_E_UNSUPPORTED, msg);
  }
  return 0;
}

}  // namespace tc
}  // namespace nfs
