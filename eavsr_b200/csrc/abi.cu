// Library-level pieces of the C ABI: version, per-thread error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <cuda.h>

#include "common.cuh"

namespace eavsr {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch error so that the next call starts clean
    set_error("%s: %s", what, cudaGetErrorString(e));
    return EAVSR_ERR_CUDA;
  }
  return EAVSR_OK;
}

bool encode_tensor_map(void* tensor_map, int dtype, int rank, const void* base, const unsigned long long* dims,
                       const unsigned long long* strides_bytes, const unsigned* box, int swizzle) {
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Fn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<Fn>(p);
  }();
  if (!fn || rank < 1 || rank > 5) return false;
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  return fn(reinterpret_cast<CUtensorMap*>(tensor_map),
            dtype == EAVSR_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
            const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace eavsr

extern "C" int eavsr_version(void) { return 100; }
extern "C" const char* eavsr_last_error(void) { return eavsr::g_err; }
extern "C" uint64_t eavsr_launch_count(void) { return eavsr::g_launches.load(std::memory_order_relaxed); }
