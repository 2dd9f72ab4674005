// Library-level pieces of the C ABI: version, per-thread error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

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

}  // namespace eavsr

extern "C" int eavsr_version(void) { return 100; }
extern "C" const char* eavsr_last_error(void) { return eavsr::g_err; }
extern "C" uint64_t eavsr_launch_count(void) { return eavsr::g_launches.load(std::memory_order_relaxed); }
