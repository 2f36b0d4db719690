// Library-wide state of the C ABI: thread-local error text, launch counter,
// device properties.
#include "common.cuh"
#include <stdarg.h>

namespace grasp {

static thread_local char tl_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace grasp

extern "C" int grasp_abi_version(void) { return GRASP_ABI_VERSION; }
extern "C" const char* grasp_last_error(void) { return grasp::tl_error; }
extern "C" uint64_t grasp_launch_count(void) { return grasp::g_launches.load(std::memory_order_relaxed); }
