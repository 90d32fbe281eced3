// C-ABI plumbing: thread-local error string, launch checking, version / capability queries.
#include "common.cuh"
#include "vaesne_b200.h"
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

namespace vaesne {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

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
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return V_ECUDA;
  }
  return V_OK;
}
}  // namespace vaesne

extern "C" const char* vaesne_last_error(void) { return vaesne::g_err; }

extern "C" long long vaesne_launch_count(void) { return vaesne::g_launches.load(); }

extern "C" int vaesne_abi_version(void) { return VAESNE_B200_ABI_VERSION; }

extern "C" int vaesne_is_emulated(void) {
#ifdef VAESNE_EMU
  return 1;
#else
  return 0;
#endif
}
