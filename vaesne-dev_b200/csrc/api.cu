// C-ABI plumbing: thread-local error string, launch checking, version / capability queries.
#include "common.cuh"
#include "vaesne_b200.h"
#include <stdarg.h>
#include <stdio.h>

namespace vaesne {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
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

extern "C" int vaesne_abi_version(void) { return VAESNE_B200_ABI_VERSION; }

extern "C" int vaesne_is_emulated(void) {
#ifdef VAESNE_EMU
  return 1;
#else
  return 0;
#endif
}
