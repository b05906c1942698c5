// Error reporting and versioning for the pcab200 C ABI.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "pcab200.h"

static thread_local char g_err[512] = "";

void pcab_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pcab_last_error(void) { return g_err; }
extern "C" int pcab_version(void) { return 200; }

int pcab_sm_count() {
  static int cached[64] = {};
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) return 148;
  const bool tracked = d >= 0 && d < 64;
  if (tracked) {
    const int v = __atomic_load_n(&cached[d], __ATOMIC_ACQUIRE);
    if (v > 0) return v;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
  if (tracked) __atomic_store_n(&cached[d], n, __ATOMIC_RELEASE);
  return n;
}
