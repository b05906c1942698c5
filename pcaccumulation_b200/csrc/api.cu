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
extern "C" int pcab_version(void) { return 100; }
