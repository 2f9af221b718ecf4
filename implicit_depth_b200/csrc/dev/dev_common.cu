// Error plumbing of the dev-probe library (libb200probe.so): same contract as the product library's.
#include <stdarg.h>
#include <stdio.h>

#include "../common.cuh"

static thread_local char g_err[512] = "";
extern "C" const char* b200_last_error(void) { return g_err; }
void b200_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
