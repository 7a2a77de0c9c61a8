// Library-level entry points: version and thread-local error string.
#include <stdarg.h>

#include "common.cuh"

namespace rcb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rcb

extern "C" int rcb_version(void) { return 100; }
extern "C" const char* rcb_last_error(void) { return rcb::g_err; }
