// api.cu -- version + thread-local error string of the C ABI (include/gsplat_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace gs {
static thread_local char g_error[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof g_error, fmt, ap);
  va_end(ap);
}
}  // namespace gs

extern "C" int gs_version(void) { return 100; }
extern "C" const char *gs_last_error_string(void) { return gs::g_error; }
