// Error reporting, version and launch accounting of libmpa_b200.so.
#include <stdarg.h>
#include <string.h>

#include "mpa_common.cuh"

namespace mpa {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mpa

extern "C" {
const char* mpa_last_error(void) { return mpa::g_err; }
int mpa_version(void) { return 100; }
uint64_t mpa_launch_count(void) { return mpa::g_launches.load(); }
}
