// Version + thread-local error reporting of libb200ret.
#include <stdarg.h>

#include "common.cuh"

namespace b2r {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b2r

extern "C" int b2r_version(void) { return B2R_VERSION; }
extern "C" const char *b2r_last_error(void) { return b2r::g_err; }
extern "C" unsigned long long b2r_launch_count(void) {
    return __atomic_load_n(&b2r::g_launches, __ATOMIC_RELAXED);
}
