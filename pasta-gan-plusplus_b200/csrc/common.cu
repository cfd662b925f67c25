#include "common.cuh"
#include <string.h>

namespace pgpp {
static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}
} // namespace pgpp

extern "C" int pgpp_version(void) { return 100; }
extern "C" const char* pgpp_last_error(void) { return pgpp::t_error; }
extern "C" int64_t pgpp_launch_count(void) { return (int64_t)pgpp::g_launches.load(); }
