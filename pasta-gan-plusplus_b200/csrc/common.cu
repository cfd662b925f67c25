#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace pgpp {
static thread_local char t_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

static EnvFlags read_env() {
    EnvFlags f;
    f.igemm_no_reuse = getenv("PGPP_IGEMM_NO_REUSE") != nullptr;
    f.igemm_no_slab2 = getenv("PGPP_IGEMM_NO_SLAB2") != nullptr;
    f.igemm_no_resident = getenv("PGPP_IGEMM_NO_RESIDENT") != nullptr;
    f.igemm_no_stack = getenv("PGPP_IGEMM_NO_STACK") != nullptr;
    f.igemm_no_lean_epilogue = getenv("PGPP_IGEMM_NO_LEAN_EPILOGUE") != nullptr;
    f.igemm_no_tma_store = getenv("PGPP_IGEMM_NO_TMA_STORE") != nullptr;
    f.igemm_slab9 = getenv("PGPP_IGEMM_SLAB9") != nullptr;
    f.wgrad_no_reuse = getenv("PGPP_WGRAD_NO_REUSE") != nullptr;
    f.wgrad_no_pair = getenv("PGPP_WGRAD_NO_PAIR") != nullptr;
    f.ba_nostream = getenv("PGPP_BA_NOSTREAM") != nullptr;
    f.fir_packed_no_tile = getenv("PGPP_FIR_PACKED_NO_TILE") != nullptr;
    const char* e = getenv("PGPP_IGEMM_DEBUG");
    f.igemm_debug = e ? atoi(e) : 0;
    return f;
}
static EnvFlags g_env = read_env();
const EnvFlags& env_flags() { return g_env; }
} // namespace pgpp

extern "C" void pgpp_refresh_env(void) { pgpp::g_env = pgpp::read_env(); }

extern "C" int pgpp_version(void) { return 100; }
extern "C" const char* pgpp_last_error(void) { return pgpp::t_error; }
extern "C" int64_t pgpp_launch_count(void) { return (int64_t)pgpp::g_launches.load(); }
