// Error string, ABI version and launch counter of libmcnerf.so.
#include <atomic>
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void mcnerf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void mcnerf_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" int mcnerf_abi_version(void) { return MCNERF_ABI_VERSION; }
extern "C" const char* mcnerf_last_error(void) { return g_err; }
extern "C" uint64_t mcnerf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
