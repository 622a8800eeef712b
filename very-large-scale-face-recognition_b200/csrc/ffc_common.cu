#include <stdarg.h>

#include <atomic>

#include "ffc_common.cuh"

namespace ffc {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
}  // namespace ffc

extern "C" const char* ffc_last_error(void) { return ffc::g_err; }
extern "C" const char* ffc_version(void) { return "ffc_b200 0.1 (sm_100a; tcgen05/TMEM/TMA head, device LRU)"; }
extern "C" int64_t ffc_launch_count(void) { return ffc::g_launches.load(); }
