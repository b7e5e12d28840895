// api.cu -- process-wide plumbing of libvitk: error text, launch counter, ABI version.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "vitk_common.cuh"

namespace vitk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VITK_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

}  // namespace vitk

extern "C" int vitk_abi_version(void) { return VITK_ABI_VERSION; }
extern "C" const char* vitk_last_error(void) { return vitk::g_err; }
extern "C" int64_t vitk_launch_count(void) { return vitk::g_launches.load(std::memory_order_relaxed); }
extern "C" void vitk_reset_launch_count(void) { vitk::g_launches.store(0, std::memory_order_relaxed); }
