// Error state, version and launch accounting of libb200ssl.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace b200ssl {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace b200ssl

extern "C" {

int b200ssl_version(void) { return B200SSL_VERSION; }

const char* b200ssl_last_error(void) { return b200ssl::g_err; }

long long b200ssl_launch_count(void) {
  return b200ssl::g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
