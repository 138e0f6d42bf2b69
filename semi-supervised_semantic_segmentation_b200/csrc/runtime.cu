// Error state, version and launch accounting of libb200ssl.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-kernel CUDA-event timing (off by default; bench.py switches it on for one extra pass) ----
struct ProfRec {
  const char* name;
  cudaEvent_t start, stop;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
static thread_local ProfRec g_open = {nullptr, nullptr, nullptr};
static thread_local cudaStream_t g_open_stream = nullptr;

void prof_begin(const char* name, cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r;
  r.name = name;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, stream);
  g_open = r;
  g_open_stream = stream;
}

void prof_end() {
  if (!g_open.name) return;
  cudaEventRecord(g_open.stop, g_open_stream);
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(g_open);
  }
  g_open.name = nullptr;
}

}  // namespace b200ssl

extern "C" {

int b200ssl_version(void) { return B200SSL_VERSION; }

const char* b200ssl_last_error(void) { return b200ssl::g_err; }

long long b200ssl_launch_count(void) {
  return b200ssl::g_launches.load(std::memory_order_relaxed);
}

size_t b200ssl_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(b200ssl_ema_chunk);
    case 1: return sizeof(b200ssl_lovasz_desc);
    case 2: return sizeof(b200ssl_step_desc);
    case 3: return sizeof(b200ssl_sgd_chunk);
    case 4: return sizeof(b200ssl_sgd_hyper);
    default: return 0;
  }
}

void b200ssl_prof_enable(int on) { b200ssl::g_prof_on.store(on ? 1 : 0); }

long long b200ssl_prof_report(char* buf, size_t capacity) {
  using namespace b200ssl;
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof_recs);
  }
  struct Agg { long long n = 0; double ms = 0.0; double min_ms = 1e30; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (const ProfRec& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      if (!agg.count(r.name)) order.push_back(r.name);
      Agg& a = agg[r.name];
      a.n += 1;
      a.ms += ms;
      if (ms < a.min_ms) a.min_ms = ms;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  std::string out;
  char line[256];
  for (const std::string& k : order) {
    const Agg& a = agg[k];
    snprintf(line, sizeof(line), "%s %lld %.6f %.6f\n", k.c_str(), a.n, a.ms, a.min_ms);
    out += line;
  }
  if (buf && capacity) {
    const size_t n = out.size() < capacity - 1 ? out.size() : capacity - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (long long)out.size();
}

}  // extern "C"
