// Error reporting, version and launch accounting of libmpa_b200.so.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

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

namespace mpa {
// ---- optional per-kernel CUDA-event profiler (bench.py roofline leg) ----
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
std::atomic<int> g_prof_on{0};

ProfScope::ProfScope(const char* name, cudaStream_t stream) : name_(name), stream_(stream) {
  active_ = g_prof_on.load(std::memory_order_relaxed) != 0;
  if (!active_) return;
  if (cudaEventCreate(&a_) != cudaSuccess || cudaEventCreate(&b_) != cudaSuccess) {
    active_ = false;
    return;
  }
  cudaEventRecord(a_, stream_);
}
ProfScope::~ProfScope() {
  if (!active_) return;
  cudaEventRecord(b_, stream_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back({name_, a_, b_});
}
}  // namespace mpa

extern "C" {
void mpa_profile_enable(int on) { mpa::g_prof_on.store(on ? 1 : 0); }

/* Waits for the recorded events, writes "name launches total_ms\n" lines into
 * buf (NUL terminated), clears the records; returns the number of bytes needed. */
size_t mpa_profile_report(char* buf, size_t cap) {
  std::lock_guard<std::mutex> lk(mpa::g_prof_mu);
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : mpa::g_prof) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  mpa::g_prof.clear();
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf != nullptr && cap > 0) {
    size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return out.size() + 1;
}

const char* mpa_last_error(void) { return mpa::g_err; }
int mpa_version(void) { return 100; }
uint64_t mpa_launch_count(void) { return mpa::g_launches.load(); }
}
