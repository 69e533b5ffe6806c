// Library-level entry points: ABI version, thread-local error text, launch counter.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_layout.cuh"

namespace effq {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

long long stream_cap(int per_sm) {
  static const bool persist = [] { const char* v = getenv("EFFQ_STREAM_PERSIST"); return v && *v == '1'; }();
  return persist ? (long long)sm_count() * per_sm : (1ll << 30);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

}  // namespace effq

extern "C" {
int effq_abi_version(void) { return EFFQ_ABI_VERSION; }
const char* effq_last_error(void) { return effq::g_err; }
uint64_t effq_launch_count(void) { return effq::g_launches.load(std::memory_order_relaxed); }
void effq_reset_launch_count(void) { effq::g_launches.store(0, std::memory_order_relaxed); }
}
