// Library-level plumbing: error text, launch counter, device attributes.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "hg_common.cuh"

namespace hg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("HG_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace hg

#ifdef HG_TRACE
namespace hg {
int trace_set_gemm(void* p);
int trace_set_env(void* p);
int trace_set_heads(void* p);
}  // namespace hg
// trace builds only (not part of the ABI): device timeline table, u64 [4 * 64 * 4 * 512] (hg_common.cuh); null switches off
extern "C" int hg_trace_set(void* buf) { return hg::trace_set_gemm(buf) | hg::trace_set_env(buf) | hg::trace_set_heads(buf); }
#endif

extern "C" const char* hg_last_error(void) { return hg::g_err; }
extern "C" int hg_abi_version(void) { return HG_ABI_VERSION; }
extern "C" int64_t hg_launch_count(void) { return hg::g_launches.load(std::memory_order_relaxed); }
