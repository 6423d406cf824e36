// Error reporting and launch accounting for the C ABI.
#include <atomic>
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace mmrec {
namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {   // read per call so that one process can compare both settings
  const char *e = getenv("MMREC_PDL");
  return !(e && atoi(e) == 0);
}
}  // namespace mmrec

extern "C" int mmrec_abi_version(void) { return 1; }
extern "C" const char *mmrec_last_error(void) { return mmrec::g_err; }
extern "C" int64_t mmrec_launch_count(void) { return mmrec::g_launches.load(); }
