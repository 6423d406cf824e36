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
  // Off unless MMREC_PDL=1: measured on the B200 (profiles/r02_pdl.txt) the chained SpMM launches of a
  // propagation get 4 % shorter (12.7 -> 12.15 us each, bit-identical results), the training step does
  // not move (2.370 ms both ways: the other streams of the step fill the same gaps), and the dependents'
  // early start inflates every per-kernel duration a profiler reports.
  const char *e = getenv("MMREC_PDL");
  return e && atoi(e) != 0;
}
}  // namespace mmrec

extern "C" int mmrec_abi_version(void) { return 1; }
extern "C" const char *mmrec_last_error(void) { return mmrec::g_err; }
extern "C" int64_t mmrec_launch_count(void) { return mmrec::g_launches.load(); }
