// Library-level entry points: version, per-thread error string, launch counter.
#include "common.h"

namespace cpcsv {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace cpcsv

extern "C" int cpcsv_version(void) { return 100; }
extern "C" const char* cpcsv_last_error_string(void) { return cpcsv::g_err; }
extern "C" int64_t cpcsv_launch_count(void) {
  return cpcsv::g_launches.load(std::memory_order_relaxed);
}
