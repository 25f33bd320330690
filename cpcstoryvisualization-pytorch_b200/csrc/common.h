// Shared host-side helpers: error reporting, launch counting, checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/cpcsv.h"

namespace cpcsv {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// call after every kernel launch
inline int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

#define CPCSV_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) return ::cpcsv::fail(-1, __VA_ARGS__); \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int num_sms();

}  // namespace cpcsv
