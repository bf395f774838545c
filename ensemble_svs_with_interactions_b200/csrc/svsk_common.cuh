// Shared host-side helpers for libsvsk: thread-local error string, argument checks, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/svsk.h"

namespace svsk {

char* last_error_buffer();  // thread-local, 512 bytes (svsk_api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

#define SVSK_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) return ::svsk::fail((code), __VA_ARGS__); \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// 0 when the current device is sm_100; caches the answer per device.
int require_sm100();

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace svsk
