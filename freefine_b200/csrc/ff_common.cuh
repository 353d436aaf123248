// Shared helpers of libfreefine_b200.so: thread-local error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/freefine_b200.h"

namespace ff {

char* err_buf();  // thread-local, 512 bytes (ff_api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Call right after a kernel launch: reports launch-configuration errors without synchronising.
inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(FF_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return FF_OK;
}

}  // namespace ff

#define FF_REQUIRE(cond, ...)                            \
  do {                                                   \
    if (!(cond)) return ff::fail(FF_E_INVALID, __VA_ARGS__); \
  } while (0)
