// Shared host-side helpers for the C ABI: error reporting and argument checks.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/sageattn_b200.h"

namespace sab {

char* error_buffer();  // thread-local, 512 bytes (capi.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define SAB_REQUIRE(cond, code, ...)                 \
  do {                                               \
    if (!(cond)) return ::sab::fail(code, __VA_ARGS__); \
  } while (0)

#define SAB_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::sab::fail(SAB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                               \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace sab
