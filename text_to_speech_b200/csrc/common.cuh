// Shared host-side error plumbing for libwg_b200.so (no exception crosses the C ABI: every entry
// point catches wg::Fail and turns it into a wg_status + message).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/wg_b200.h"

namespace wg {

struct Fail {
  int code;
  std::string msg;
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Fail{code, buf};
}

#define WG_CK(expr)                                                                                   \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      ::wg::fail(WG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Entry points run on the engine's device and put the caller's current device back on exit (torch reads it with
// cudaGetDevice: a process driving several engines must not find it changed behind its back).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    WG_CK(cudaSetDevice(dev));   // always: this also binds the device's primary context to a thread that has none yet
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

}  // namespace wg
