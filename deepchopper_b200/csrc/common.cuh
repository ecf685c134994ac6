// Shared plumbing of libdcb200: error reporting, the ctx (device + stream + workspaces).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/dcb200.h"

namespace dcb {

void set_error(const char* fmt, ...);

#define DCB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      dcb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DCB200_ECUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define DCB_CHECK(rc_expr)            \
  do {                                \
    int _rc = (rc_expr);              \
    if (_rc != DCB200_OK) return _rc; \
  } while (0)

#define DCB_ARG(cond)                                                      \
  do {                                                                     \
    if (!(cond)) {                                                         \
      dcb::set_error("invalid argument: %s (%s:%d)", #cond, __FILE__, __LINE__); \
      return DCB200_EINVAL;                                                \
    }                                                                      \
  } while (0)

// A grow-only device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return DCB200_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return DCB200_ENOMEM;
    }
    cap = want;
    return DCB200_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace dcb

struct dcb200_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int64_t launches = 0;
  // named workspaces (activations, staging), grow-only
  std::map<std::string, dcb::DevBuf> ws;
  dcb::DevBuf& buf(const char* name) { return ws[name]; }
};

#define DCB_LAUNCH_CHECK(ctx)                                                              \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      dcb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DCB200_ECUDA;                                                                 \
    }                                                                                      \
  } while (0)
