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

namespace dcb {
enum KernelKind { K_ENCODE = 0, K_EMBED, K_INPROJ, K_CONV, K_OUTPROJ, K_FC1, K_FC2, K_HEAD1, K_HEAD2, K_SMOOTH, K_OTHER, K_SCONV, K_TOEP, K_MLP, K_BLOCK, K_NKINDS };
enum TraceKind { TRACE_NONE = 0, TRACE_INPROJ, TRACE_BLOCK, TRACE_TOEPLITZ };
// The default value of the ctx option "fft_min_len" means "let the measured cost model of model.cu:use_fft_conv pick
// between the blocked FFT long convolution (lconv.cu) and the tensor-core Toeplitz kernel (toeplitz.cu)"; any other
// value is a plain threshold on the padded batch length
constexpr int kDefaultFftMinLen = 6144;
struct ProfRec {
  int kind;
  cudaEvent_t a, b;
};
}  // namespace dcb

struct dcb200_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  int64_t launches = 0;
  // options (dcb200_ctx_set_option); the trace kind comes from DCB200_TRACE, read once when the ctx is created
  int fft_min_len = dcb::kDefaultFftMinLen;
  int smooth_warp_kernel = 0;  // option "smooth_warp_kernel": 0 = by launch size, 1 = always the warp-per-read kernel, 2 = always the tile kernel
  int trace_kind = dcb::TRACE_NONE;
  bool traced_once = false;
  // named workspaces (activations, staging), grow-only
  std::map<std::string, dcb::DevBuf> ws;
  dcb::DevBuf& buf(const char* name) { return ws[name]; }
  // kernels whose dynamic shared-memory limit has been raised on THIS device (the attribute is per device, and one
  // process may hold contexts on several)
  std::map<const void*, size_t> smem_attr;
  int ensure_smem(const void* func, size_t bytes) {
    auto it = smem_attr.find(func);
    if (it != smem_attr.end() && it->second >= bytes) return DCB200_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      dcb::set_error("cudaFuncSetAttribute(max dynamic smem = %zu) failed: %s", bytes, cudaGetErrorString(e));
      return DCB200_ECUDA;
    }
    smem_attr[func] = bytes;
    return DCB200_OK;
  }
  // optional per-kernel CUDA-event timing (bench.py's roofline numbers)
  bool profiling = false;
  std::vector<dcb::ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[dcb::K_NKINDS] = {0};
  int64_t prof_cnt[dcb::K_NKINDS] = {0};
  cudaEvent_t prof_event() {
    if (!prof_pool.empty()) {
      cudaEvent_t e = prof_pool.back();
      prof_pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

namespace dcb {
// Brackets the kernel launches of one scope with events on the ctx stream when profiling is on.
struct ProfScope {
  dcb200_ctx* ctx;
  ProfRec rec;
  bool on;
  ProfScope(dcb200_ctx* c, int kind) : ctx(c), on(c->profiling) {
    if (on) {
      rec.kind = kind;
      rec.a = ctx->prof_event();
      rec.b = ctx->prof_event();
      cudaEventRecord(rec.a, ctx->stream);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(rec.b, ctx->stream);
      ctx->prof_recs.push_back(rec);
    }
  }
};
}  // namespace dcb

#define DCB_LAUNCH_CHECK(ctx)                                                              \
  do {                                                                                     \
    (ctx)->launches++;                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      dcb::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return DCB200_ECUDA;                                                                 \
    }                                                                                      \
  } while (0)
