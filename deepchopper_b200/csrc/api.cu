// extern "C" surface of libdcb200 (see include/dcb200.h).
#include "common.cuh"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

namespace dcb {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int encode_device(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                  const int32_t* len, const int32_t* lpad_rows, int32_t R, int32_t Lpad, int32_t Lrow, uint8_t* tok,
                  float* qual);
int smooth_chop_device(dcb200_ctx* ctx, const int8_t* labels, const float* logits, int64_t total, const int64_t* starts,
                       const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                       int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action,
                       int8_t* smoothed);
int weights_create(dcb200_ctx* ctx, const char* const* names, const float* const* data, const int64_t* numel,
                   int32_t n, dcb200_weights** out);
int weights_destroy(dcb200_weights* w);
int forward_device(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* tok, const float* qual, int32_t B, int32_t L,
                   float* logits, uint8_t* labels, int stop_stage);

__global__ void row_starts_kernel(const int32_t* __restrict__ len, const int32_t* __restrict__ lpad_rows, int R, int Lrow,
                                  int Lpad, int64_t* __restrict__ starts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) starts[r] = (int64_t)r * Lrow + ((lpad_rows ? lpad_rows[r] : Lpad) - 1 - len[r]);
}

static int check_params(const dcb200_chop_params* p) {
  DCB_ARG(p != nullptr);
  DCB_ARG(p->smooth_window_size >= 1 && p->smooth_window_size < (1 << 20));
  DCB_ARG(p->min_interval_size >= 0);
  DCB_ARG(p->approved_interval_number >= 0 && p->approved_interval_number <= (1 << 20));
  DCB_ARG(p->max_process_intervals >= 0);
  DCB_ARG(p->min_read_length_after_chop >= 0);
  DCB_ARG(p->min_read_length >= 0);
  DCB_ARG(p->chop_type >= DCB200_CHOP_TERMINAL && p->chop_type <= DCB200_CHOP_ALL);
  return DCB200_OK;
}

}  // namespace dcb

using namespace dcb;

extern "C" {

const char* dcb200_last_error(void) { return g_err; }
int dcb200_version(void) { return 100; }

void dcb200_chop_params_default(dcb200_chop_params* p) {
  if (!p) return;
  p->smooth_window_size = 21;
  p->min_interval_size = 13;
  p->approved_interval_number = 20;
  p->max_process_intervals = 4;
  p->min_read_length_after_chop = 20;
  p->min_read_length = 150;
  p->chop_type = DCB200_CHOP_ALL;
  p->output_chopped_seqs = 0;
}

int dcb200_ctx_create(int device, void* stream, dcb200_ctx** out) {
  DCB_ARG(out != nullptr);
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device visible (%s); libdcb200 has no CPU fallback", cudaGetErrorString(e));
    return DCB200_ENODEV;
  }
  DCB_ARG(device >= 0 && device < ndev);
  cudaDeviceProp prop;
  DCB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libdcb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return DCB200_ENODEV;
  }
  DCB_CUDA(cudaSetDevice(device));
  dcb200_ctx* ctx = new dcb200_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (stream) {
    ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
      return DCB200_ECUDA;
    }
    ctx->owns_stream = true;
  }
  if (const char* te = getenv("DCB200_TRACE")) {  // kernel timeline tracer (tools/trace_*.py), read once per ctx
    if (!strcmp(te, "inproj")) ctx->trace_kind = TRACE_INPROJ;
    else if (!strcmp(te, "block")) ctx->trace_kind = TRACE_BLOCK;
    else if (!strcmp(te, "toeplitz")) ctx->trace_kind = TRACE_TOEPLITZ;
  }
  *out = ctx;
  return DCB200_OK;
}

int dcb200_ctx_set_option(dcb200_ctx* ctx, const char* name, int64_t value) {
  DCB_ARG(ctx && name);
  if (!strcmp(name, "fft_min_len")) {
    DCB_ARG(value >= 0 && value <= INT_MAX);
    ctx->fft_min_len = (int)value;
    return DCB200_OK;
  }
  if (!strcmp(name, "smooth_warp_kernel")) {
    DCB_ARG(value >= 0 && value <= 2);
    ctx->smooth_warp_kernel = (int)value;
    return DCB200_OK;
  }
  set_error("unknown ctx option '%s'", name);
  return DCB200_EINVAL;
}

int64_t dcb200_ctx_get_option(dcb200_ctx* ctx, const char* name) {
  if (!ctx || !name) return -1;
  if (!strcmp(name, "fft_min_len")) return ctx->fft_min_len;
  if (!strcmp(name, "smooth_warp_kernel")) return ctx->smooth_warp_kernel;
  return -1;
}

int dcb200_ctx_destroy(dcb200_ctx* ctx) {
  if (!ctx) return DCB200_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->ws) kv.second.release();
  for (auto& r : ctx->prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return DCB200_OK;
}

int dcb200_ctx_sync(dcb200_ctx* ctx) {
  DCB_ARG(ctx != nullptr);
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCB200_OK;
}

void* dcb200_ctx_stream(dcb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t dcb200_ctx_launch_count(dcb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dcb200_encode_batch(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                        const int32_t* len, int32_t R, int32_t Lpad, int32_t Lrow, uint8_t* tok, float* qual) {
  DCB_ARG(ctx && bytes && seq_off && qual_off && len && tok && qual);
  DCB_ARG(R >= 0 && Lpad > 0 && Lrow >= Lpad && Lrow % 4 == 0);
  DCB_ARG((reinterpret_cast<uintptr_t>(tok) & 3) == 0 && (reinterpret_cast<uintptr_t>(qual) & 15) == 0);
  DCB_CUDA(cudaSetDevice(ctx->device));
  return encode_device(ctx, bytes, seq_off, qual_off, len, nullptr, R, Lpad, Lrow, tok, qual);
}

int dcb200_encode_batch_rows(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                             const int32_t* len, const int32_t* lpad_rows, int32_t R, int32_t Lpad, int32_t Lrow,
                             uint8_t* tok, float* qual) {
  DCB_ARG(ctx && bytes && seq_off && qual_off && len && lpad_rows && tok && qual);
  DCB_ARG(R >= 0 && Lpad > 0 && Lrow >= Lpad && Lrow % 4 == 0);
  DCB_ARG((reinterpret_cast<uintptr_t>(tok) & 3) == 0 && (reinterpret_cast<uintptr_t>(qual) & 15) == 0);
  DCB_CUDA(cudaSetDevice(ctx->device));
  return encode_device(ctx, bytes, seq_off, qual_off, len, lpad_rows, R, Lpad, Lrow, tok, qual);
}

int dcb200_weights_create(dcb200_ctx* ctx, const char* const* names, const float* const* data, const int64_t* numel,
                          int32_t n_tensors, dcb200_weights** out) {
  DCB_ARG(ctx && names && data && numel && out && n_tensors > 0);
  DCB_CUDA(cudaSetDevice(ctx->device));
  return weights_create(ctx, names, data, numel, n_tensors, out);
}

int dcb200_weights_destroy(dcb200_weights* w) { return weights_destroy(w); }

int dcb200_forward(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* tok, const float* qual, int32_t B,
                   int32_t L, float* logits, uint8_t* labels) {
  DCB_ARG(ctx && w && tok && qual);
  DCB_ARG(B > 0 && L > 0 && L % 128 == 0 && L <= 32768);
  DCB_ARG((int64_t)B * L <= INT_MAX / 2);  // token indices are 32-bit inside the kernels
  DCB_CUDA(cudaSetDevice(ctx->device));
  return forward_device(ctx, w, tok, qual, B, L, logits, labels, 1 << 30);
}

int dcb200_forward_debug(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* tok, const float* qual, int32_t B,
                         int32_t L, float* logits, uint8_t* labels, int32_t stop_stage) {
  DCB_ARG(ctx && w && tok && qual);
  DCB_ARG(B > 0 && L > 0 && L % 128 == 0 && L <= 32768);
  DCB_ARG((int64_t)B * L <= INT_MAX / 2);
  DCB_CUDA(cudaSetDevice(ctx->device));
  DCB_CHECK(forward_device(ctx, w, tok, qual, B, L, logits, labels, stop_stage));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCB200_OK;
}

static int prof_collect(dcb200_ctx* ctx) {
  if (ctx->prof_recs.empty()) return DCB200_OK;
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ctx->prof_ms[r.kind] += ms;
      ctx->prof_cnt[r.kind] += 1;
    }
    ctx->prof_pool.push_back(r.a);
    ctx->prof_pool.push_back(r.b);
  }
  ctx->prof_recs.clear();
  return DCB200_OK;
}

int dcb200_ctx_profile(dcb200_ctx* ctx, int enable) {
  DCB_ARG(ctx != nullptr);
  DCB_CUDA(cudaSetDevice(ctx->device));
  DCB_CHECK(prof_collect(ctx));
  ctx->profiling = enable != 0;
  return DCB200_OK;
}

int dcb200_ctx_profile_read(dcb200_ctx* ctx, double* ms, int64_t* counts, int32_t n, int32_t reset) {
  DCB_ARG(ctx && ms && counts && n >= 0);
  DCB_CUDA(cudaSetDevice(ctx->device));
  DCB_CHECK(prof_collect(ctx));
  for (int i = 0; i < n; ++i) {
    ms[i] = i < K_NKINDS ? ctx->prof_ms[i] : 0.0;
    counts[i] = i < K_NKINDS ? ctx->prof_cnt[i] : 0;
  }
  if (reset)
    for (int i = 0; i < K_NKINDS; ++i) {
      ctx->prof_ms[i] = 0.0;
      ctx->prof_cnt[i] = 0;
    }
  return DCB200_OK;
}

const char* dcb200_kernel_kind_name(int32_t kind) {
  static const char* names[K_NKINDS] = {"encode", "embed_ln", "in_proj", "fft_conv", "out_proj", "fc1", "fc2",
                                        "head1", "head2", "smooth_chop", "other", "shortconv_gate", "toeplitz_conv", "mlp", "block"};
  return (kind >= 0 && kind < K_NKINDS) ? names[kind] : nullptr;
}

int dcb200_ctx_read_workspace(dcb200_ctx* ctx, const char* name, void* host_dst, int64_t bytes) {
  DCB_ARG(ctx && name && host_dst && bytes >= 0);
  DCB_CUDA(cudaSetDevice(ctx->device));
  auto it = ctx->ws.find(name);
  if (it == ctx->ws.end() || (size_t)bytes > it->second.cap) {
    set_error("workspace '%s' missing or smaller than %lld bytes", name, (long long)bytes);
    return DCB200_EINVAL;
  }
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  DCB_CUDA(cudaMemcpy(host_dst, it->second.p, (size_t)bytes, cudaMemcpyDeviceToHost));
  return DCB200_OK;
}

int dcb200_smooth_chop(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                       const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                       int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action) {
  DCB_ARG(ctx && (labels || labels_bytes == 0) && starts && lens && n_adapter && adapter_iv && n_keep && keep_iv && action);
  DCB_ARG(R >= 0 && labels_bytes >= 0);
  DCB_CHECK(check_params(p));
  DCB_CUDA(cudaSetDevice(ctx->device));
  return smooth_chop_device(ctx, labels, nullptr, labels_bytes, starts, lens, qual_lens, R, p, n_adapter, adapter_iv,
                            n_keep, keep_iv, action, nullptr);
}

int dcb200_smooth_chop_logits(dcb200_ctx* ctx, const float* logits, int64_t n_tokens, const int64_t* starts,
                              const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                              int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv,
                              uint8_t* action) {
  DCB_ARG(ctx && (logits || n_tokens == 0) && starts && lens && n_adapter && adapter_iv && n_keep && keep_iv && action);
  DCB_ARG(R >= 0 && n_tokens >= 0 && (reinterpret_cast<uintptr_t>(logits) & 7) == 0);
  DCB_CHECK(check_params(p));
  DCB_CUDA(cudaSetDevice(ctx->device));
  return smooth_chop_device(ctx, nullptr, logits, n_tokens, starts, lens, qual_lens, R, p, n_adapter, adapter_iv, n_keep,
                            keep_iv, action, nullptr);
}

int dcb200_majority_voting(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                           const int32_t* lens, int64_t R, int32_t window, int8_t* out) {
  DCB_ARG(ctx && (labels || labels_bytes == 0) && starts && lens && (out || labels_bytes == 0));
  DCB_ARG(R >= 0 && window >= 1 && window < (1 << 20));
  DCB_CUDA(cudaSetDevice(ctx->device));
  dcb200_chop_params p;
  dcb200_chop_params_default(&p);
  p.smooth_window_size = window;
  p.min_read_length = 0;
  if (labels_bytes == 0 || R == 0) return DCB200_OK;
  return smooth_chop_device(ctx, labels, nullptr, labels_bytes, starts, lens, nullptr, R, &p, nullptr, nullptr, nullptr,
                            nullptr, nullptr, out);
}

// ---- host-buffer forms ---------------------------------------------------------------------------

static int stage_in(dcb200_ctx* ctx, const char* name, const void* host, size_t bytes, void** dev) {
  dcb::DevBuf& b = ctx->buf(name);
  DCB_CHECK(b.reserve(bytes ? bytes : 1));
  if (bytes) DCB_CUDA(cudaMemcpyAsync(b.p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev = b.p;
  return DCB200_OK;
}

static int stage_out(dcb200_ctx* ctx, const char* name, size_t bytes, void** dev) {
  dcb::DevBuf& b = ctx->buf(name);
  DCB_CHECK(b.reserve(bytes ? bytes : 1));
  *dev = b.p;
  return DCB200_OK;
}

int dcb200_smooth_chop_host(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                            const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                            int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv,
                            uint8_t* action) {
  DCB_ARG(ctx && (labels || labels_bytes == 0) && starts && lens && n_adapter && adapter_iv && n_keep && keep_iv && action);
  DCB_ARG(R >= 0 && labels_bytes >= 0);
  DCB_CHECK(check_params(p));
  DCB_CUDA(cudaSetDevice(ctx->device));
  if (R == 0) return DCB200_OK;
  const int ap = p->approved_interval_number;
  void *d_lab, *d_st, *d_len, *d_ql = nullptr, *d_na, *d_ad, *d_nk, *d_kp, *d_act;
  DCB_CHECK(stage_in(ctx, "h_labels", labels, (size_t)labels_bytes, &d_lab));
  DCB_CHECK(stage_in(ctx, "h_starts", starts, (size_t)R * 8, &d_st));
  DCB_CHECK(stage_in(ctx, "h_lens", lens, (size_t)R * 4, &d_len));
  if (qual_lens) DCB_CHECK(stage_in(ctx, "h_qlens", qual_lens, (size_t)R * 4, &d_ql));
  DCB_CHECK(stage_out(ctx, "h_nad", (size_t)R * 4, &d_na));
  DCB_CHECK(stage_out(ctx, "h_ad", (size_t)R * ap * 8, &d_ad));
  DCB_CHECK(stage_out(ctx, "h_nk", (size_t)R * 4, &d_nk));
  DCB_CHECK(stage_out(ctx, "h_kp", (size_t)R * (ap + 1) * 8, &d_kp));
  DCB_CHECK(stage_out(ctx, "h_act", (size_t)R, &d_act));
  DCB_CUDA(cudaMemsetAsync(d_ad, 0, (size_t)R * ap * 8, ctx->stream));
  DCB_CUDA(cudaMemsetAsync(d_kp, 0, (size_t)R * (ap + 1) * 8, ctx->stream));
  DCB_CHECK(smooth_chop_device(ctx, (const int8_t*)d_lab, nullptr, labels_bytes, (const int64_t*)d_st, (const int32_t*)d_len,
                               (const int32_t*)d_ql, R, p, (int32_t*)d_na, (int32_t*)d_ad, (int32_t*)d_nk, (int32_t*)d_kp,
                               (uint8_t*)d_act, nullptr));
  DCB_CUDA(cudaMemcpyAsync(n_adapter, d_na, (size_t)R * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (ap) DCB_CUDA(cudaMemcpyAsync(adapter_iv, d_ad, (size_t)R * ap * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(n_keep, d_nk, (size_t)R * 4, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(keep_iv, d_kp, (size_t)R * (ap + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(action, d_act, (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCB200_OK;
}

int dcb200_majority_voting_host(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                                const int32_t* lens, int64_t R, int32_t window, int8_t* out) {
  DCB_ARG(ctx && (labels || labels_bytes == 0) && starts && lens && (out || labels_bytes == 0));
  DCB_ARG(R >= 0 && window >= 1 && window < (1 << 20));
  DCB_CUDA(cudaSetDevice(ctx->device));
  if (R == 0 || labels_bytes == 0) return DCB200_OK;
  void *d_lab, *d_st, *d_len, *d_out;
  DCB_CHECK(stage_in(ctx, "h_labels", labels, (size_t)labels_bytes, &d_lab));
  DCB_CHECK(stage_in(ctx, "h_starts", starts, (size_t)R * 8, &d_st));
  DCB_CHECK(stage_in(ctx, "h_lens", lens, (size_t)R * 4, &d_len));
  DCB_CHECK(stage_out(ctx, "h_smoothed", (size_t)labels_bytes, &d_out));
  // positions outside any read keep their input value
  DCB_CUDA(cudaMemcpyAsync(d_out, d_lab, (size_t)labels_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  DCB_CHECK(dcb200_majority_voting(ctx, (const int8_t*)d_lab, labels_bytes, (const int64_t*)d_st, (const int32_t*)d_len, R,
                                   window, (int8_t*)d_out));
  DCB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)labels_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCB200_OK;
}

}  // extern "C"

static int predict_batch_host_impl(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* bytes, int64_t n_bytes,
                                   const int64_t* seq_off, const int64_t* qual_off, const int32_t* len,
                                   const int32_t* lpad_rows, const int32_t* qual_lens, int32_t R, int32_t Lpad,
                                   const dcb200_chop_params* p, float* logits_out, uint8_t* labels_out, int32_t* n_adapter,
                                   int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action) {
  DCB_ARG(ctx && w && bytes && seq_off && qual_off && len && n_adapter && adapter_iv && n_keep && keep_iv && action);
  DCB_ARG(R > 0 && n_bytes > 0 && Lpad > 0 && Lpad <= 32768);
  DCB_CHECK(check_params(p));
  DCB_CUDA(cudaSetDevice(ctx->device));
  const int ap = p->approved_interval_number;
  const int Lrow = (Lpad + 127) / 128 * 128;  // row stride; columns >= Lpad are inert right filler (causal model)
  const size_t T = (size_t)R * Lrow;
  void *d_bytes, *d_so, *d_qo, *d_len, *d_ql = nullptr, *d_tok, *d_q, *d_lab, *d_logits = nullptr, *d_st, *d_lp = nullptr;
  void *d_na, *d_ad, *d_nk, *d_kp, *d_act;
  DCB_CHECK(stage_in(ctx, "p_bytes", bytes, (size_t)n_bytes, &d_bytes));
  DCB_CHECK(stage_in(ctx, "p_seq_off", seq_off, (size_t)R * 8, &d_so));
  DCB_CHECK(stage_in(ctx, "p_qual_off", qual_off, (size_t)R * 8, &d_qo));
  DCB_CHECK(stage_in(ctx, "p_len", len, (size_t)R * 4, &d_len));
  if (qual_lens) DCB_CHECK(stage_in(ctx, "p_qlens", qual_lens, (size_t)R * 4, &d_ql));
  if (lpad_rows) DCB_CHECK(stage_in(ctx, "p_lpad_rows", lpad_rows, (size_t)R * 4, &d_lp));
  // a bad offset from across the plain-pointer ABI must not become an out-of-bounds device read
  DCB_ARG((int64_t)R * Lrow <= INT_MAX / 2);
  for (int r = 0; r < R; ++r) {
    DCB_ARG(len[r] >= 0 && len[r] + 1 <= (lpad_rows ? lpad_rows[r] : Lpad));
    DCB_ARG(!lpad_rows || lpad_rows[r] <= Lpad);
    DCB_ARG(seq_off[r] >= 0 && seq_off[r] + len[r] <= n_bytes && qual_off[r] >= 0 && qual_off[r] + len[r] <= n_bytes);
  }
  // label-row starts: read r occupies columns [Lpad-len-1, Lpad-1) of row r (left pad, SEP last); computed on the device
  // from the lengths that are uploaded anyway (no host staging vector, no extra synchronisation)
  DCB_CHECK(stage_out(ctx, "p_starts", (size_t)R * 8, &d_st));
  row_starts_kernel<<<(R + 255) / 256, 256, 0, ctx->stream>>>((const int32_t*)d_len, (const int32_t*)d_lp, R, Lrow, Lpad,
                                                              (int64_t*)d_st);
  DCB_LAUNCH_CHECK(ctx);
  DCB_CHECK(stage_out(ctx, "p_tok", T, &d_tok));
  DCB_CHECK(stage_out(ctx, "p_qual", T * 4, &d_q));
  DCB_CHECK(stage_out(ctx, "p_labels", T, &d_lab));
  if (logits_out) DCB_CHECK(stage_out(ctx, "p_logits", T * 8, &d_logits));
  DCB_CHECK(stage_out(ctx, "h_nad", (size_t)R * 4, &d_na));
  DCB_CHECK(stage_out(ctx, "h_ad", (size_t)R * ap * 8, &d_ad));
  DCB_CHECK(stage_out(ctx, "h_nk", (size_t)R * 4, &d_nk));
  DCB_CHECK(stage_out(ctx, "h_kp", (size_t)R * (ap + 1) * 8, &d_kp));
  DCB_CHECK(stage_out(ctx, "h_act", (size_t)R, &d_act));
  DCB_CUDA(cudaMemsetAsync(d_ad, 0, (size_t)R * ap * 8, ctx->stream));
  DCB_CUDA(cudaMemsetAsync(d_kp, 0, (size_t)R * (ap + 1) * 8, ctx->stream));
  DCB_CHECK(encode_device(ctx, (const uint8_t*)d_bytes, (const int64_t*)d_so, (const int64_t*)d_qo, (const int32_t*)d_len,
                          (const int32_t*)d_lp, R,
                          Lpad, Lrow, (uint8_t*)d_tok, (float*)d_q));
  DCB_CHECK(forward_device(ctx, w, (const uint8_t*)d_tok, (const float*)d_q, R, Lrow, (float*)d_logits, (uint8_t*)d_lab, 1 << 30));
  DCB_CHECK(smooth_chop_device(ctx, (const int8_t*)d_lab, nullptr, (int64_t)T, (const int64_t*)d_st, (const int32_t*)d_len,
                               (const int32_t*)d_ql, R, p, (int32_t*)d_na, (int32_t*)d_ad, (int32_t*)d_nk, (int32_t*)d_kp,
                               (uint8_t*)d_act, nullptr));
  if (logits_out)
    DCB_CUDA(cudaMemcpy2DAsync(logits_out, (size_t)Lpad * 8, d_logits, (size_t)Lrow * 8, (size_t)Lpad * 8, R,
                               cudaMemcpyDeviceToHost, ctx->stream));
  if (labels_out)
    DCB_CUDA(cudaMemcpy2DAsync(labels_out, (size_t)Lpad, d_lab, (size_t)Lrow, (size_t)Lpad, R, cudaMemcpyDeviceToHost,
                               ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(n_adapter, d_na, (size_t)R * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (ap) DCB_CUDA(cudaMemcpyAsync(adapter_iv, d_ad, (size_t)R * ap * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(n_keep, d_nk, (size_t)R * 4, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(keep_iv, d_kp, (size_t)R * (ap + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaMemcpyAsync(action, d_act, (size_t)R, cudaMemcpyDeviceToHost, ctx->stream));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  return DCB200_OK;
}

extern "C" {

int dcb200_predict_batch_host(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* bytes, int64_t n_bytes,
                              const int64_t* seq_off, const int64_t* qual_off, const int32_t* len,
                              const int32_t* qual_lens, int32_t R, int32_t Lpad, const dcb200_chop_params* p,
                              float* logits_out, uint8_t* labels_out, int32_t* n_adapter, int32_t* adapter_iv,
                              int32_t* n_keep, int32_t* keep_iv, uint8_t* action) {
  return predict_batch_host_impl(ctx, w, bytes, n_bytes, seq_off, qual_off, len, nullptr, qual_lens, R, Lpad, p, logits_out,
                                 labels_out, n_adapter, adapter_iv, n_keep, keep_iv, action);
}

int dcb200_predict_batch_host_rows(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* bytes, int64_t n_bytes,
                                   const int64_t* seq_off, const int64_t* qual_off, const int32_t* len,
                                   const int32_t* lpad_rows, const int32_t* qual_lens, int32_t R, int32_t Lpad,
                                   const dcb200_chop_params* p, float* logits_out, uint8_t* labels_out, int32_t* n_adapter,
                                   int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action) {
  DCB_ARG(lpad_rows != nullptr);
  return predict_batch_host_impl(ctx, w, bytes, n_bytes, seq_off, qual_off, len, lpad_rows, qual_lens, R, Lpad, p, logits_out,
                                 labels_out, n_adapter, adapter_iv, n_keep, keep_iv, action);
}

}  // extern "C"
