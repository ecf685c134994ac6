// Weights + forward orchestration of the HyenaDNA-small-32k token classifier:
//   embedding -> 4 x [LN1 -> in_proj -> short conv/gate/long conv/gate -> out_proj + res -> LN2 -> MLP + res]
//   -> ln_f -> head (deepchopper/models/llm/head.py:94-102) -> logits / labels
// Reference entry: deepchopper/models/basic_module.py:90-100 -> deepchopper/models/llm/hyena.py:29-41.
#include "common.cuh"
#include "block.h"
#include "gemm.h"
#include "head1.h"
#include "inproj.h"
#include "lconv.h"
#include "toeplitz.h"

#include <math.h>
#include <string.h>

#include <mutex>
#include <string>

namespace dcb {

constexpr int kD = 256;
constexpr int kLayers = 4;
constexpr int kInner = 1024;
constexpr int kFilterOrder = 64;
constexpr int kEmbDim = 5;
constexpr int kVocab = 16;

struct LayerW {
  __nv_bfloat16 *w_in = nullptr, *w_out = nullptr, *w_fc1 = nullptr, *w_fc2 = nullptr;
  float *b_in = nullptr, *b_out = nullptr, *b_fc1 = nullptr, *b_fc2 = nullptr;
  float *short_w = nullptr, *short_b = nullptr, *filt_D = nullptr;
  float* k = nullptr;  // [256][Lmax] implicit filter, evaluated once
  __nv_bfloat16* toep = nullptr;  // Toeplitz core-matrix table of k' (toeplitz.cu), built on first use
  float4* lc_K = nullptr;         // filter spectra of the blocked FFT convolution (lconv.cu), built on first use
  CUtensorMap tm_in_mc;
  std::vector<float> hb_fc1, hb_fc2, hb_out;  // host copies: passed to the block kernel as constant-bank parameters
  CUtensorMap tm_w1u, tm_w2u, tm_wou;  // per-CTA halves of the weight tiles for the block kernel (128-row boxes)
};

}  // namespace dcb

struct dcb200_weights {
  int device = 0;
  int Lmax = 0;
  float* emb = nullptr;  // [16][256]
  __nv_bfloat16* emb_u = nullptr;  // [16][256]: LayerNorm (norm1 of layer 0, affine folded away) of every embedding row
  dcb::LayerW layer[dcb::kLayers];
  __nv_bfloat16 *wh1 = nullptr, *wh2 = nullptr;
  float *bh1 = nullptr, *bh2 = nullptr, *w3 = nullptr, *b3 = nullptr;
  CUtensorMap tm_h1, tm_h2;  // tm_h1: boxes of 128 rows (head1.cu), tm_h2: 256 rows (gemm.cu)
  int toep_cap = 0;           // read length the Toeplitz tables were built for
  float2* lc_tw = nullptr;    // twiddle tables of the blocked FFT convolution
  int lc_nbK = 0;             // block distances the filter spectra cover
  std::mutex lazy_mu;         // guards the lazily built tables above (several ctxs / threads may share the weights)
  std::vector<void*> allocs;
};

namespace dcb {

// ---- small kernels --------------------------------------------------------------------------------

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

// Implicit filter MLP (modeling_hyena.py HyenaFilter.filter): one thread per position t.
//   h = Lin6( sin(f5 * Lin4( sin(f3 * Lin2( sin(f1 * Lin0(z_t)) )) )) ),  k[c][t] = h_c * (exp(-t |delta_c|) + 0.05)
struct FilterW {
  const float *z, *t;              // [Lmax,5], [Lmax]
  const float *w0, *b0, *f1;       // [64,5], [64], [64]
  const float *w2, *b2, *f3;       // [64,64], [64], [64]
  const float *w4, *b4, *f5;       // [64,64], [64], [64]
  const float *w6;                 // [256,64]
  const float *deltas;             // [256]
};

__global__ void __launch_bounds__(128) implicit_filter_kernel(FilterW w, int Lmax, float* __restrict__ k) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Lmax) return;
  float h1[kFilterOrder], h2[kFilterOrder];
  float zt[kEmbDim];
  for (int i = 0; i < kEmbDim; ++i) zt[i] = w.z[(size_t)t * kEmbDim + i];
  for (int j = 0; j < kFilterOrder; ++j) {
    float a = w.b0[j];
    for (int i = 0; i < kEmbDim; ++i) a = fmaf(w.w0[j * kEmbDim + i], zt[i], a);
    h1[j] = sinf(w.f1[j] * a);
  }
  for (int j = 0; j < kFilterOrder; ++j) {
    float a = w.b2[j];
    for (int i = 0; i < kFilterOrder; ++i) a = fmaf(w.w2[j * kFilterOrder + i], h1[i], a);
    h2[j] = sinf(w.f3[j] * a);
  }
  for (int j = 0; j < kFilterOrder; ++j) {
    float a = w.b4[j];
    for (int i = 0; i < kFilterOrder; ++i) a = fmaf(w.w4[j * kFilterOrder + i], h2[i], a);
    h1[j] = sinf(w.f5[j] * a);
  }
  const float tt = w.t[t];
  for (int c = 0; c < kD; ++c) {
    float a = 0.f;
    for (int i = 0; i < kFilterOrder; ++i) a = fmaf(w.w6[c * kFilterOrder + i], h1[i], a);
    k[(size_t)c * Lmax + t] = a * (expf(-tt * fabsf(w.deltas[c])) + 0.05f);
  }
}

// Embedding + LayerNorm1 of layer 0.  There are 16 token ids, so LayerNorm(embedding row) has 16 possible values: they are
// computed once per weight set (one warp per id, 8 features per lane) ...
__global__ void __launch_bounds__(32) embed_table_kernel(const float* __restrict__ emb, __nv_bfloat16* __restrict__ emb_u) {
  const int lane = threadIdx.x & 31;
  const int id = blockIdx.x;
  const float4* e4 = reinterpret_cast<const float4*>(emb + id * kD + lane * 8);
  const float4 a = __ldg(e4), c = __ldg(e4 + 1);
  float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s * (1.0f / kD);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) v = fmaf(x[i] - mean, x[i] - mean, v);
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const float rstd = rsqrtf(v * (1.0f / kD) + 1e-5f);
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float y0 = (x[2 * i] - mean) * rstd, y1 = (x[2 * i + 1] - mean) * rstd;  // affine part folded into in_linear
    __nv_bfloat162 r = __floats2bfloat162_rn(y0, y1);
    o[i] = *reinterpret_cast<uint32_t*>(&r);
  }
  *reinterpret_cast<uint4*>(emb_u + id * kD + lane * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ... and the per-batch kernel is a gather-copy: one warp per token writes the fp32 residual row (1 KB: two 16-byte
// stores per lane, each instruction one contiguous 512-byte run) and the bf16 LayerNorm row (512 B: one store per lane)
// from the two tables (24 KB, L1-resident).  (Writing 32 contiguous bytes per lane as two 16-byte stores made every store
// instruction cover half of each 32-byte sector: 0.69 of the HBM roofline.)
__global__ void __launch_bounds__(256) embed_ln_kernel(const uint8_t* __restrict__ tok, const float* __restrict__ emb,
                                                       const __nv_bfloat16* __restrict__ emb_u, int T,
                                                       float* __restrict__ h, __nv_bfloat16* __restrict__ u) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int t = warp; t < T; t += nwarps) {
    int id = tok[t];
    id = id < kVocab ? id : kVocab - 1;
    const float4* e4 = reinterpret_cast<const float4*>(emb + id * kD);
    const float4 a = __ldg(e4 + lane), c = __ldg(e4 + 32 + lane);
    const uint4 y = __ldg(reinterpret_cast<const uint4*>(emb_u + id * kD) + lane);
    float4* h4 = reinterpret_cast<float4*>(h + (size_t)t * kD);
    h4[lane] = a;
    h4[32 + lane] = c;
    reinterpret_cast<uint4*>(u + (size_t)t * kD)[lane] = y;
  }
}

// ---- weights --------------------------------------------------------------------------------------

struct StateDict {
  const char* const* names;
  const float* const* data;
  const int64_t* numel;
  int n;
  // find a tensor whose name ends with `suffix` (prefixes such as "net.backbone.backbone." are ignored)
  int find(const std::string& suffix) const {
    for (int i = 0; i < n; ++i) {
      const size_t ln = strlen(names[i]);
      if (ln >= suffix.size() && memcmp(names[i] + ln - suffix.size(), suffix.data(), suffix.size()) == 0) {
        if (ln == suffix.size() || names[i][ln - suffix.size() - 1] == '.') return i;
      }
    }
    return -1;
  }
};

static int upload_f32(dcb200_ctx* ctx, dcb200_weights* w, const StateDict& sd, const std::string& key, int64_t numel,
                      float** out) {
  const int i = sd.find(key);
  if (i < 0) {
    set_error("state dict has no tensor ending in '%s'", key.c_str());
    return DCB200_EWEIGHT;
  }
  if (sd.numel[i] != numel) {
    set_error("tensor '%s' has %lld elements, expected %lld", sd.names[i], (long long)sd.numel[i], (long long)numel);
    return DCB200_EWEIGHT;
  }
  void* d = nullptr;
  DCB_CUDA(cudaMalloc(&d, (size_t)numel * 4));
  w->allocs.push_back(d);
  DCB_CUDA(cudaMemcpyAsync(d, sd.data[i], (size_t)numel * 4, cudaMemcpyHostToDevice, ctx->stream));
  *out = static_cast<float*>(d);
  return DCB200_OK;
}

static int host_copy(const StateDict& sd, const std::string& key, int64_t numel, std::vector<float>& out) {
  const int i = sd.find(key);
  if (i < 0 || sd.numel[i] != numel) {
    set_error("state dict has no tensor ending in '%s' with %lld elements", key.c_str(), (long long)numel);
    return DCB200_EWEIGHT;
  }
  out.assign(sd.data[i], sd.data[i] + numel);
  return DCB200_OK;
}

static int upload_bf16(dcb200_ctx* ctx, dcb200_weights* w, const StateDict& sd, const std::string& key, int64_t numel,
                       __nv_bfloat16** out) {
  float* tmp = nullptr;
  DCB_CHECK(upload_f32(ctx, w, sd, key, numel, &tmp));
  void* d = nullptr;
  DCB_CUDA(cudaMalloc(&d, (size_t)numel * 2));
  w->allocs.push_back(d);
  f32_to_bf16_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, ctx->stream>>>(tmp, static_cast<__nv_bfloat16*>(d), (size_t)numel);
  DCB_LAUNCH_CHECK(ctx);
  *out = static_cast<__nv_bfloat16*>(d);
  return DCB200_OK;
}

static int upload_f32_host(dcb200_ctx* ctx, dcb200_weights* w, const std::vector<float>& h, float** out) {
  void* d = nullptr;
  DCB_CUDA(cudaMalloc(&d, h.size() * 4));
  w->allocs.push_back(d);
  DCB_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));  // `h` is a temporary of the caller
  *out = static_cast<float*>(d);
  return DCB200_OK;
}

// A LayerNorm's affine part followed by a Linear is a Linear: W (z g + b) + c = (W diag(g)) z + (c + W b).  Every
// LayerNorm of the model feeds exactly one Linear (norm1 -> in_linear, norm2 -> fc1, ln_f -> head.linear1), so the
// kernels only normalise: 640 fewer vector elements per token row in the block kernel's epilogues.  The fold is done in
// fp64 on the host; the folded weight is then rounded to bf16 like any other weight.
static int upload_folded_linear(dcb200_ctx* ctx, dcb200_weights* w, const StateDict& sd, const std::string& wkey,
                                const std::string& bkey, const std::string& gkey, const std::string& betakey, int out_f,
                                int in_f, __nv_bfloat16** w_out, float** b_out, std::vector<float>* b_host) {
  std::vector<float> W, c, g, beta;
  DCB_CHECK(host_copy(sd, wkey, (int64_t)out_f * in_f, W));
  DCB_CHECK(host_copy(sd, bkey, out_f, c));
  DCB_CHECK(host_copy(sd, gkey, in_f, g));
  DCB_CHECK(host_copy(sd, betakey, in_f, beta));
  for (int o = 0; o < out_f; ++o) {
    double acc = c[o];
    float* row = W.data() + (size_t)o * in_f;
    for (int k = 0; k < in_f; ++k) {
      acc += (double)row[k] * (double)beta[k];
      row[k] = (float)((double)row[k] * (double)g[k]);
    }
    c[o] = (float)acc;
  }
  float* tmp = nullptr;
  DCB_CHECK(upload_f32_host(ctx, w, W, &tmp));
  void* d = nullptr;
  DCB_CUDA(cudaMalloc(&d, W.size() * 2));
  w->allocs.push_back(d);
  f32_to_bf16_kernel<<<(unsigned)((W.size() + 255) / 256), 256, 0, ctx->stream>>>(tmp, static_cast<__nv_bfloat16*>(d), W.size());
  DCB_LAUNCH_CHECK(ctx);
  *w_out = static_cast<__nv_bfloat16*>(d);
  DCB_CHECK(upload_f32_host(ctx, w, c, b_out));
  if (b_host) *b_host = c;
  return DCB200_OK;
}

int weights_destroy(dcb200_weights* w) {
  if (!w) return DCB200_OK;
  cudaSetDevice(w->device);
  cudaDeviceSynchronize();
  for (void* p : w->allocs) cudaFree(p);
  delete w;
  return DCB200_OK;
}

static int weights_fill(dcb200_ctx* ctx, dcb200_weights* w, const StateDict& sd) {
  DCB_CHECK(upload_f32(ctx, w, sd, "embeddings.word_embeddings.weight", kVocab * kD, &w->emb));
  {
    void* eu = nullptr;
    DCB_CUDA(cudaMalloc(&eu, (size_t)kVocab * kD * sizeof(__nv_bfloat16)));
    w->allocs.push_back(eu);
    w->emb_u = static_cast<__nv_bfloat16*>(eu);
    embed_table_kernel<<<kVocab, 32, 0, ctx->stream>>>(w->emb, w->emb_u);
    DCB_LAUNCH_CHECK(ctx);
  }
  // positional table length (= max_seq_len of the checkpoint)
  const int iz = sd.find("layers.0.mixer.filter_fn.pos_emb.z");
  if (iz < 0 || sd.numel[iz] % kEmbDim != 0) {
    set_error("state dict has no usable 'layers.0.mixer.filter_fn.pos_emb.z'");
    return DCB200_EWEIGHT;
  }
  w->Lmax = (int)(sd.numel[iz] / kEmbDim);
  for (int l = 0; l < kLayers; ++l) {
    LayerW& lw = w->layer[l];
    const std::string p = "layers." + std::to_string(l) + ".";
    DCB_CHECK(upload_folded_linear(ctx, w, sd, p + "mixer.in_linear.weight", p + "mixer.in_linear.bias", p + "norm1.weight",
                                   p + "norm1.bias", 3 * kD, kD, &lw.w_in, &lw.b_in, nullptr));
    DCB_CHECK(upload_bf16(ctx, w, sd, p + "mixer.out_linear.weight", kD * kD, &lw.w_out));
    DCB_CHECK(upload_f32(ctx, w, sd, p + "mixer.out_linear.bias", kD, &lw.b_out));
    DCB_CHECK(upload_f32(ctx, w, sd, p + "mixer.short_filter.weight", 3 * kD * 3, &lw.short_w));
    DCB_CHECK(upload_f32(ctx, w, sd, p + "mixer.short_filter.bias", 3 * kD, &lw.short_b));
    DCB_CHECK(upload_f32(ctx, w, sd, p + "mixer.filter_fn.bias", kD, &lw.filt_D));
    // (the LayerNorms' affine parts are folded into the Linear that follows each of them: the kernels only normalise)
    DCB_CHECK(upload_folded_linear(ctx, w, sd, p + "mlp.fc1.weight", p + "mlp.fc1.bias", p + "norm2.weight",
                                   p + "norm2.bias", kInner, kD, &lw.w_fc1, &lw.b_fc1, &lw.hb_fc1));
    DCB_CHECK(upload_bf16(ctx, w, sd, p + "mlp.fc2.weight", kD * kInner, &lw.w_fc2));
    DCB_CHECK(upload_f32(ctx, w, sd, p + "mlp.fc2.bias", kD, &lw.b_fc2));
    DCB_CHECK(host_copy(sd, p + "mlp.fc2.bias", kD, lw.hb_fc2));
    DCB_CHECK(host_copy(sd, p + "mixer.out_linear.bias", kD, lw.hb_out));
    // implicit filter, evaluated once for the whole positional table
    FilterW fw;
    float *z, *t, *w0, *b0, *f1, *w2, *b2, *f3, *w4, *b4, *f5, *w6, *dl;
    const std::string f = p + "mixer.filter_fn.";
    DCB_CHECK(upload_f32(ctx, w, sd, f + "pos_emb.z", (int64_t)w->Lmax * kEmbDim, &z));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "pos_emb.t", w->Lmax, &t));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.0.weight", kFilterOrder * kEmbDim, &w0));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.0.bias", kFilterOrder, &b0));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.1.freq", kFilterOrder, &f1));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.2.weight", kFilterOrder * kFilterOrder, &w2));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.2.bias", kFilterOrder, &b2));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.3.freq", kFilterOrder, &f3));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.4.weight", kFilterOrder * kFilterOrder, &w4));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.4.bias", kFilterOrder, &b4));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.5.freq", kFilterOrder, &f5));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "implicit_filter.6.weight", kD * kFilterOrder, &w6));
    DCB_CHECK(upload_f32(ctx, w, sd, f + "modulation.deltas", kD, &dl));
    fw.z = z; fw.t = t; fw.w0 = w0; fw.b0 = b0; fw.f1 = f1; fw.w2 = w2; fw.b2 = b2; fw.f3 = f3;
    fw.w4 = w4; fw.b4 = b4; fw.f5 = f5; fw.w6 = w6; fw.deltas = dl;
    void* kbuf = nullptr;
    DCB_CUDA(cudaMalloc(&kbuf, (size_t)kD * w->Lmax * 4));
    w->allocs.push_back(kbuf);
    lw.k = static_cast<float*>(kbuf);
    implicit_filter_kernel<<<(w->Lmax + 127) / 128, 128, 0, ctx->stream>>>(fw, w->Lmax, lw.k);
    DCB_LAUNCH_CHECK(ctx);
    DCB_CHECK(make_tmap_2d(&lw.tm_in_mc, lw.w_in, 3 * kD, kD, 64));  // inproj_conv: half boxes, multicast across a cluster of 2
    DCB_CHECK(make_tmap_2d(&lw.tm_w1u, lw.w_fc1, kInner, kD, 128));
    DCB_CHECK(make_tmap_2d(&lw.tm_w2u, lw.w_fc2, kD, kInner, 128));
    DCB_CHECK(make_tmap_2d(&lw.tm_wou, lw.w_out, kD, kD, 128));
  }
  DCB_CHECK(upload_folded_linear(ctx, w, sd, "head.linear1.weight", "head.linear1.bias", "ln_f.weight", "ln_f.bias", kInner,
                                 kD, &w->wh1, &w->bh1, nullptr));
  DCB_CHECK(upload_bf16(ctx, w, sd, "head.linear2.weight", kInner * kInner, &w->wh2));
  DCB_CHECK(upload_f32(ctx, w, sd, "head.linear2.bias", kInner, &w->bh2));
  DCB_CHECK(upload_f32(ctx, w, sd, "head.linear3.weight", 2 * kInner, &w->w3));
  DCB_CHECK(upload_f32(ctx, w, sd, "head.linear3.bias", 2, &w->b3));
  DCB_CHECK(make_tmap_2d(&w->tm_h1, w->wh1, kInner, kD, 128));
  DCB_CHECK(make_tmap_2d(&w->tm_h2, w->wh2, kInner, kInner, 256));
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));  // host staging buffers of the caller may now be released
  return DCB200_OK;
}

int weights_create(dcb200_ctx* ctx, const char* const* names, const float* const* data, const int64_t* numel,
                   int32_t n, dcb200_weights** out) {
  *out = nullptr;
  dcb200_weights* w = new dcb200_weights();
  w->device = ctx->device;
  StateDict sd{names, data, numel, n};
  const int rc = weights_fill(ctx, w, sd);
  if (rc != DCB200_OK) {
    weights_destroy(w);
    return rc;
  }
  *out = w;
  return DCB200_OK;
}


// Toeplitz core-matrix tables of all layers, built on first use for reads up to 8192 tokens (35 MB per layer); a longer
// batch that still takes the Toeplitz kernel (ctx option fft_min_len raised) rebuilds them for the model's 32768.
// The tables belong to the shared weights: the build is serialised, finished before anyone launches against it
// (another ctx / stream may share the weights), and a superseded smaller table is released.
static int ensure_toeplitz(dcb200_ctx* ctx, dcb200_weights* w, int L) {
  std::lock_guard<std::mutex> lock(w->lazy_mu);
  if (w->toep_cap >= L) return DCB200_OK;
  const int cap = L <= 8192 ? 8192 : kToepMaxL;
  DCB_CUDA(cudaDeviceSynchronize());  // nobody may still be reading a table that is about to be replaced
  for (int l = 0; l < kLayers; ++l) {
    void* t = nullptr;
    DCB_CUDA(cudaMalloc(&t, toeplitz_table_bytes(cap)));
    DCB_CHECK(launch_toeplitz_table(ctx, w->layer[l].k, w->Lmax, w->Lmax, w->layer[l].filt_D, cap,
                                    static_cast<__nv_bfloat16*>(t)));
    if (void* old = w->layer[l].toep) {
      for (auto& a : w->allocs)
        if (a == old) a = t;
      cudaFree(old);
    } else {
      w->allocs.push_back(t);
    }
    w->layer[l].toep = static_cast<__nv_bfloat16*>(t);
  }
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  w->toep_cap = cap;
  return DCB200_OK;
}

// Twiddles + filter spectra of the blocked FFT convolution: all four block distances (the model's 32768 tokens) at the
// first long batch, 64 KB per channel and distance (67 MB per layer).  Same sharing rules as ensure_toeplitz.
static int ensure_lconv(dcb200_ctx* ctx, dcb200_weights* w, int L) {
  std::lock_guard<std::mutex> lock(w->lazy_mu);
  const int need = lconv_blocks_for(L);
  if (w->lc_nbK >= need) return DCB200_OK;
  const int nbK = lconv_blocks_for(lconv_max_len());
  if (!w->lc_tw) {
    void* t = nullptr;
    DCB_CUDA(cudaMalloc(&t, lconv_twiddle_bytes()));
    w->allocs.push_back(t);
    w->lc_tw = static_cast<float2*>(t);
    DCB_CHECK(launch_lconv_twiddles(ctx, w->lc_tw));
  }
  for (int l = 0; l < kLayers; ++l) {
    void* kf = nullptr;
    DCB_CUDA(cudaMalloc(&kf, lconv_spectrum_bytes(nbK)));
    w->allocs.push_back(kf);
    w->layer[l].lc_K = static_cast<float4*>(kf);
    DCB_CHECK(launch_lconv_filter_spectrum(ctx, w->layer[l].k, w->Lmax, w->Lmax, w->layer[l].filt_D, w->lc_tw, nbK,
                                           w->layer[l].lc_K));
  }
  DCB_CUDA(cudaStreamSynchronize(ctx->stream));
  w->lc_nbK = nbK;
  return DCB200_OK;
}

// Which long-convolution kernel a batch of B rows of padded length L takes.  Measured on B200 (profiles/r02_summary.md):
// the Toeplitz kernel works on M-tiles of 128 batch rows and its MMA work grows with L: a full row tile costs ~0.037 ns
// per token^2 and layer (0.29 ns per token per 1024 tokens of L).  A row tile with at most 64 real rows (the reference's
// batch 16 is one such tile) costs half of that, 0.0185 ns per token^2: its MMAs run on TMA zero fill, the board leaves
// its 1 kW power cap and the tensor pipe runs at full clock.  The FFT kernel works per (row, channel) sequence and costs a
// fixed time per 8192-token block -- 1.57 / 1.76 / 1.95 / 2.06 ns per token slot for reads of 1 / 2 / 3 / 4 blocks --
// whether the last block is full or not.  With full row tiles the FFT wins from ~6.7 k tokens up to 8192, loses again
// while a second block is mostly empty (8.3 k - 9.9 k) and wins everywhere above; with 16 rows it wins from ~3.4 k tokens.
// An explicit "fft_min_len" option is a plain threshold on L.
static bool use_fft_conv(const dcb200_ctx* ctx, int B, int L) {
  if (L > lconv_max_len()) return false;
  if (ctx->fft_min_len != kDefaultFftMinLen) return L >= ctx->fft_min_len;
  static const double kSlotNs[4] = {1.57, 1.76, 1.95, 2.06};
  const int nb = lconv_blocks_for(L);
  const double fft_ns = (double)B * nb * 8192.0 * kSlotNs[nb - 1];
  const int rem = B % 128;
  const double tiles = B / 128 + (rem > 64 ? 1.0 : (rem > 0 ? 0.5 : 0.0));
  const double toeplitz_ns = tiles * 128.0 * 0.29e-3 * L * L;
  return fft_ns < toeplitz_ns;
}

int forward_device(dcb200_ctx* ctx, const dcb200_weights* wc, const uint8_t* tok, const float* qual, int32_t B, int32_t L,
                   float* logits, uint8_t* labels, int stop_stage) {
  int stage = 0;
#define DCB_STAGE_DONE()               \
  do {                                 \
    if (stage++ >= stop_stage) return DCB200_OK; \
  } while (0)
  dcb200_weights* w = const_cast<dcb200_weights*>(wc);  // lazily built convolution tables
  if (w->device != ctx->device) {
    set_error("weights live on device %d, ctx on %d", w->device, ctx->device);
    return DCB200_EINVAL;
  }
  // Long convolution: tensor-core Toeplitz GEMMs (MMA work grows with L) below the measured crossover, the blocked
  // shared-memory FFT (O(L log L), fp32) above it (use_fft_conv).  ctx option "fft_min_len" overrides (the tests run both
  // kernels at the same sizes).
  const bool fft = use_fft_conv(ctx, B, L);
  if (fft) DCB_CHECK(ensure_lconv(ctx, w, L));
  else DCB_CHECK(ensure_toeplitz(ctx, w, L));
  const size_t T = (size_t)B * L;
  DevBuf& bhA = ctx->buf("act_hA");
  DevBuf& bu = ctx->buf("act_u");
  DevBuf& by = ctx->buf("act_y");
  DevBuf& bg = ctx->buf("act_g");
  DevBuf& bvv = ctx->buf("act_vv");
  DevBuf& bgt = ctx->buf("act_gate");
  DCB_CHECK(bhA.reserve(T * kD * 4));
  DCB_CHECK(bu.reserve(T * kD * 2));
  DCB_CHECK(by.reserve(T * kD * 2));
  DCB_CHECK(bg.reserve(T * kInner * 2));
  DCB_CHECK(bvv.reserve(T * kD * 2));
  DCB_CHECK(bgt.reserve(T * kD * 2));
  float* hA = bhA.as<float>();
  __nv_bfloat16* u = bu.as<__nv_bfloat16>();
  __nv_bfloat16* y = by.as<__nv_bfloat16>();
  __nv_bfloat16* g = bg.as<__nv_bfloat16>();
  __nv_bfloat16* vv = bvv.as<__nv_bfloat16>();
  __nv_bfloat16* gate = bgt.as<__nv_bfloat16>();

  CUtensorMap tm_vv, tm_gate, tm_yr, tm_vv_st, tm_gate_st, tm_u144, tm_u, tm_y, tm_g, tm_hA;
  if (!fft) {
    DCB_CHECK(make_tmap_3d_rows(&tm_vv, vv, B, kD, L));
    DCB_CHECK(make_tmap_3d_rows(&tm_gate, gate, B, kD, L));
    DCB_CHECK(make_tmap_3d_rows(&tm_yr, y, B, kD, L));
  }
  DCB_CHECK(make_tmap_3d_chbox(&tm_vv_st, vv, B, kD, L));
  DCB_CHECK(make_tmap_3d_chbox(&tm_gate_st, gate, B, kD, L));
  DCB_CHECK(make_tmap_2d(&tm_u144, u, T, kD, 144));
  DCB_CHECK(make_tmap_2d_f32(&tm_hA, hA, T, kD));
  DCB_CHECK(make_tmap_2d(&tm_u, u, T, kD, 128));
  DCB_CHECK(make_tmap_3d_cm(&tm_y, y, B, kD, L));
  DCB_CHECK(make_tmap_2d(&tm_g, g, T, kInner, 128));

  auto trace_buf = [&](int kind, int layer, long long** out) -> int {
    *out = nullptr;
    if (ctx->trace_kind != kind || layer != 0) return DCB200_OK;
    DevBuf& bt = ctx->buf("trace");
    DCB_CHECK(bt.reserve(3 * 4096 * 2 * 8));
    DCB_CUDA(cudaMemsetAsync(bt.p, 0, 3 * 4096 * 2 * 8, ctx->stream));
    *out = bt.as<long long>();
    return DCB200_OK;
  };

  {
    int blocks = (int)((T + 7) / 8);
    const int cap = ctx->sm_count * 8 * 4;
    if (blocks > cap) blocks = cap;
    ProfScope prof(ctx, K_EMBED);
    embed_ln_kernel<<<blocks, 256, 0, ctx->stream>>>(tok, w->emb, w->emb_u, (int)T, hA, u);
    DCB_LAUNCH_CHECK(ctx);
  }
  DCB_STAGE_DONE();
  for (int l = 0; l < kLayers; ++l) {
    LayerW& lw = w->layer[l];
    {
      // in_proj + short conv + first gate in one kernel: z never exists
      InprojParams ip;
      ip.num_tiles = (int)(T / 128);
      ip.L = L;
      ip.b_in = lw.b_in;
      ip.short_w = lw.short_w;
      ip.short_b = lw.short_b;
      DCB_CHECK(trace_buf(TRACE_INPROJ, l, &ip.trace));
      DCB_CHECK(launch_inproj_conv(ctx, tm_u144, lw.tm_in_mc, tm_vv_st, tm_gate_st, ip));
    }
    DCB_STAGE_DONE();
    if (fft) DCB_CHECK(launch_lconv(ctx, vv, gate, y, lw.lc_K, w->lc_nbK, w->lc_tw, B, L));
    else DCB_CHECK(launch_toeplitz_conv(ctx, lw.toep, w->toep_cap, tm_vv, tm_gate, tm_yr, B, L));
    DCB_STAGE_DONE();
    {
      // out_proj + LN2 + MLP + residual + next LN in one kernel: h1 and m never leave the SM
      static thread_local BlockParams bp;  // 11 KB: keep it off the stack
      bp.num_pairs = (int)((T / 128 + 1) / 2);
      bp.T = (int)T;
      bp.L = L;
      DCB_CHECK(trace_buf(TRACE_BLOCK, l, &bp.trace));
      memcpy(bp.bo, lw.hb_out.data(), sizeof(bp.bo));
      memcpy(bp.b1, lw.hb_fc1.data(), sizeof(bp.b1));
      memcpy(bp.b2, lw.hb_fc2.data(), sizeof(bp.b2));
      // residual stream updated in place: every tile reads its rows before it writes them
      DCB_CHECK(launch_block(ctx, tm_y, lw.tm_wou, lw.tm_w1u, lw.tm_w2u, tm_hA, tm_hA, tm_u, bp));
    }
    DCB_STAGE_DONE();
    DCB_STAGE_DONE();
    DCB_STAGE_DONE();
  }
  {
    Head1Params hp;
    hp.num_pairs = (int)((T + 255) / 256);
    hp.T = (int)T;
    hp.bias = w->bh1;
    hp.qual = qual;
    CUtensorMap tm_r_st;
    DCB_CHECK(make_tmap_2d(&tm_r_st, g, T, kInner, 32));
    DCB_CHECK(launch_head1(ctx, tm_u, w->tm_h1, tm_r_st, hp));
  }
  DCB_STAGE_DONE();
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.T = (int)T;
  p.L = L;
  p.num_outer = (int)(T / 128);
  p.bias = w->bh2;
  p.r_in = g;
  p.w3 = w->w3;
  p.b3 = w->b3;
  p.logits = logits;
  p.labels = labels;
  DCB_CHECK(launch_gemm(ctx, G_HEAD2, tm_g, w->tm_h2, p));
  return DCB200_OK;
}

}  // namespace dcb
