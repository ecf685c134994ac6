// Blocked shared-memory FFT long convolution (see lconv_core.cuh for the algorithm): the product path for reads longer
// than the measured crossover with the tensor-core Toeplitz kernel (toeplitz.cu), whose MMA work grows with L.
// Reference: fftconv of the HF modeling_hyena.py, call site deepchopper/models/llm/hyena.py:34-41.
#include "common.cuh"
#include "lconv.h"
#include "lconv_core.cuh"

namespace dcb {

using namespace lc;

// T1[n] = W_8192^n, T2[k] = W_4096^k, T3[slot] = -i W_16384^k(slot): evaluated in double, rounded once
__global__ void lconv_twiddle_kernel(float2* __restrict__ tw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTwTotal) return;
  double s, c;
  if (i < kTwT2) {
    sincospi(-2.0 * (double)i / (double)kP, &s, &c);
  } else if (i < kTwT3) {
    sincospi(-2.0 * (double)(i - kTwT2) / (double)kH, &s, &c);
  } else {
    double s0, c0;
    sincospi(-2.0 * (double)slot_freq(i - kTwT3) / (double)(2 * kP), &s0, &c0);
    c = s0;   // -i (c0 + i s0) = s0 - i c0
    s = -c0;
  }
  tw[i] = make_float2((float)c, (float)s);
}

// Filter spectra: for (channel c, block distance d) the 2P-point real sequence
//   c_d[m] = k'[dP + m] (m < P),  0 (m = P),  k'[dP + m - 2P] (m > P),      k'[0] = k[0] + D
// transformed like a data block and stored in slot layout, scaled by 1 / (8 N).
__global__ void __launch_bounds__(kThreads) lconv_filter_spectrum_kernel(const float* __restrict__ k, int k_stride, int k_len,
                                                                         const float* __restrict__ D,
                                                                         const float2* __restrict__ tw, int nbK,
                                                                         float4* __restrict__ K) {
  extern __shared__ float2 lconv_smem[];
  float2* X = lconv_smem;
  const int c = blockIdx.x, d = blockIdx.y;
  const int tid = threadIdx.x;
  const float* kc = k + (size_t)c * k_stride;
  const float Dc = D[c];
  auto tap = [&](int m) -> float {  // c_d[m]
    int u;
    if (m < kP) u = d * kP + m;
    else if (m == kP) return 0.f;
    else u = d * kP + m - 2 * kP;
    float v = (u >= 0 && u < k_len) ? kc[u] : 0.f;
    if (u == 0) v += Dc;
    return v;
  };
  for (int n = tid; n < kH; n += kThreads) {
    const float2 zlo = make_float2(tap(2 * n), tap(2 * n + 1));
    const float2 zhi = make_float2(tap(2 * (n + kH)), tap(2 * (n + kH) + 1));
    prologue_store_full(X, tw + kTwT1, n, zlo, zhi);
  }
  __syncthreads();
  const ThreadTw t = load_thread_tw(tw, tid);
  radix16_pass<0, false>(X, t.p0a, t.p0b, tid);
  __syncthreads();
  radix16_pass<1, false>(X, t.p1a, t.p1b, tid);
  __syncthreads();
  radix16_pass<2, false>(X, t.p1a, t.p1b, tid);
  __syncthreads();
  float4* dst = K + ((size_t)c * nbK + d) * kSlots;
  const float scale = 1.0f / (8.0f * (float)kP);
  for (int slot = tid; slot < kSlots; slot += kThreads) dst[slot] = spectrum_slot(X, tw + kTwT3, slot, scale);
}

struct LconvParams {
  const __nv_bfloat16* vv;
  const __nv_bfloat16* gate;
  __nv_bfloat16* y;
  const float4* K;     // [256][nbK][kSlots]
  const float2* tw;
  float4* scratch;     // [gridDim.x][kMaxBlocks - 1][kSlots]
  int B, L, nb, nbK;
};

// Two CTAs per SM (128 registers, no spills): measured 15-20 % faster than three CTAs at 80 registers with the twiddle
// arrays spilling to local memory.
__global__ void __launch_bounds__(kThreads, 2) lconv_kernel(const LconvParams p) {
  extern __shared__ float2 lconv_smem[];
  float2* X = lconv_smem;
  const int tid = threadIdx.x;
  const float2* T3 = p.tw + kTwT3;
  const ThreadTw t = load_thread_tw(p.tw, tid);
  float4* S = p.scratch + (size_t)blockIdx.x * (kMaxBlocks - 1) * kSlots;
  // L2 policies: block / filter spectra are re-read (keep), vv / gate / y pass through once (stream)
  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  const int n_items = 256 * p.B;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    // channel-major order: the CTAs in flight share the filter spectra of a few channels (L2 hits)
    const int c = item / p.B, b = item % p.B;
    const size_t row = ((size_t)b * 256 + c) * p.L;
    const uint32_t* v32 = reinterpret_cast<const uint32_t*>(p.vv + row);   // one word = the bf16 pair of a complex point
    const uint32_t* g32 = reinterpret_cast<const uint32_t*>(p.gate + row);
    uint32_t* y32 = reinterpret_cast<uint32_t*>(p.y + row);
    const float4* Kc = p.K + (size_t)c * p.nbK * kSlots;
    for (int i = 0; i < p.nb; ++i) {
      const int n0 = i * kH;                  // first complex point (token pair) of this block
      const int lim = p.L / 2 - n0;           // points of this block that lie inside the read
      {
        uint32_t zraw[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int n = tid + 256 * m;
          zraw[m] = n < lim ? ld_u32_hint(v32 + n0 + n, pol_stream) : 0u;
        }
        fwd_pass0_fused(X, zraw, t, tid);
      }
      __syncthreads();
      radix16_pass<1, false>(X, t.p1a, t.p1b, tid);
      __syncthreads();
      radix16_pass<2, false>(X, t.p1a, t.p1b, tid);
      __syncthreads();
#pragma unroll 1
      for (int it = 0; it < kSlots / kThreads; it += 4) pointwise_group<4>(X, T3, Kc, S, i, p.nb, it, tid, pol_keep);
      __syncthreads();
      radix16_pass<2, true>(X, t.p1a, t.p1b, tid);
      __syncthreads();
      radix16_pass<1, true>(X, t.p1a, t.p1b, tid);
      __syncthreads();
      {
        uint32_t out[16];
        inv_pass0_fused(X, t, tid, [&](int m) -> uint32_t {
          const int n = tid + 256 * m;
          return n < lim ? ld_u32_hint(g32 + n0 + n, pol_stream) : 0u;
        }, out);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const int n = tid + 256 * m;
          if (n < lim) st_u32_hint(y32 + n0 + n, out[m], pol_stream);
        }
      }
      __syncthreads();  // the next block's pass 0 overwrites the array
    }
  }
  // The scratch is dead now: drop its lines from L2 instead of writing them back.  (Measured: raising the persisting-L2
  // set-aside for the evict_last lines makes this kernel 10 % faster and in_proj / block 50 % slower -- the carve-out is
  // device-wide -- so the set-aside stays at its default and the policies are only hints.)
  if (p.nb > 1) {
    const char* sb = reinterpret_cast<const char*>(S);
    for (int ofs = tid * 128; ofs < (p.nb - 1) * kSlots * (int)sizeof(float4); ofs += kThreads * 128)
      asm volatile("discard.global.L2 [%0], 128;" ::"l"(sb + ofs) : "memory");
  }
}

constexpr size_t kLconvSmem = (size_t)kXFloat2 * sizeof(float2);

int lconv_max_len() { return kMaxBlocks * kP; }
size_t lconv_twiddle_bytes() { return (size_t)kTwTotal * sizeof(float2); }
size_t lconv_spectrum_bytes(int nbK) { return (size_t)256 * nbK * kSlots * sizeof(float4); }
int lconv_blocks_for(int L) { return (L + kP - 1) / kP; }

int launch_lconv_twiddles(dcb200_ctx* ctx, float2* tw) {
  lconv_twiddle_kernel<<<(kTwTotal + 255) / 256, 256, 0, ctx->stream>>>(tw);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_lconv_filter_spectrum(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, const float2* tw,
                                 int nbK, float4* K) {
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&lconv_filter_spectrum_kernel), kLconvSmem));
  dim3 grid(256, nbK);
  lconv_filter_spectrum_kernel<<<grid, kThreads, kLconvSmem, ctx->stream>>>(k, k_stride, k_len, D, tw, nbK, K);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_lconv(dcb200_ctx* ctx, const __nv_bfloat16* vv, const __nv_bfloat16* gate, __nv_bfloat16* y, const float4* K,
                 int nbK, const float2* tw, int B, int L) {
  if (L <= 0 || L % 128 != 0 || L > lconv_max_len()) {
    set_error("fft long conv: L=%d must be a multiple of 128 and <= %d", L, lconv_max_len());
    return DCB200_EINVAL;
  }
  LconvParams p;
  p.vv = vv;
  p.gate = gate;
  p.y = y;
  p.K = K;
  p.tw = tw;
  p.B = B;
  p.L = L;
  p.nb = lconv_blocks_for(L);
  p.nbK = nbK;
  if (p.nb > nbK) {
    set_error("fft long conv: filter spectra built for %d blocks, L=%d needs %d", nbK, L, p.nb);
    return DCB200_EINVAL;
  }
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&lconv_kernel), kLconvSmem));
  int per_sm = 0;
  DCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lconv_kernel, kThreads, kLconvSmem));
  if (per_sm < 1) per_sm = 1;
  const int n_items = 256 * B;
  int grid = per_sm * ctx->sm_count;
  if (grid > n_items) grid = n_items;
  DevBuf& sc = ctx->buf("lconv_scratch");
  DCB_CHECK(sc.reserve((size_t)grid * (kMaxBlocks - 1) * kSlots * sizeof(float4)));
  p.scratch = sc.as<float4>();
  ProfScope prof(ctx, K_CONV);
  lconv_kernel<<<grid, kThreads, kLconvSmem, ctx->stream>>>(p);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
