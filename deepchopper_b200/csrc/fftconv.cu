// Hyena long convolution, fused with the short depthwise conv and both gates (SURVEY K4+K5, Appendix A):
//
//   zc = causal 3-tap depthwise conv of z (the in_proj output, channel-major bf16 [B,768,L])
//   x0, x1, v = zc[0:256], zc[256:512], zc[512:768]
//   y = ( causal_conv(v*x1, k) + (v*x1)*D ) * x0           -> bf16 channel-major [B,256,L]
//
// The reference does rfft/irfft of size 2L through cuFFT on fp32 [B,256,L] tensors (fftconv in the HF
// modeling_hyena.py, called from deepchopper/models/llm/hyena.py:34-41).  Here one CTA owns one
// channel of TWO batch rows: the two real signals are packed as re/im of one complex sequence (the
// filter is real, so the complex convolution convolves both rows at once), transformed in shared
// memory with a radix-16/8/4/2 decimation-in-frequency FFT (fp32, output left in digit-reversed order),
// multiplied by the cached filter spectrum (stored in the same digit-reversed order, pre-scaled by
// 1/N, with the D skip term folded in as k[0] += D) and inverted with the mirrored
// decimation-in-time passes -- no bit-reversal pass, no HBM round trip: HBM sees z once (6 B per
// token-channel) and y once (2 B).  FFT size N = next power of two >= 2L; the linear (causal)
// convolution is exact for any such N.
#include "common.cuh"
#include "fftconv.h"

namespace dcb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }  // one pad slot per 16: conflict-free strided access

template <int R> __device__ __forceinline__ constexpr int brev(int q) {
  int r = 0;
  for (int b = 1; b < R; b <<= 1) {
    r = (r << 1) | (q & 1);
    q >>= 1;
  }
  return r;
}

// multiply by exp(-+ 2 pi i k16 / 16); k16 is a compile-time constant after unrolling
template <bool INV> __device__ __forceinline__ float2 mul_w16(float2 d, int k16) {
  constexpr float C[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                          0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
  constexpr float S[8] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                          1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
  if (k16 == 0) return d;
  if (k16 == 4) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  const float c = C[k16], s = INV ? S[k16] : -S[k16];
  return make_float2(fmaf(d.x, c, -d.y * s), fmaf(d.x, s, d.y * c));
}

// R-point DFT in registers (radix-2 decimation in frequency); output element q lands in a[brev<R>(q)].
template <int R, bool INV> __device__ __forceinline__ void dft_small(float2 (&a)[R]) {
#pragma unroll
  for (int len = R; len >= 2; len >>= 1) {
    const int half = len >> 1;
#pragma unroll
    for (int blk = 0; blk < R; blk += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 u = a[blk + j], v = a[blk + j + half];
        a[blk + j] = make_float2(u.x + v.x, u.y + v.y);
        a[blk + j + half] = mul_w16<INV>(make_float2(u.x - v.x, u.y - v.y), j * (16 / len));
      }
    }
  }
}

// powers w[q] = w1^q for q in [1,R) from the table entries W^(j), W^(2j), W^(4j), W^(8j)
template <int R> __device__ __forceinline__ void twiddles(float2 (&w)[R], const float2* __restrict__ tw, int e) {
  w[0] = make_float2(1.f, 0.f);
  if (R >= 2) w[1] = __ldg(tw + e);
  if (R >= 4) {
    w[2] = __ldg(tw + 2 * e);
    w[3] = cmul(w[1], w[2]);
  }
  if (R >= 8) {
    w[4] = __ldg(tw + 4 * e);
#pragma unroll
    for (int q = 5; q < 8; ++q) w[q] = cmul(w[4], w[q - 4]);
  }
  if (R >= 16) {
    w[8] = __ldg(tw + 8 * e);
#pragma unroll
    for (int q = 9; q < 16; ++q) w[q] = cmul(w[8], w[q - 8]);
  }
}

// One in-place pass over blocks of length n (n = current sub-transform length), radix R.
//   forward (DIF):  x[base + q s] = W_n^(j q) * DFT_R(x[base + m s])_q
//   inverse (DIT):  exact mirror with conjugate twiddles (scale R, folded into the filter spectrum)
template <int R, bool INV, bool MULKF>
__device__ __forceinline__ void fft_pass(float2* X, int N, int n, const float2* __restrict__ tw,
                                         const float2* __restrict__ KF) {
  const int s = n / R;
  const int ls = 31 - __clz(s);
  const int tstep = N / n;
  for (int t = threadIdx.x; t < N / R; t += blockDim.x) {
    const int blk = t >> ls, j = t & (s - 1);
    const int base = blk * n + j;
    float2 a[R];
#pragma unroll
    for (int m = 0; m < R; ++m) a[m] = X[padi(base + m * s)];
    if (MULKF) {
#pragma unroll
      for (int m = 0; m < R; ++m) a[m] = cmul(a[m], __ldg(KF + base + m * s));
    }
    float2 w[R];
    if (s > 1) twiddles<R>(w, tw, j * tstep);
    if (!INV) {
      dft_small<R, false>(a);
#pragma unroll
      for (int q = 0; q < R; ++q) {
        float2 val = a[brev<R>(q)];
        if (s > 1 && q > 0) val = cmul(val, w[q]);
        X[padi(base + q * s)] = val;
      }
    } else {
      if (s > 1) {
#pragma unroll
        for (int q = 1; q < R; ++q) a[q] = cmulc(a[q], w[q]);
      }
      dft_small<R, true>(a);
#pragma unroll
      for (int pz = 0; pz < R; ++pz) X[padi(base + brev<R>(pz) * s)] = a[pz];
    }
  }
}

template <bool INV, bool MULKF>
__device__ __forceinline__ void fft_pass_r(int R, float2* X, int N, int n, const float2* tw, const float2* KF) {
  switch (R) {
    case 16: fft_pass<16, INV, MULKF>(X, N, n, tw, KF); break;
    case 8: fft_pass<8, INV, MULKF>(X, N, n, tw, KF); break;
    case 4: fft_pass<4, INV, MULKF>(X, N, n, tw, KF); break;
    default: fft_pass<2, INV, MULKF>(X, N, n, tw, KF); break;
  }
}

// forward transform of the sequence in X (length N), result in digit-reversed order
__device__ __forceinline__ void fft_forward(const FftPlan& pl, float2* X, const float2* tw) {
  int n = pl.N;
  for (int i = 0; i < pl.npass; ++i) {
    fft_pass_r<false, false>(pl.radix[i], X, pl.N, n, tw, nullptr);
    n /= pl.radix[i];
    __syncthreads();
  }
}
// inverse of fft_forward (times N), first pass multiplies by KF
__device__ __forceinline__ void fft_inverse_mul(const FftPlan& pl, float2* X, const float2* tw, const float2* KF) {
  int n = 1;
  for (int i = pl.npass - 1; i >= 0; --i) {
    n *= pl.radix[i];
    if (i == pl.npass - 1) fft_pass_r<true, true>(pl.radix[i], X, pl.N, n, tw, KF);
    else fft_pass_r<true, false>(pl.radix[i], X, pl.N, n, tw, nullptr);
    __syncthreads();
  }
}

__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ float bf2f(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// short conv of 8 consecutive tokens of one channel row: out[t] = w0 z[t-2] + w1 z[t-1] + w2 z[t] + b
__device__ __forceinline__ void short_conv8(const __nv_bfloat16* __restrict__ row, int t0, float w0, float w1, float w2,
                                            float bias, float (&out)[8]) {
  float z[10];
  float cur[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(row + t0)), cur);
  z[0] = t0 >= 2 ? bf2f(row + t0 - 2) : 0.f;
  z[1] = t0 >= 1 ? bf2f(row + t0 - 1) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) z[2 + i] = cur[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = fmaf(w0, z[i], fmaf(w1, z[i + 1], fmaf(w2, z[i + 2], bias)));
}

__global__ void __launch_bounds__(256) fftconv_kernel(const ConvParams p) {
  extern __shared__ float2 smem_f2[];
  float2* X = smem_f2;                       // N + N/16 complex
  float2* G = smem_f2 + (p.plan.N + p.plan.N / 16);  // L: short-conv'ed x0 gate of both rows
  const int c = blockIdx.y;
  const int b0 = 2 * blockIdx.x;
  const bool has1 = b0 + 1 < p.B;
  const int L = p.L, N = p.plan.N;
  const __nv_bfloat16* zb0 = p.z + (size_t)b0 * 768 * L;
  const __nv_bfloat16* zb1 = p.z + (size_t)(has1 ? b0 + 1 : b0) * 768 * L;
  // short-filter taps of the three channels feeding this output channel
  float sw[3][3], sbias[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int ch = g * 256 + c;
    sw[g][0] = __ldg(p.short_w + ch * 3 + 0);
    sw[g][1] = __ldg(p.short_w + ch * 3 + 1);
    sw[g][2] = __ldg(p.short_w + ch * 3 + 2);
    sbias[g] = __ldg(p.short_b + ch);
  }
  // ---- prologue: short conv + first gate -> X (re = row b0, im = row b0+1), x0 gate -> G ----------
  for (int g8 = threadIdx.x; g8 < N / 8; g8 += blockDim.x) {
    const int t0 = g8 * 8;
    if (t0 < L) {
      float x0a[8], x1a[8], va[8], x0b[8], x1b[8], vb[8];
      short_conv8(zb0 + (size_t)c * L, t0, sw[0][0], sw[0][1], sw[0][2], sbias[0], x0a);
      short_conv8(zb0 + (size_t)(256 + c) * L, t0, sw[1][0], sw[1][1], sw[1][2], sbias[1], x1a);
      short_conv8(zb0 + (size_t)(512 + c) * L, t0, sw[2][0], sw[2][1], sw[2][2], sbias[2], va);
      if (has1) {
        short_conv8(zb1 + (size_t)c * L, t0, sw[0][0], sw[0][1], sw[0][2], sbias[0], x0b);
        short_conv8(zb1 + (size_t)(256 + c) * L, t0, sw[1][0], sw[1][1], sw[1][2], sbias[1], x1b);
        short_conv8(zb1 + (size_t)(512 + c) * L, t0, sw[2][0], sw[2][1], sw[2][2], sbias[2], vb);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x0b[i] = x1b[i] = vb[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        X[padi(t0 + i)] = make_float2(va[i] * x1a[i], vb[i] * x1b[i]);
        G[t0 + i] = make_float2(x0a[i], x0b[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) X[padi(t0 + i)] = make_float2(0.f, 0.f);
    }
  }
  __syncthreads();
  fft_forward(p.plan, X, p.tw);
  fft_inverse_mul(p.plan, X, p.tw, p.KF + (size_t)c * N);
  // ---- epilogue: second gate, bf16, coalesced 16B stores -------------------------------------------
  __nv_bfloat16* y0 = p.y + ((size_t)b0 * 256 + c) * L;
  __nv_bfloat16* y1 = p.y + ((size_t)(b0 + 1) * 256 + c) * L;
  for (int g8 = threadIdx.x; g8 < L / 8; g8 += blockDim.x) {
    const int t0 = g8 * 8;
    uint32_t o0[4], o1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = X[padi(t0 + 2 * i)], b = X[padi(t0 + 2 * i + 1)];
      const float2 ga = G[t0 + 2 * i], gb = G[t0 + 2 * i + 1];
      __nv_bfloat162 r0 = __floats2bfloat162_rn(a.x * ga.x, b.x * gb.x);
      __nv_bfloat162 r1 = __floats2bfloat162_rn(a.y * ga.y, b.y * gb.y);
      o0[i] = *reinterpret_cast<uint32_t*>(&r0);
      o1[i] = *reinterpret_cast<uint32_t*>(&r1);
    }
    *reinterpret_cast<uint4*>(y0 + t0) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
    if (has1) *reinterpret_cast<uint4*>(y1 + t0) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
  }
}

// Filter spectrum: KF[c][:] = FFT_N( k[c][0:N/2] with k[0] += D[c], zero padded ) / N, digit-reversed order.
__global__ void __launch_bounds__(256) filter_spectrum_kernel(const float* __restrict__ k, int k_stride, int k_len,
                                                              const float* __restrict__ D, FftPlan plan,
                                                              const float2* __restrict__ tw, float2* __restrict__ KF) {
  extern __shared__ float2 smem_f2[];
  float2* X = smem_f2;
  const int c = blockIdx.x;
  const int N = plan.N;
  const int taps = min(N / 2, k_len);
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    float v = 0.f;
    if (t < taps) v = k[(size_t)c * k_stride + t];
    if (t == 0) v += D[c];
    X[padi(t)] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft_forward(plan, X, tw);
  const float inv = 1.0f / (float)N;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    const float2 v = X[padi(t)];
    KF[(size_t)c * N + t] = make_float2(v.x * inv, v.y * inv);
  }
}

FftPlan make_plan(int N) {
  FftPlan pl;
  pl.N = N;
  int lg = 0;
  while ((1 << lg) < N) ++lg;
  pl.npass = 0;
  const int rem = lg % 4;
  if (rem) pl.radix[pl.npass++] = 1 << rem;
  for (int i = 0; i < lg / 4; ++i) pl.radix[pl.npass++] = 16;
  return pl;
}

size_t conv_smem_bytes(int N, int L) { return (size_t)(N + N / 16) * 8 + (size_t)L * 8; }

static int conv_threads(int N) {
  int t = N / 16;
  if (t > 256) t = 256;
  if (t < 32) t = 32;
  return t;
}

int launch_fftconv(dcb200_ctx* ctx, const ConvParams& p) {
  const size_t smem = conv_smem_bytes(p.plan.N, p.L);
  if (smem > 227 * 1024) {
    set_error("fftconv: L=%d needs %zu bytes of shared memory (long-read path not built yet)", p.L, smem);
    return DCB200_EINVAL;
  }
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&fftconv_kernel), 227 * 1024));
  dim3 grid((p.B + 1) / 2, 256);
  ProfScope prof(ctx, K_CONV);
  fftconv_kernel<<<grid, conv_threads(p.plan.N), smem, ctx->stream>>>(p);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_filter_spectrum(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, const FftPlan& plan,
                           const float2* tw, float2* KF) {
  const size_t smem = (size_t)(plan.N + plan.N / 16) * 8;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&filter_spectrum_kernel), 227 * 1024));
  filter_spectrum_kernel<<<256, conv_threads(plan.N), smem, ctx->stream>>>(k, k_stride, k_len, D, plan, tw, KF);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
