#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct dcb200_ctx;

namespace dcb {

struct FftPlan {
  int N;
  int npass;
  int radix[4];  // forward (DIF) order: first pass works on blocks of length N
};

struct ConvParams {
  const __nv_bfloat16* z;  // [B,768,L] in_proj output (before the short conv)
  __nv_bfloat16* y;        // [B,256,L]
  const float* short_w;    // [768,3]
  const float* short_b;    // [768]
  const float2* KF;        // [256,N] filter spectrum of this layer for this N
  const float2* tw;        // [N] exp(-2 pi i k / N)
  int B, L;
  FftPlan plan;
};

FftPlan make_plan(int N);
size_t conv_smem_bytes(int N, int L);
int launch_fftconv(dcb200_ctx* ctx, const ConvParams& p);
int launch_filter_spectrum(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, const FftPlan& plan,
                           const float2* tw, float2* KF);

}  // namespace dcb
