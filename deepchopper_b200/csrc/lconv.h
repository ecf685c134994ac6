#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct dcb200_ctx;

namespace dcb {

// Blocked shared-memory FFT long convolution (lconv.cu): y = gate * causal_conv(vv, k'), activations bf16 [B,256,L]
int lconv_max_len();                       // 32768
size_t lconv_twiddle_bytes();
size_t lconv_spectrum_bytes(int nbK);      // per layer: [256][nbK][4096] float4
int lconv_blocks_for(int L);               // 8192-token blocks of a read of L tokens
int launch_lconv_twiddles(dcb200_ctx* ctx, float2* tw);
int launch_lconv_filter_spectrum(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, const float2* tw,
                                 int nbK, float4* K);
int launch_lconv(dcb200_ctx* ctx, const __nv_bfloat16* vv, const __nv_bfloat16* gate, __nv_bfloat16* y, const float4* K,
                 int nbK, const float2* tw, int B, int L);

}  // namespace dcb
