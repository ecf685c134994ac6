#pragma once
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
constexpr int kToepMaxBlocks = 32;  // Toeplitz path covers L <= 32 * 128 = 4096 tokens
int launch_toeplitz_build(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, int nb_max,
                          __nv_bfloat16* T);
int launch_shortconv_gate(dcb200_ctx* ctx, const __nv_bfloat16* z, const float* sw, const float* sb, int B, int L,
                          __nv_bfloat16* vv, __nv_bfloat16* gate);
}  // namespace dcb
