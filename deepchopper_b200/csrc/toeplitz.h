#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
constexpr int kToepMaxL = 8192;  // the tensor-core (Toeplitz) long convolution covers L <= 8192 tokens
// per-layer table of Toeplitz core matrices (35 MB), built once per weight set
size_t toeplitz_table_bytes();
int launch_toeplitz_table(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, __nv_bfloat16* E);
// y = gate * causal_conv(vv, k'): all three activations bf16 [B,256,L], tensor maps from make_tmap_3d_rows
int launch_toeplitz_conv(dcb200_ctx* ctx, const __nv_bfloat16* E, const CUtensorMap& tm_vv, const CUtensorMap& tm_gate,
                         const CUtensorMap& tm_y, int B, int L);
int launch_shortconv_gate(dcb200_ctx* ctx, const __nv_bfloat16* z, const float* sw, const float* sb, int B, int L,
                          __nv_bfloat16* vv, __nv_bfloat16* gate);
}  // namespace dcb
