#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
constexpr int kToepMaxL = 32768;  // the tensor-core (Toeplitz) long convolution covers the model's whole range
// per-layer table of Toeplitz core matrices for reads up to `cap` tokens (16 cap + 4096 bytes per channel: 35 MB at
// cap = 8192, 135 MB at 32768), built once per weight set
size_t toeplitz_table_bytes(int cap);
int launch_toeplitz_table(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, int cap, __nv_bfloat16* E);
// y = gate * causal_conv(vv, k'): all three activations bf16 [B,256,L], tensor maps from make_tmap_3d_rows
int launch_toeplitz_conv(dcb200_ctx* ctx, const __nv_bfloat16* E, int cap, const CUtensorMap& tm_vv, const CUtensorMap& tm_gate,
                         const CUtensorMap& tm_y, int B, int L);
}  // namespace dcb
