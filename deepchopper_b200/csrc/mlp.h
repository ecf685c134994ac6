#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
struct MlpParams {
  int num_pairs;      // ceil(T / 256): 256-token tiles, one per CTA pair
  const float* b1;    // [1024]
  const float* b2;    // [256]
  const float* ln_g;  // [256] LayerNorm applied to the block output (next layer's norm1, or ln_f)
  const float* ln_b;  // [256]
};
// tm_m / tm_u: bf16 [T,256] box 64 x 128;  tm_w1: [1024,256] box 64 x 64 rows;  tm_w2: [256,1024] box 64 x 128 rows;
// tm_hin / tm_hout: fp32 [T,256] box 32 x 128 (make_tmap_2d_f32)
int launch_mlp(dcb200_ctx* ctx, const CUtensorMap& tm_m, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
               const CUtensorMap& tm_hin, const CUtensorMap& tm_hout, const CUtensorMap& tm_u, const MlpParams& p);
}  // namespace dcb
