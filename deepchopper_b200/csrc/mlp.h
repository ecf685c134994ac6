#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
// Passed BY VALUE as a kernel parameter: the vectors sit in the constant bank, where the epilogue's warp-uniform reads
// are broadcasts (the kernel leaves no L1 for __ldg).  7 KB; CUDA >= 12.1 allows 32 KB of parameters.
struct MlpParams {
  int num_pairs;      // ceil(T / 256): 256-token tiles, one per CTA pair
  int T;              // tokens
  const float* h_in;  // fp32 [T,256] residual (read through tm_hin)
  float* h_out;       // fp32 [T,256]
  __nv_bfloat16* u_out;  // bf16 [T,256]
  long long* trace;   // optional timeline trace buffer (3 x 4096 x 2 int64), or null
  float b1[1024];
  float b2[256];
  float ln_g[256];    // LayerNorm applied to the block output (next layer's norm1, or ln_f)
  float ln_b[256];
};
// tm_m / tm_u: bf16 [T,256] box 64 x 128;  tm_w1: [1024,256] box 64 x 128 rows;  tm_w2: [256,1024] box 64 x 128 rows;
// tm_hin / tm_hout: fp32 [T,256] box 32 x 128 (make_tmap_2d_f32); tm_hin is only used for L2 prefetch
int launch_mlp(dcb200_ctx* ctx, const CUtensorMap& tm_m, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
               const CUtensorMap& tm_hin, const CUtensorMap& tm_hout, const CUtensorMap& tm_u, const MlpParams& p);
}  // namespace dcb
