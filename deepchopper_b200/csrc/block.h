#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
// Passed BY VALUE as a kernel parameter (constant bank; see mlp.h).  6 KB.  The LayerNorms (norm2, and the next
// layer's norm1 / ln_f applied to the block output) only normalise: their affine parts are folded into the weights of
// the Linear that follows each of them (model.cu upload_folded_linear).
struct BlockParams {
  int num_pairs;      // ceil(T / 256): 256-token tiles, one per CTA pair
  int T;              // tokens
  int L;              // padded read length (tiles never straddle reads: L % 128 == 0)
  long long* trace;   // optional timeline trace buffer, or null
  float bo[256];      // out_linear bias
  float b1[1024];
  float b2[256];
};
// tm_y: bf16 [B,256,L] box 64 (L) x 64 (C) x 1 (make_tmap_3d_cm);  tm_wo: [256,256] box 64 x 128 rows;
// tm_w1: [1024,256] box 64 x 128 rows;  tm_w2: [256,1024] box 64 x 128 rows;
// tm_hin / tm_hout: fp32 [T,256] box 32 x 128;  tm_u: bf16 [T,256] box 64 x 128
int launch_block(dcb200_ctx* ctx, const CUtensorMap& tm_y, const CUtensorMap& tm_wo, const CUtensorMap& tm_w1,
                 const CUtensorMap& tm_w2, const CUtensorMap& tm_hin, const CUtensorMap& tm_hout, const CUtensorMap& tm_u,
                 const BlockParams& p);
}  // namespace dcb
