#pragma once
#include <cuda.h>
struct dcb200_ctx;
namespace dcb {
struct Head1Params {
  int num_pairs;      // ceil(T / 256): 256-token tiles, one per CTA pair
  int T;              // tokens (multiple of 128)
  const float* bias;  // [1024] (ln_f affine folded in)
  const float* qual;  // [T] normalised quality per token
};
// tm_u: bf16 [T,256] box 64 x 128 rows;  tm_w: Wh1 bf16 [1024,256] box 64 x 128 rows;
// tm_r: r bf16 [T,1024] box 64 x 32 rows (TMA stores)
int launch_head1(dcb200_ctx* ctx, const CUtensorMap& tm_u, const CUtensorMap& tm_w, const CUtensorMap& tm_r,
                 const Head1Params& p);
}  // namespace dcb
