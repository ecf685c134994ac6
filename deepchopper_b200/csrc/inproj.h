#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
struct dcb200_ctx;
namespace dcb {
struct InprojParams {
  int num_tiles;         // T / 128
  int L;                 // padded read length (multiple of 128)
  const float* b_in;     // [768] in_linear bias
  const float* short_w;  // [768,3] short depthwise filter
  const float* short_b;  // [768]
  long long* trace;      // optional timeline trace buffer (DCB200_TRACE=inproj), or null
};
// tm_u: bf16 [T,256] box 64 x 144 rows;  tm_w: W_in [768,256] box 64 x 64 rows (half a weight box per CTA of a cluster, multicast);
// tm_vv / tm_gate: bf16 [B,256,L] box 64 (L) x 128 (C) x 1 (make_tmap_3d_chbox), TMA stores
int launch_inproj_conv(dcb200_ctx* ctx, const CUtensorMap& tm_u, const CUtensorMap& tm_w, const CUtensorMap& tm_vv,
                       const CUtensorMap& tm_gate, const InprojParams& p);
}  // namespace dcb
