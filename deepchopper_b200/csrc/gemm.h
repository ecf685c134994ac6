#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

struct dcb200_ctx;

namespace dcb {

enum GemmMode { G_HEAD2 = 3 };

struct GemmParams {
  int T;          // tokens = B * L (multiple of 128)
  int L;          // padded read length (multiple of 128)
  int num_outer;  // T / 128 token tiles
  const float* bias;
  const __nv_bfloat16* r_in;  // HEAD2: r [T,1024]
  const float* w3;          // HEAD2: [2,1024]
  const float* b3;          // HEAD2: [2]
  float* logits;            // HEAD2: [T,2] or null
  uint8_t* labels;          // HEAD2: [T] or null
};

int launch_gemm(dcb200_ctx* ctx, int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p);
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
int make_tmap_2d_f32(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols);
int make_tmap_3d_cm(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L);
int make_tmap_3d_chbox(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L);
int make_tmap_3d_rows(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L);

}  // namespace dcb
