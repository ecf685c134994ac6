// Persistent, warp-specialised tcgen05 GEMMs with fused epilogues for the Hyena layers and the head.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128/256, K=16) into TMEM
//   warp 2      TMEM allocator
//   warps 4-7   epilogue: tcgen05.ld the fp32 accumulator (thread == one of the 128 rows) and apply the
//               fused epilogue while the MMA warp fills the other TMEM accumulator stage
//
// Modes (reference ops they replace, SURVEY Appendix A / K3,K6,K7,K8):
//   INPROJ   z^T = W_in . LN1(h)^T + b      -> bf16 channel-major [B,768,L]  (operand roles swapped so
//            the accumulator rows are channels and a thread writes contiguous tokens)
//   OUTPROJ  h' = y . W_o^T + b + h ; m = LN2(h')   (A = y read MN-major straight from the conv's
//            channel-major output: no transpose pass)
//   FC1      g = gelu_tanh(m . W1^T + b1)   -> bf16 [T,1024]
//   FC2      h'' = g . W2^T + b2 + h' ; u = LN(h'')  (next layer's LN1, or ln_f after the last layer)
//   HEAD1    r = relu(hf . Wh1^T + b1) + q  -> bf16 [T,1024]        (head.py:94-97)
//   HEAD2    o = relu(r . Wh2^T + b2 + r) ; logits = o . W3^T + b3 ; label = l1 > l0   (head.py:98-102)
#include "common.cuh"
#include "ptx.cuh"
#include "gemm.h"

namespace dcb {

using namespace ptx;

constexpr int kStages = 4;
constexpr int kTileM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle row

template <int MODE> struct Traits;
template <> struct Traits<G_INPROJ>  { static constexpr int K = 256,  NT = 128, INNER = 6; static constexpr bool A_MN = false; };
template <> struct Traits<G_OUTPROJ> { static constexpr int K = 256,  NT = 256, INNER = 1; static constexpr bool A_MN = true;  };
template <> struct Traits<G_FC1>     { static constexpr int K = 256,  NT = 256, INNER = 4; static constexpr bool A_MN = false; };
template <> struct Traits<G_FC2>     { static constexpr int K = 1024, NT = 256, INNER = 1; static constexpr bool A_MN = false; };
template <> struct Traits<G_HEAD1>   { static constexpr int K = 256,  NT = 256, INNER = 4; static constexpr bool A_MN = false; };
template <> struct Traits<G_HEAD2>   { static constexpr int K = 1024, NT = 256, INNER = 4; static constexpr bool A_MN = false; };

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))   (F.gelu(approximate="tanh"))
  const float u = 0.7978845608028654f * fmaf(0.044715f * x * x, x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

template <int MODE>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Tr = Traits<MODE>;
  constexpr int NT = Tr::NT;
  constexpr int KCH = Tr::K / kBlockK;
  constexpr uint32_t A_BYTES = kTileM * 128;
  constexpr uint32_t B_BYTES = NT * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * NT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem ptr
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_outer = p.num_outer;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int o = blockIdx.x; o < num_outer; o += gridDim.x) {
        const int tok0 = o * kTileM;
        for (int i = 0; i < Tr::INNER; ++i) {
          for (int kc = 0; kc < KCH; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            if (MODE == G_INPROJ) {
              tma_load_2d(a_dst, &tmA, full_bar(stage), kc * kBlockK, i * kTileM);  // W_in rows (channels)
              tma_load_2d(b_dst, &tmB, full_bar(stage), kc * kBlockK, tok0);        // tokens
            } else {
              if (Tr::A_MN) {
                const int b = tok0 / p.L, l0 = tok0 % p.L;
                tma_load_3d(a_dst, &tmA, full_bar(stage), l0, kc * kBlockK, b);
                tma_load_3d(a_dst + 8192, &tmA, full_bar(stage), l0 + 64, kc * kBlockK, b);
              } else {
                tma_load_2d(a_dst, &tmA, full_bar(stage), kc * kBlockK, tok0);
              }
              tma_load_2d(b_dst, &tmB, full_bar(stage), kc * kBlockK, i * NT);
            }
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, NT, Tr::A_MN, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int o = blockIdx.x; o < num_outer; o += gridDim.x) {
        for (int i = 0; i < Tr::INNER; ++i) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * NT;
          for (int kc = 0; kc < KCH; ++kc) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
            const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              const uint64_t adesc = Tr::A_MN ? make_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                              : make_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t bdesc = make_desc_sw128(b_addr + k * 32, 16, 1024);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (kc | k) ? 1u : 0u);
            }
            umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int e = warp - 4;            // == warp % 4: the TMEM lane quadrant this warp may access
    const int row_in_tile = e * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t v[32];
    for (int o = blockIdx.x; o < num_outer; o += gridDim.x) {
      const int tok0 = o * kTileM;
      float lg0 = 0.f, lg1 = 0.f;  // HEAD2 partial logits
      for (int i = 0; i < Tr::INNER; ++i) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * NT + ((uint32_t)(e * 32) << 16);

        if (MODE == G_INPROJ) {
          const int ch = i * kTileM + row_in_tile;
          const float bias = __ldg(p.bias + ch);
          const int b = tok0 / p.L, l0 = tok0 % p.L;
          __nv_bfloat16* dst = p.out_bf16 + ((size_t)b * 768 + ch) * p.L + l0;
#pragma unroll 1
          for (int c = 0; c < NT / 32; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
            uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(v[8 * q + 0]) + bias, __uint_as_float(v[8 * q + 1]) + bias);
              w.y = pack_bf16(__uint_as_float(v[8 * q + 2]) + bias, __uint_as_float(v[8 * q + 3]) + bias);
              w.z = pack_bf16(__uint_as_float(v[8 * q + 4]) + bias, __uint_as_float(v[8 * q + 5]) + bias);
              w.w = pack_bf16(__uint_as_float(v[8 * q + 6]) + bias, __uint_as_float(v[8 * q + 7]) + bias);
              d4[q] = w;
            }
          }
        } else if (MODE == G_OUTPROJ || MODE == G_FC2) {
          const size_t row = (size_t)tok0 + row_in_tile;
          const float4* res4 = reinterpret_cast<const float4*>(p.resid + row * 256);
          float4* h4 = p.h_out ? reinterpret_cast<float4*>(p.h_out + row * 256) : nullptr;
          float sum = 0.f;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 r = __ldg(res4 + c * 8 + q);
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias) + c * 8 + q);
              float4 x;
              x.x = __uint_as_float(v[4 * q + 0]) + bb.x + r.x;
              x.y = __uint_as_float(v[4 * q + 1]) + bb.y + r.y;
              x.z = __uint_as_float(v[4 * q + 2]) + bb.z + r.z;
              x.w = __uint_as_float(v[4 * q + 3]) + bb.w + r.w;
              sum += (x.x + x.y) + (x.z + x.w);
              if (h4) h4[c * 8 + q] = x;
              v[4 * q + 0] = __float_as_uint(x.x);
              v[4 * q + 1] = __float_as_uint(x.y);
              v[4 * q + 2] = __float_as_uint(x.z);
              v[4 * q + 3] = __float_as_uint(x.w);
            }
            tmem_st32(t_row + c * 32, v);
          }
          tmem_st_wait();
          const float mean = sum * (1.0f / 256.0f);
          float var = 0.f;
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = __uint_as_float(v[j]) - mean;
              var = fmaf(d, d, var);
            }
          }
          const float rstd = rsqrtf(var * (1.0f / 256.0f) + 1e-5f);
          uint4* u4 = reinterpret_cast<uint4*>(p.out_bf16 + row * 256);
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float y[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int col = c * 32 + q * 8 + j;
                y[j] = fmaf((__uint_as_float(v[q * 8 + j]) - mean) * rstd, __ldg(p.ln_g + col), __ldg(p.ln_b + col));
              }
              uint4 w;
              w.x = pack_bf16(y[0], y[1]);
              w.y = pack_bf16(y[2], y[3]);
              w.z = pack_bf16(y[4], y[5]);
              w.w = pack_bf16(y[6], y[7]);
              u4[c * 4 + q] = w;
            }
          }
        } else if (MODE == G_FC1 || MODE == G_HEAD1) {
          const size_t row = (size_t)tok0 + row_in_tile;
          const float qv = (MODE == G_HEAD1) ? __ldg(p.qual + row) : 0.f;
          uint4* g4 = reinterpret_cast<uint4*>(p.out_bf16 + row * 1024 + i * NT);
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float y[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x = __uint_as_float(v[q * 8 + j]) + __ldg(p.bias + i * NT + c * 32 + q * 8 + j);
                y[j] = (MODE == G_FC1) ? gelu_tanh(x) : (fmaxf(x, 0.f) + qv);
              }
              uint4 w;
              w.x = pack_bf16(y[0], y[1]);
              w.y = pack_bf16(y[2], y[3]);
              w.z = pack_bf16(y[4], y[5]);
              w.w = pack_bf16(y[6], y[7]);
              g4[c * 4 + q] = w;
            }
          }
        } else {  // G_HEAD2
          const size_t row = (size_t)tok0 + row_in_tile;
          const uint4* r4 = reinterpret_cast<const uint4*>(p.r_in + row * 1024 + i * NT);
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 rr = __ldg(r4 + c * 4 + q);
              const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int col = i * NT + c * 32 + q * 8 + j;
                const uint32_t pr = rw[j >> 1];
                const float rres = __uint_as_float((j & 1) ? (pr & 0xffff0000u) : (pr << 16));
                const float ov = fmaxf(__uint_as_float(v[q * 8 + j]) + __ldg(p.bias + col) + rres, 0.f);
                lg0 = fmaf(ov, __ldg(p.w3 + col), lg0);
                lg1 = fmaf(ov, __ldg(p.w3 + 1024 + col), lg1);
              }
            }
          }
          if (i == Tr::INNER - 1) {
            lg0 += __ldg(p.b3);
            lg1 += __ldg(p.b3 + 1);
            if (p.logits) reinterpret_cast<float2*>(p.logits)[row] = make_float2(lg0, lg1);
            if (p.labels) p.labels[row] = lg1 > lg0 ? 1 : 0;
          }
        }
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int MODE> static size_t smem_bytes() {
  using Tr = Traits<MODE>;
  return (size_t)kStages * (kTileM * 128 + Tr::NT * 128) + 1024 + 256;
}

template <int MODE> static int launch_mode(dcb200_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p) {
  static bool configured = false;
  const size_t smem = smem_bytes<MODE>();
  if (!configured) {
    DCB_CUDA(cudaFuncSetAttribute(gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = p.num_outer < ctx->sm_count ? p.num_outer : ctx->sm_count;
  static const int kinds[6] = {K_INPROJ, K_OUTPROJ, K_FC1, K_FC2, K_HEAD1, K_HEAD2};
  ProfScope prof(ctx, kinds[MODE]);
  gemm_kernel<MODE><<<grid, 256, smem, ctx->stream>>>(a, b, p);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_gemm(dcb200_ctx* ctx, int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p) {
  switch (mode) {
    case G_INPROJ: return launch_mode<G_INPROJ>(ctx, a, b, p);
    case G_OUTPROJ: return launch_mode<G_OUTPROJ>(ctx, a, b, p);
    case G_FC1: return launch_mode<G_FC1>(ctx, a, b, p);
    case G_FC2: return launch_mode<G_FC2>(ctx, a, b, p);
    case G_HEAD1: return launch_mode<G_HEAD1>(ctx, a, b, p);
    case G_HEAD2: return launch_mode<G_HEAD2>(ctx, a, b, p);
  }
  set_error("bad gemm mode %d", mode);
  return DCB200_EINVAL;
}

// ---- tensor maps ---------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 row-major [rows][cols] (cols contiguous); box = box_rows x 64 columns, 128B swizzle
int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DCB200_ECUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d rows=%llu cols=%llu) failed: %d", (unsigned long long)rows, (unsigned long long)cols, (int)r);
    return DCB200_ECUDA;
  }
  return DCB200_OK;
}

// bf16 [B][C][L] (L contiguous); box = 64 (L) x 64 (C) x 1, 128B swizzle: an MN-major operand tile
int make_tmap_3d_cm(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DCB200_ECUDA;
  }
  cuuint64_t dims[3] = {L, C, B};
  cuuint64_t strides[2] = {L * 2, C * L * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return DCB200_ECUDA;
  }
  return DCB200_OK;
}

}  // namespace dcb
