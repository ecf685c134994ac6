// Persistent, warp-specialised tcgen05 GEMM with a fused epilogue for the second half of the classification head (the
// Hyena layers' projections live in inproj.cu and block.cu, the head's first Linear in head1.cu).
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a 3/4-stage smem ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128/256, K=16) into TMEM
//   warp 2      TMEM allocator
//   warps 4-11  epilogue (two warps per TMEM lane quadrant, each owning half of the columns): tcgen05.ld
//               the fp32 accumulator, transpose through a private smem stage so that every global load /
//               store instruction covers whole 128-byte lines, apply the fused epilogue, while the MMA
//               warp fills the other TMEM accumulator stage
//
// Modes (reference ops they replace, SURVEY Appendix A / K3,K6,K7,K8):
//   HEAD2    o = relu(r . Wh2^T + b2 + r) ; logits = o . W3^T + b3 ; label = l1 > l0   (head.py:98-102)
#include "common.cuh"
#include "ptx.cuh"
#include "gemm.h"

namespace dcb {

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kBlockK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kEpiWarps = 8;  // two warps per TMEM lane quadrant, each owning half of the accumulator columns
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kStagePitch = 144;                    // bytes per staged row: 128 B payload + 16 B pad (conflict-free)
constexpr int kStageBytes = 32 * kStagePitch;       // per epilogue warp
constexpr int kVecFloats = 3 * 1024;                // column vectors cached in smem (bias / LN gamma,beta / linear3)

template <int MODE> struct Traits;
template <> struct Traits<G_HEAD2> { static constexpr int K = 1024, NT = 256, INNER = 4, STAGES = 3; };

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Tr = Traits<MODE>;
  constexpr int NT = Tr::NT;
  constexpr int kStages = Tr::STAGES;
  constexpr int KCH = Tr::K / kBlockK;
  constexpr uint32_t A_BYTES = kTileM * 128;
  constexpr uint32_t B_BYTES = NT * 128;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * NT;
  constexpr int HALF = NT / 2;  // accumulator columns owned by one epilogue warp

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_area = smem + kStages * STAGE_BYTES;                       // kEpiWarps * kStageBytes
  float* vec = reinterpret_cast<float*>(stage_area + kEpiWarps * kStageBytes);  // kVecFloats
  float2* stats = reinterpret_cast<float2*>(vec + kVecFloats);              // [2 parity][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 2 * 2 * 128);
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
      mbar_init(tempty_bar(s), 32 * kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), TMEM_COLS);
    tmem_relinquish();
  }
  // column vectors -> smem (epilogue warps read them as broadcast / lane-fixed float4)
  if (MODE == G_HEAD2) {
    for (int i = threadIdx.x; i < 1024; i += kThreads) {
      vec[i] = p.bias[i];
      vec[1024 + i] = p.w3[i];
      vec[2048 + i] = p.w3[1024 + i];
    }
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
        constexpr int nk = KCH;
        for (int i = 0; i < Tr::INNER; ++i) {
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
            const uint32_t b_dst = a_dst + A_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(a_dst, &tmA, full_bar(stage), kc * kBlockK, tok0);
            tma_load_2d(b_dst, &tmB, full_bar(stage), kc * kBlockK, i * NT);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, the elected lane issues (ptx.cuh umma_bf16_x4_e) =====
    {
      const uint32_t el = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, NT, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int o = blockIdx.x; o < num_outer; o += gridDim.x) {
        constexpr int nk = KCH;
        for (int i = 0; i < Tr::INNER; ++i) {
          mbar_wait(tempty_bar(acc), acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * NT;
          for (int kc = 0; kc < nk; ++kc) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
            const uint32_t b_addr = a_addr + A_BYTES;
            // K-major tiles advance 32 B per K = 16 slice
            umma_bf16_x4_e<1>(d_tmem, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc,
                              kc ? 1u : 0u, el);
            umma_commit_e(empty_bar(stage), el);  // frees the smem slot once these MMAs retire
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_e(tfull_bar(acc), el);  // accumulator complete -> epilogue
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    // Row-owner layout (what tcgen05.ld 32x32b gives): thread == accumulator row `quad*32 + lane`.
    // T layout (what global memory wants): lane -> (row 4i + lane/8, 16-byte piece lane%8) of a 128-byte
    // wide row chunk, so one warp instruction touches 4 full 128-byte lines.  The per-warp smem stage
    // converts between the two.
    const int e = warp - 4;
    const int quad = warp & 3;   // TMEM lane quadrant this warp may access
    const int half = e >> 2;     // which half of the accumulator columns
    uint8_t* stg = stage_area + e * kStageBytes;
    const uint32_t stg_own = smem_u32(stg) + lane * kStagePitch;                       // my row (row-owner)
    const int trow0 = lane >> 3, piece = lane & 7;
    const uint8_t* stg_t = stg + trow0 * kStagePitch + piece * 16;                     // + i * 4 * kStagePitch
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t tile_parity = 0;
    uint32_t v[32];
    for (int o = blockIdx.x; o < num_outer; o += gridDim.x) {
      const int tok0 = o * kTileM;
      float lg0[8], lg1[8];  // HEAD2 partial logits of my 8 T-layout rows
#pragma unroll
      for (int i = 0; i < 8; ++i) lg0[i] = lg1[i] = 0.f;
      for (int it = 0; it < Tr::INNER; ++it) {
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * NT + half * HALF + ((uint32_t)(quad * 32) << 16);

        {
          const size_t row_t = (size_t)tok0 + quad * 32 + trow0;
#pragma unroll 1
          for (int c = 0; c < HALF / 32; ++c) {
            tmem_ld32(t_row + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_own + q * 16), "r"(v[4 * q]),
                           "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                           : "memory");
            __syncwarp();
            const int col = it * NT + half * HALF + c * 32 + piece * 4;
            const float4 bb = *reinterpret_cast<const float4*>(vec + col);
            const float4 wa = *reinterpret_cast<const float4*>(vec + 1024 + col);
            const float4 wb = *reinterpret_cast<const float4*>(vec + 2048 + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 a = *reinterpret_cast<const float4*>(stg_t + i * 4 * kStagePitch);
              const uint2 rr = __ldg(reinterpret_cast<const uint2*>(p.r_in + (row_t + 4 * i) * 1024 + col));
              const float o0 = fmaxf(a.x + bb.x + __uint_as_float(rr.x << 16), 0.f);
              const float o1 = fmaxf(a.y + bb.y + __uint_as_float(rr.x & 0xffff0000u), 0.f);
              const float o2 = fmaxf(a.z + bb.z + __uint_as_float(rr.y << 16), 0.f);
              const float o3 = fmaxf(a.w + bb.w + __uint_as_float(rr.y & 0xffff0000u), 0.f);
              lg0[i] = fmaf(o0, wa.x, fmaf(o1, wa.y, fmaf(o2, wa.z, fmaf(o3, wa.w, lg0[i]))));
              lg1[i] = fmaf(o0, wb.x, fmaf(o1, wb.y, fmaf(o2, wb.z, fmaf(o3, wb.w, lg1[i]))));
            }
            __syncwarp();
          }
          if (it == Tr::INNER - 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
              for (int d = 1; d < 8; d <<= 1) {
                lg0[i] += __shfl_xor_sync(0xffffffffu, lg0[i], d);
                lg1[i] += __shfl_xor_sync(0xffffffffu, lg1[i], d);
              }
            }
            float2* st = stats + tile_parity * 256;
            if (half == 1 && piece == 0) {
#pragma unroll
              for (int i = 0; i < 8; ++i) st[quad * 32 + trow0 + 4 * i] = make_float2(lg0[i], lg1[i]);
            }
            named_bar_sync(1 + quad, 64);
            if (half == 0 && piece == 0) {
              const float b30 = __ldg(p.b3), b31 = __ldg(p.b3 + 1);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 o2 = st[quad * 32 + trow0 + 4 * i];
                const float l0 = lg0[i] + o2.x + b30, l1 = lg1[i] + o2.y + b31;
                const size_t row = row_t + 4 * i;
                if (p.logits) reinterpret_cast<float2*>(p.logits)[row] = make_float2(l0, l1);
                if (p.labels) p.labels[row] = l1 > l0 ? 1 : 0;
              }
            }
            tile_parity ^= 1;
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
  return (size_t)Tr::STAGES * (kTileM * 128 + Tr::NT * 128) + kEpiWarps * kStageBytes + kVecFloats * 4 + 2 * 2 * 128 * 8 +
         1024 + 256;
}

template <int MODE> static int launch_mode(dcb200_ctx* ctx, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p) {
  const size_t smem = smem_bytes<MODE>();
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&gemm_kernel<MODE>), smem));
  int grid = p.num_outer < ctx->sm_count ? p.num_outer : ctx->sm_count;
  ProfScope prof(ctx, K_HEAD2);
  gemm_kernel<MODE><<<grid, kThreads, smem, ctx->stream>>>(a, b, p);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_gemm(dcb200_ctx* ctx, int mode, const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p) {
  switch (mode) {
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

// fp32 row-major [rows][cols]; box = 128 rows x 32 columns (128 bytes), 128B swizzle: the epilogues' residual / h tiles
int make_tmap_2d_f32(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DCB200_ECUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d f32 rows=%llu) failed: %d", (unsigned long long)rows, (int)r);
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


// bf16 [B][C][L] (L contiguous); box = 64 (L) x 128 (C) x 1: 128 channel rows of 64 tokens (in_proj epilogue stores)
int make_tmap_3d_chbox(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DCB200_ECUDA;
  }
  cuuint64_t dims[3] = {L, C, B};
  cuuint64_t strides[2] = {L * 2, C * L * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d chbox) failed: %d", (int)r);
    return DCB200_ECUDA;
  }
  return DCB200_OK;
}

// bf16 [B][C][L] (L contiguous); box = 64 (L) x 1 (C) x 128 (B), 128B swizzle: a K-major operand tile whose
// rows are batch rows of one channel
int make_tmap_3d_rows(CUtensorMap* m, const void* base, uint64_t B, uint64_t C, uint64_t L) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return DCB200_ECUDA;
  }
  cuuint64_t dims[3] = {L, C, B};
  cuuint64_t strides[2] = {L * 2, C * L * 2};
  cuuint32_t box[3] = {64, 1, 128};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d rows) failed: %d", (int)r);
    return DCB200_ECUDA;
  }
  return DCB200_OK;
}

}  // namespace dcb
