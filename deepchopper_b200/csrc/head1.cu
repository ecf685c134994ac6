// First Linear of the classification head on CTA pairs (head.py:94-97):
//   r = relu(hf . Wh1^T + b1) + q      hf = ln_f output, bf16 [T,256];  Wh1 [1024,256];  q = normalised quality per token
//   -> bf16 [T,1024]
// Same shape as fc1 of the Hyena MLP, and the same machinery as block.cu: a cluster of two CTAs owns a 256-token tile,
// cta_group::2 MMAs (M = 256, N = 256) accumulate one 256-column group of r at a time into one of two TMEM stages, each
// CTA supplying its own 128 token rows of A and half of the weight rows of B.
//
// Why a pair: the single-CTA kernel (gemm_kernel<HEAD1>, removed) pulled 768 KB through its TMA ring per 128-token tile
// (all of Wh1, 512 KB, plus the A tile re-loaded for each of the four column groups) and ran at 19.6 k cycles per tile
// against 8.2 k of MMA -- neither its epilogue (rewritten with 16 warps and TMA stores: -2 %) nor HBM latency (an L2
// prefetch of the next A tile: +3 %) was the limit, the bytes per tile through L2 -> shared memory were.  Here a CTA
// loads half of the weights (256 KB per tile) and keeps its A tile resident (64 KB, refilled box by box as the tile's
// last group retires): 320 KB per 128 tokens.
//
//   warp 0  weight-ring producer (both CTAs)      warp 1  MMA issuer (leader CTA)      warp 2  TMEM allocator
//   warp 3  A-tile loader (both CTAs)             warps 4-19  epilogue: 4 TMEM lane quadrants x 4 parts of 64 columns;
//           the accumulator slice goes to registers and the stage straight back to the MMA warp; bias + relu + quality,
//           bf16 rows into a private 128B-swizzled 4 KB box, one TMA store per warp and group
#include "common.cuh"
#include "gemm.h"
#include "head1.h"
#include "ptx.cuh"

#include <string.h>

namespace dcb {

using namespace ptx;

namespace {

constexpr int kThreads = 640;               // 4 service warps + 16 epilogue warps
constexpr int kSlots = 5;                   // weight ring
constexpr uint32_t kUnitBytes = 128 * 128;  // 128 rows x 64 bf16
constexpr uint32_t kABytes = 4 * kUnitBytes;
constexpr uint32_t kBoxBytes = 32 * 128;    // epilogue staging box per warp: 32 token rows x 64 bf16 columns
constexpr int kGroups = 4;                  // 1024 output features / 256

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
head1_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmR, const Head1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_base = smem_u32(smem);
  const uint32_t w_base = a_base + kABytes;
  const uint32_t s_base = w_base + kSlots * kUnitBytes;  // 16 staging boxes
  float* vec = reinterpret_cast<float*>(smem + kABytes + kSlots * kUnitBytes + 16 * kBoxBytes);  // bias, 1024 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec + 1024);
  const uint32_t bar_base = smem_u32(bars);
  // "L" = only the leader's copy is used (waited on by the leader's MMA thread; the peer's threads and TMA loads signal
  // it through its shared::cluster address), "B" = both copies, signalled by multicast commits.
  enum { A_FULL = 0 /*L, one per K box*/, A_EMPTY = 4 /*B*/, W_FULL = 8 /*L*/, W_EMPTY = W_FULL + kSlots /*B*/,
         T_FULL = W_EMPTY + kSlots /*B, one per TMEM stage*/, T_EMPTY = T_FULL + 2 /*L*/, N_BARS = T_EMPTY + 2 };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  auto lbar = [&](int i) { return mapa(bar(i), 0); };  // the leader's copy (shared::cluster address)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmU);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar(A_FULL + s), 1);
      mbar_init(bar(A_EMPTY + s), 1);
    }
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar(W_FULL + s), 1);
      mbar_init(bar(W_EMPTY + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(T_FULL + s), 1);
      mbar_init(bar(T_EMPTY + s), 32);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 1024; i += kThreads) vec[i] = p.bias[i];
  if (warp == 2) {
    tmem_alloc_2sm(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int num_pairs = p.num_pairs;
  const int pair0 = (int)cluster_id_x(), pair_step = (int)cluster_nclusters_x();

  if (warp == 0) {
    // ===== weight-ring producer (both CTAs): my 128 rows of each 256-row group, K box by K box =====
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int pr = pair0; pr < num_pairs; pr += pair_step) {
        for (int g = 0; g < kGroups; ++g) {
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
            if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
            tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmW, lbar(W_FULL + slot), kb * 64, g * 256 + (int)rank * 128);
            if (++slot == kSlots) {
              slot = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader only): the whole warp runs the loop, the elected lane issues =====
    if (leader) {
      const uint32_t el = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, false, false);
      int slot = 0, acc = 0;
      uint32_t wphase = 0, acc_phase = 0, n = 0;
      for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
        for (int g = 0; g < kGroups; ++g) {
          mbar_wait_cluster(bar(T_EMPTY + acc), acc_phase ^ 1);  // the epilogue warps of both CTAs have drained the stage
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 256;
          for (int kb = 0; kb < 4; ++kb) {
            if (g == 0) mbar_wait(bar(A_FULL + kb), n & 1);
            mbar_wait(bar(W_FULL + slot), wphase);
            tc_fence_after();
            const uint32_t a_addr = a_base + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
            umma_bf16_x4_e<2>(d, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc, kb ? 1u : 0u, el);
            umma_commit_2sm_e(bar(W_EMPTY + slot), 3, el);
            if (g == kGroups - 1) umma_commit_2sm_e(bar(A_EMPTY + kb), 3, el);  // K box kb of the A tile may be refilled
            if (++slot == kSlots) {
              slot = 0;
              wphase ^= 1;
            }
          }
          umma_commit_2sm_e(bar(T_FULL + acc), 3, el);
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== A-tile loader (both CTAs): my 128 token rows, K box by K box, as the previous tile's last group retires =====
    if (lane == 0) {
      uint32_t n = 0;
      for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
        const int tok0 = pr * 256 + (int)rank * 128;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(A_EMPTY + kb), (n & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(bar(A_FULL + kb), 2 * kUnitBytes);
          tma_load_2d_2sm(a_base + kb * kUnitBytes, &tmU, lbar(A_FULL + kb), kb * 64, tok0);
        }
        // hf comes from HBM (the block kernel wrote 0.5 GB of it): the next tile's rows go to L2 a tile ahead
        if (pr + pair_step < num_pairs)
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_2d(&tmU, kb * 64, (pr + pair_step) * 256 + (int)rank * 128);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps = 4 TMEM lane quadrants x 4 column parts of 64 =====
    const int quad = warp & 3;
    const int part = (warp - 4) >> 2;
    const int row = quad * 32 + lane;  // my token row of this CTA's half tile
    const uint32_t box = s_base + (uint32_t)(warp - 4) * kBoxBytes;
    const uint32_t own = box + lane * 128;
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t t_empty0 = lbar(T_EMPTY), t_empty1 = lbar(T_EMPTY + 1);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t v[32], w2[32];
    for (int pr = pair0; pr < num_pairs; pr += pair_step) {
      const int tok0 = pr * 256 + (int)rank * 128;
      const int tok = tok0 + row;
      const float rowv = tok < p.T ? __ldg(p.qual + tok) : 0.f;  // quality of my token row
#pragma unroll 1
      for (int g = 0; g < kGroups; ++g) {
        mbar_wait(bar(T_FULL + acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + lane_off + acc * 256 + part * 64;
        tmem_ld32(t_row, v);
        tmem_ld32(t_row + 32, w2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(acc ? t_empty1 : t_empty0);  // my slice sits in registers
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        const int col0 = g * 256 + part * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&x)[32] = c ? w2 : v;
          const float4* b4 = reinterpret_cast<const float4*>(vec + col0 + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ba = b4[2 * q], bb = b4[2 * q + 1];
            const float bj[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float y0 = fmaxf(__uint_as_float(x[q * 8 + 2 * j]) + bj[2 * j], 0.f) + rowv;
              const float y1 = fmaxf(__uint_as_float(x[q * 8 + 2 * j + 1]) + bj[2 * j + 1], 0.f) + rowv;
              x[q * 4 + j] = pack_bf16(y0, y1);  // (in place: index q*4+j <= q*8+2j)
            }
          }
        }
        if (lane == 0) bulk_wait_read<0>();  // my previous store (one group ago) has read the box
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t (&x)[32] = c ? w2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(own + (((uint32_t)(c * 4 + q) ^ sw) << 4)),
                         "r"(x[q * 4]), "r"(x[q * 4 + 1]), "r"(x[q * 4 + 2]), "r"(x[q * 4 + 3])
                         : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmR, box, col0, tok0 + quad * 32);  // rows beyond T are clipped by the tensor map
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still be signalling my barriers / the leader's MMAs reading my smem until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

int launch_head1(dcb200_ctx* ctx, const CUtensorMap& tm_u, const CUtensorMap& tm_w, const CUtensorMap& tm_r,
                 const Head1Params& p) {
  const size_t smem = kABytes + kSlots * kUnitBytes + 16 * kBoxBytes + 1024 * 4 + 32 * 8 + 1024;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&head1_kernel), smem));
  int clusters = ctx->sm_count / 2;
  if (p.num_pairs < clusters) clusters = p.num_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(ctx, K_HEAD1);
  DCB_CUDA(cudaLaunchKernelEx(&cfg, head1_kernel, tm_u, tm_w, tm_r, p));
  ctx->launches++;
  return DCB200_OK;
}

}  // namespace dcb
