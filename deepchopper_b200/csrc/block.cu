// One Hyena block tail in ONE kernel (HF modeling_hyena.py HyenaBlock.forward after the long conv):
//   h1 = out_linear(y) + h0 ;  m = LN2(h1) ;  h2 = fc2( gelu_tanh( fc1(m) ) ) + h1 ;  u = LN(h2)  (next norm1 / ln_f)
// i.e. gemm_kernel<OUTPROJ> + block_kernel fused: h1 and m never leave the SM (out_proj alone moved 3 KB per token at
// 62-73 % of HBM peak and was the second-largest kernel).  HBM traffic per token: read y (512 B) + h0 (1 KB), write
// h2 (1 KB) + u (512 B).
//
// Structure = mlp.cu (CTA pairs, cta_group::2, M = 256, N = 256 MMAs; see there) plus a per-tile prologue:
//   * the y tile (channel-major, read MN-major by UMMA) is TMA-loaded into the buffer that will hold m,
//   * out_proj: acc1 = y . Wo^T (one N = 256 series; Wo streams through the weight ring),
//   * epilogue O (all 16 epilogue warps): x1 = acc1 + bo + h0 (residual boxes TMA-staged through the four staging
//     slots, two rounds), LN2 statistics, m -> bf16 K-major A tile written over the y tile, acc2 := x1 + b2 by
//     tcgen05.st, so that fc2 simply accumulates onto the residual,
//   * then the MLP chunk loop and final epilogue of mlp.cu (without its residual injection).
#include "common.cuh"
#include "gemm.h"
#include "block.h"
#include "ptx.cuh"

#include <string.h>

namespace dcb {

using namespace ptx;

namespace {

constexpr int kThreads = 640;  // 4 service warps + 16 epilogue warps
constexpr int kSlots = 5;  // weight ring (80 KB); the sixth slot's room holds the bias / LayerNorm vectors
// fp32 vectors in shared memory, read with warp-uniform (broadcast) LDS.128: an indexed constant-bank load (LDC.64)
// issues at a few cycles per warp, and at 640 values per thread per tile that was ~14 % of the kernel.
enum { V_BO = 0, V_B2 = 256, V_B1 = 512, V_FLOATS = 1536 };
constexpr uint32_t kUnitBytes = 128 * 128;  // 128 rows x 64 bf16 (or 2 x 64 rows x 64 bf16)
constexpr uint32_t kABytes = 4 * kUnitBytes;
constexpr uint32_t kGBytes = 4 * kUnitBytes;  // one gelu chunk: 128 rows x 256 k (four 64-wide K boxes)
constexpr int kChunks = 4;                  // 1024 hidden / 256: one chunk = one fc1 group

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float gelu_tanh(float x) {
  const float u = x * fmaf(0.0356774081f, x * x, 0.7978845608f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&w)[4]) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

}  // namespace

template <bool kTrace>
__global__ void __launch_bounds__(kThreads, 1)
block_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWo,
           const __grid_constant__ CUtensorMap tmW1,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmHin,
           const __grid_constant__ CUtensorMap tmHout, const __grid_constant__ CUtensorMap tmU, const BlockParams p) {
  // The 224 KB of operand tiles leave no room for alignment slack: the dynamic window itself must be 1024-byte
  // aligned (it is when the kernel has no static shared memory); trap loudly otherwise.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();
  const uint32_t a_base = smem_u32(smem);
  const uint32_t w_base = a_base + kABytes;
  const uint32_t g_base = w_base + kSlots * kUnitBytes;  // G: K box `part` doubles as part's staging slot between tiles
  const uint32_t vec_base = g_base + kGBytes;
  float* vecs = reinterpret_cast<float*>(smem + kABytes + kSlots * kUnitBytes + kGBytes);
  uint8_t* tail = smem + kABytes + kSlots * kUnitBytes + kGBytes + V_FLOATS * 4;
  float2* stats = reinterpret_cast<float2*>(tail);  // [2 part pairs][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * 128 * 8);
  const uint32_t bar_base = smem_u32(bars);
  // Barriers.  "L" = only the leader's copy is used (waited on by the leader's MMA thread; the peer's threads and
  // TMA loads signal it through its shared::cluster address), "B" = both copies, signalled by multicast commits.
  enum { A_FULL = 0 /*L*/, A_EMPTY = 1 /*B*/, W_FULL = 2 /*L*/, W_EMPTY = W_FULL + kSlots /*B*/,
         T1_FULL = W_EMPTY + kSlots /*B*/, T1_EMPTY = T1_FULL + 1 /*L*/, G_FULL = T1_EMPTY + 1 /*L*/,
         G_EMPTY = G_FULL + 1 /*B*/, T2_FULL = G_EMPTY + 1 /*B*/, T2_EMPTY = T2_FULL + 1 /*L*/,
         R_FULL = T2_EMPTY + 1 /*local: one per staging slot*/, R2_FULL = R_FULL + 4 /*local: one per K box of A*/,
         M_FULL = R2_FULL + 4 /*L*/, Y_DEAD = M_FULL + 1 /*B*/, N_BARS = Y_DEAD + 1 };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  auto lbar = [&](int i) { return mapa(bar(i), 0); };  // the leader's copy (shared::cluster address)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmWo);
    prefetch_tmap(&tmW1);
    prefetch_tmap(&tmW2);
    prefetch_tmap(&tmHin);
    prefetch_tmap(&tmHout);
    prefetch_tmap(&tmU);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bar(A_FULL), 1);
    mbar_init(bar(A_EMPTY), 1);
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(bar(W_FULL + s), 1);
      mbar_init(bar(W_EMPTY + s), 1);
    }
    mbar_init(bar(T1_FULL), 1);
    mbar_init(bar(T1_EMPTY), 32);  // one arrival per epilogue warp of both CTAs
    mbar_init(bar(G_FULL), 32);
    mbar_init(bar(G_EMPTY), 1);
    mbar_init(bar(T2_FULL), 1);
    mbar_init(bar(T2_EMPTY), 32);
    for (int s = 0; s < 8; ++s) mbar_init(bar(R_FULL + s), 1);
    mbar_init(bar(M_FULL), 32);
    mbar_init(bar(Y_DEAD), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += kThreads) {
    vecs[V_BO + i] = p.bo[i];
    vecs[V_B2 + i] = p.b2[i];
  }
  for (int i = threadIdx.x; i < 1024; i += kThreads) vecs[V_B1 + i] = p.b1[i];
  auto vec4 = [&](int idx, float (&o)[4]) {  // four consecutive vector elements, same address in every lane
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(vec_base + 4u * idx));
  };
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
  const bool traced = p.trace != nullptr && blockIdx.x == 0;
  const int pair0 = (int)cluster_id_x(), pair_step = (int)cluster_nclusters_x();

  // Register budget: the four service warps (one lane each of scalar code) give registers up so that the 512 epilogue
  // threads can hold a 64-column accumulator slice plus the LayerNorm state without spilling (with 227 KB of shared
  // memory there is no L1 to absorb local-memory traffic).  128 x 64 + 512 x 104 <= 640 x 96.
  if (warp < 4) {
  setmaxnreg_dec<64>();
  if (warp == 0) {
    // ===== weight-ring producer (both CTAs) =====
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      TracerT<kTrace> tr{traced ? p.trace + 2 * 2 * kTraceCap : nullptr, 0};
      uint32_t wfull[kSlots];
      for (int s = 0; s < kSlots; ++s) wfull[s] = lbar(W_FULL + s);
      auto advance = [&]() {
        if (++slot == kSlots) {
          slot = 0;
          phase ^= 1;
        }
      };
      // fc1 group P (256 hidden units): my 128 rows of W1 x K = 256 -> four slots of 64 k
      auto load_fc1 = [&](int P) {
        for (int kb = 0; kb < 4; ++kb) {
          tr(300 + P);
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          tr(310 + P);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmW1, wfull[slot], kb * 64, P * 256 + (int)rank * 128);
          advance();
        }
      };
      // fc2 chunk j: my 128 rows (output features) of W2 x K = 256 -> four slots of 64 k
      auto load_fc2 = [&](int j) {
        for (int kb = 0; kb < 4; ++kb) {
          tr(320 + j);
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          tr(330 + j);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmW2, wfull[slot], j * 256 + kb * 64, (int)rank * 128);
          advance();
        }
      };
      // out_proj: my 128 rows (output features) of Wo x K = 256 -> four slots of 64 k
      auto load_wo = [&]() {
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmWo, wfull[slot], kc * 64, (int)rank * 128);
          advance();
        }
      };
      for (int pr = pair0; pr < num_pairs; pr += pair_step) {
        load_wo();  // (the order of the MMA warp)
        load_fc1(0);
        for (int P = 0; P < kChunks; ++P) {
          if (P + 1 < kChunks) load_fc1(P + 1);
          load_fc2(P);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader only): the whole warp runs the loop, the elected lane issues (ptx.cuh umma_bf16_x4_e) =====
    if (leader) {
      const uint32_t el = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc2 = make_idesc_bf16(256, 256, false, false);
      int slot = 0;
      uint32_t wphase = 0;
      uint32_t n = 0;   // tile pairs done by this cluster
      uint32_t t1 = 0;  // uses of acc1 started so far (out_proj + 4 fc1 groups per tile)
      constexpr uint32_t idesc_o = make_idesc_bf16(256, 256, true, false);  // A = y tile, MN-major
      TracerT<kTrace> tr{(traced && el) ? p.trace : nullptr, 0};
      auto advance = [&]() {
        if (++slot == kSlots) {
          slot = 0;
          wphase ^= 1;
        }
      };
      for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
        // fc1 group P: acc1 (256 columns, single stage) = m . W1[256 P .. 256 P + 255]^T, one N = 256 MMA series.
        // (N = 128 MMAs re-read the 4 KB A slice per 2 KB of B and ran at half the N = 256 rate: the tensor pipe's
        // operand fetch from shared memory, ~64 B/clk, is what bounds these shapes.)
        auto fc1 = [&](int P) {
          tr(100 + P);
          mbar_wait_cluster(bar(T1_EMPTY), (t1 & 1) ^ 1);  // the epilogue has drained the previous use of acc1
          ++t1;
          tr(110 + P);
          tc_fence_after();
          const uint32_t d = tmem_base + 256;
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tr(120 + P);
            tc_fence_after();
            const uint32_t a_addr = a_base + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
            umma_bf16_x4_e<2>(d, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc2, kb ? 1u : 0u, el);
            umma_commit_2sm_e(bar(W_EMPTY + slot), 3, el);
            advance();
          }
          umma_commit_2sm_e(bar(T1_FULL), 3, el);
          if (P == kChunks - 1) umma_commit_2sm_e(bar(A_EMPTY), 3, el);
        };
        auto fc2 = [&](int j) {
          tr(200 + j);
          mbar_wait_cluster(bar(G_FULL), j & 1);  // (4 uses per tile: parity of 4 n + j)
          tr(210 + j);
          if (j == 0) mbar_wait_cluster(bar(T2_EMPTY), (n & 1) ^ 1);
          tr(220 + j);
          tc_fence_after();
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tr(230 + j);
            tc_fence_after();
            const uint32_t a_addr = g_base + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
            umma_bf16_x4_e<2>(tmem_base, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc2,
                              1u, el);  // acc2 was preloaded with x1 + b2 by epilogue O
            umma_commit_2sm_e(bar(W_EMPTY + slot), 3, el);
            advance();
          }
          umma_commit_2sm_e(bar(G_EMPTY), 3, el);
          if (j == kChunks - 1) umma_commit_2sm_e(bar(T2_FULL), 3, el);
        };
        tr(90);
        mbar_wait(bar(A_FULL), n & 1);
        tr(91);
        tc_fence_after();
        // out_proj: acc1 = y . Wo^T
        mbar_wait_cluster(bar(T1_EMPTY), (t1 & 1) ^ 1);
        ++t1;
        tc_fence_after();
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(bar(W_FULL + slot), wphase);
          tc_fence_after();
          const uint32_t a_addr = a_base + kc * kUnitBytes;  // [tokens 0-63 | tokens 64-127] x 64 channels, 8 KB each
          const uint32_t b_addr = w_base + slot * kUnitBytes;
          umma_bf16_x4_e<2>(tmem_base + 256, make_desc_sw128(a_addr, 8192, 1024), 128, make_desc_sw128(b_addr, 16, 1024), 2,
                            idesc_o, kc ? 1u : 0u, el);
          umma_commit_2sm_e(bar(W_EMPTY + slot), 3, el);
          advance();
        }
        umma_commit_2sm_e(bar(T1_FULL), 3, el);
        umma_commit_2sm_e(bar(Y_DEAD), 3, el);
        tr(92);
        mbar_wait_cluster(bar(M_FULL), n & 1);  // m tile written over the y tile, acc2 preloaded (both CTAs)
        tr(93);
        tc_fence_after();
        // fc1(P + 1) goes ahead of fc2(P): acc1 is free as soon as the epilogue holds group P in registers, G(P) only
        // after its GELU, so the tensor pipe works on fc1(P + 1) while the epilogue warps compute GELU(P).
        fc1(0);
        for (int P = 0; P < kChunks; ++P) {
          if (P + 1 < kChunks) fc1(P + 1);
          fc2(P);
        }
      }
    }
  } else if (warp == 3) {
    // ===== y-tile loader (one tile ahead) + L2 prefetch of the tile's residual rows =====
    // (Prefetching y and the residual a whole tile earlier was tried: ~142 MB flow through L2 per round of tiles, so
    // 60 % of the prefetched lines were evicted before use and DRAM reads grew by half for no gain in time.)
    if (lane == 0) {
      const uint32_t afull = lbar(A_FULL);
      auto load_y = [&](int pr) {
        const int tok0 = pr * 256 + (int)rank * 128;
        const int b = tok0 / p.L, l0 = tok0 % p.L;
        if (leader) mbar_arrive_expect_tx(bar(A_FULL), 2 * kABytes);
        for (int kc = 0; kc < 4; ++kc) {
          tma_load_3d_2sm(a_base + kc * kUnitBytes, &tmY, afull, l0, kc * 64, b);
          tma_load_3d_2sm(a_base + kc * kUnitBytes + 8192, &tmY, afull, l0 + 64, kc * 64, b);
        }
        for (int kb = 0; kb < 8; ++kb) tma_prefetch_2d(&tmHin, kb * 32, tok0);
      };
      if (pair0 < num_pairs) load_y(pair0);
      uint32_t n = 0;
      for (int pr = pair0; pr + pair_step < num_pairs; pr += pair_step, ++n) {
        mbar_wait(bar(A_EMPTY), n & 1);  // the last fc1 group of tile n has retired: the buffer is free for tile n+1
        load_y(pr + pair_step);
      }
    }
  }
  } else {
    setmaxnreg_inc<104>();
    // ===== epilogue: 16 warps = 4 TMEM lane quadrants x 4 column parts =====
    const int quad = warp & 3;
    const int part = (warp - 4) >> 2;  // 0..3
    const int row = quad * 32 + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool storer = (lane == 0 && quad == 0);  // one per part
    const uint32_t t1_empty = lbar(T1_EMPTY);
    const uint32_t g_full = lbar(G_FULL);
    const uint32_t t2_empty = lbar(T2_EMPTY);
    uint32_t n = 0;
    uint32_t v[32];
    TracerT<kTrace> tr{(traced && warp == 4 && lane == 0) ? p.trace + 2 * kTraceCap : nullptr, 0};
    const uint64_t kC0 = f2_pack(0.7978845608f, 0.7978845608f), kC1 = f2_pack(0.0356774081f, 0.0356774081f);
    const uint64_t kHalf = f2_pack(0.5f, 0.5f);
    // Final-epilogue staging: part p owns the columns [64p, 64p+64) and ONE 16 KB slot (K box p of G, dead once the
    // tile's last fc2 has retired and until GELU(0) of the next tile) through which its two fp32 h_out boxes and its
    // bf16 u box go to TMA stores.
    const uint32_t my_slot = g_base + part * kUnitBytes;
    // Epilogue O takes the first residual box (columns [64 part, +32)) through the same slot (requested at the end of
    // the previous tile) and the second from the A buffer (warp 2).
    const uint32_t my_rfull = bar(R_FULL + part);
    const uint32_t m_full = lbar(M_FULL);
    uint32_t e1 = 0;  // uses of acc1 consumed so far
    const int colA = part * 64;
    if (storer && pair0 < num_pairs) {
      mbar_arrive_expect_tx(my_rfull, kUnitBytes);
      tma_load_2d(my_slot, &tmHin, my_rfull, colA, pair0 * 256 + (int)rank * 128);
    }
    for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
      const int tok0 = pr * 256 + (int)rank * 128;
      // ---- epilogue O: x1 = acc1 + bo + h0 ; m = LN2(x1) -> A tile of fc1 ; acc2 := x1 + b2 ---------------------------
      {
        uint32_t w2[32];
        tr(380);
        mbar_wait(bar(T1_FULL), e1 & 1);
        ++e1;
        tr(381);
        tc_fence_after();
        if (n == 0 && storer) {  // first tile: the out_proj MMAs have retired, the y buffer is free for my second box
          mbar_arrive_expect_tx(bar(R2_FULL + part), kUnitBytes);
          tma_load_2d(a_base + part * kUnitBytes, &tmHin, bar(R2_FULL + part), colA + 32, tok0);
        }
        tmem_ld32(tmem_base + lane_off + 256 + colA, v);
        tmem_ld32(tmem_base + lane_off + 256 + colA + 32, w2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(t1_empty);  // acc1 drained
        tr(382);
        uint64_t sum2 = f2_pack(0.f, 0.f), sq2 = sum2;  // (even, odd) column partial sums: packed fp32x2 arithmetic
        // Both boxes were requested by my part's storer thread at the end of the previous tile's final epilogue: the one
        // in the A buffer (my columns 32-63) first, the one in my staging slot (columns 0-31) second.
#pragma unroll
        for (int st = 0; st < 2; ++st) {
          uint32_t (&x)[32] = st ? v : w2;
          const uint32_t rrow = (st ? my_slot : a_base + part * kUnitBytes) + row * 128;
          const int c0 = colA + (st ? 0 : 32);
          mbar_wait(st ? my_rfull : bar(R2_FULL + part), n & 1);
          tr(383 + st);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t r[4];
            float c4[4];
            lds128(rrow + (((uint32_t)q ^ sw) << 4), r);
            vec4(V_BO + c0 + 4 * q, c4);
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
              const uint64_t xv = f2_add(f2_add(f2_pack(__uint_as_float(x[4 * q + i]), __uint_as_float(x[4 * q + i + 1])),
                                                f2_pack(c4[i], c4[i + 1])),
                                         f2_pack(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
              sum2 = f2_add(sum2, xv);
              sq2 = f2_fma(xv, xv, sq2);
              float x0, x1;
              f2_unpack(xv, x0, x1);
              x[4 * q + i] = __float_as_uint(x0);
              x[4 * q + i + 1] = __float_as_uint(x1);
            }
          }
        }
        float sum, sq;
        {
          float a0, a1, b0, b1;
          f2_unpack(sum2, a0, a1);
          f2_unpack(sq2, b0, b1);
          sum = a0 + a1;
          sq = b0 + b1;
        }
        tr(385);
        float2* st2 = stats + (part >> 1) * 128 + row;
        if (part & 1) *st2 = make_float2(sum, sq);
        bar_sync(6 + quad, 128);
        if (!(part & 1)) {
          const float2 o = *st2;
          *st2 = make_float2(sum + o.x, sq + o.y);
        }
        bar_sync(6 + quad, 128);
        tr(386);
        float2 sa = stats[row];
        {
          const float2 sb = stats[128 + row];
          sa.x += sb.x;
          sa.y += sb.y;
        }
        const float mean = sa.x * (1.0f / 256.0f);
        const float rstd = rsqrtf(fmaxf(sa.y * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-5f);
        // m: my 64 columns = K box `part` of the fc1 A tile (over the y tile: every out_proj MMA has retired);
        // acc2 := x1 + b2: fc2 accumulates onto the residual.  Packed fp32x2: z = x rstd - mean rstd ; m = z g + b.
        const uint32_t mrow = a_base + part * kUnitBytes + row * 128;
        const uint64_t rs2 = f2_pack(rstd, rstd), nm2 = f2_pack(-mean * rstd, -mean * rstd);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&x)[32] = c ? w2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = colA + c * 32 + q * 8;
            float d8[8];
            uint32_t o[4];
            vec4(V_B2 + col, *reinterpret_cast<float(*)[4]>(d8));
            vec4(V_B2 + col + 4, *reinterpret_cast<float(*)[4]>(d8 + 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint64_t xv = f2_pack(__uint_as_float(x[q * 8 + 2 * i]), __uint_as_float(x[q * 8 + 2 * i + 1]));
              float y0, y1, a0, a1;
              f2_unpack(f2_fma(xv, rs2, nm2), y0, y1);
              o[i] = pack_bf16(y0, y1);
              f2_unpack(f2_add(xv, f2_pack(d8[2 * i], d8[2 * i + 1])), a0, a1);
              x[q * 8 + 2 * i] = __float_as_uint(a0);
              x[q * 8 + 2 * i + 1] = __float_as_uint(a1);
            }
            sts128(mrow + (((uint32_t)(c * 4 + q) ^ sw) << 4), o[0], o[1], o[2], o[3]);
          }
        }
        tr(387);
        tmem_st32(tmem_base + lane_off + colA, v);
        tmem_st32(tmem_base + lane_off + colA + 32, w2);
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(m_full);
        tr(388);
        bar_sync(6 + quad, 128);  // (stats are rewritten by the final epilogue of this tile)
      }
      // ---- GELU chunks: acc1 (one fc1 group, 256 hidden units) -> bf16 K-major tile G; my 64 of the 256 columns = my
      // row of K box `part`.  acc1 goes back to the MMA warp as soon as the group sits in registers.
#pragma unroll 1
      for (int j = 0; j < kChunks; ++j) {
        uint32_t w2[32];
        tr(400 + j);
        mbar_wait(bar(T1_FULL), e1 & 1);
        ++e1;
        tr(410 + j);
        tc_fence_after();
        tmem_ld32(tmem_base + lane_off + 256 + colA, v);
        tmem_ld32(tmem_base + lane_off + 256 + colA + 32, w2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(t1_empty);
        const int b1o = V_B1 + j * 256 + colA;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&x)[32] = c ? w2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float bj[8];
            vec4(b1o + c * 32 + q * 8, *reinterpret_cast<float(*)[4]>(bj));
            vec4(b1o + c * 32 + q * 8 + 4, *reinterpret_cast<float(*)[4]>(bj + 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // gelu_tanh on a pair: 0.5 x (1 + tanh(x (c0 + c1 x^2))), packed fp32x2 arithmetic
              const uint64_t xx = f2_add(f2_pack(__uint_as_float(x[q * 8 + 2 * i]), __uint_as_float(x[q * 8 + 2 * i + 1])),
                                         f2_pack(bj[2 * i], bj[2 * i + 1]));
              const uint64_t in = f2_fma(f2_mul(xx, xx), kC1, kC0);
              float u0, u1, t0, t1;
              f2_unpack(f2_mul(xx, in), u0, u1);
              asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
              asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
              const uint64_t hx = f2_mul(xx, kHalf);
              float y0, y1;
              f2_unpack(f2_fma(hx, f2_pack(t0, t1), hx), y0, y1);
              x[q * 4 + i] = pack_bf16(y0, y1);  // (in place: index q*4+i <= q*8+2i)
            }
          }
        }
        tr(420 + j);
        mbar_wait(bar(G_EMPTY), (j & 1) ^ 1);  // fc2 of the previous chunk has finished reading G (parity of 4 n + j - 1)
        tr(430 + j);
        const uint32_t grow = my_slot + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t (&x)[32] = c ? w2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(grow + (((uint32_t)(c * 4 + q) ^ sw) << 4), x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(g_full);
        tr(440 + j);
      }
      // ---- final: x = acc2 + b2 (+ residual, already in acc2) -> h_out ; LayerNorm -> u ------------------------------
      // My 64 columns of x stay in registers from here on, so acc2 goes back to the MMA warp after ~1 k cycles (fc2(0) of
      // the next tile only waits for that and for GELU(0)).  Global stores are plain coalesced st.global: each warp
      // transposes its own 32 rows through a private 4 KB scratch (row-owner writes, 4 rows x 128 B reads).
      tr(500);
      mbar_wait(bar(T2_FULL), n & 1);
      tr(501);
      tc_fence_after();
      uint32_t v2[32];
      tmem_ld32(tmem_base + lane_off + colA, v);
      tmem_ld32(tmem_base + lane_off + colA + 32, v2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(t2_empty);  // acc2 drained
      float sum, sq;
      {
        uint64_t sum2 = f2_pack(0.f, 0.f), sq2 = sum2;  // (b2 and the residual are already in acc2)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const uint64_t xa = f2_pack(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          const uint64_t xb = f2_pack(__uint_as_float(v2[i]), __uint_as_float(v2[i + 1]));
          sum2 = f2_add(sum2, f2_add(xa, xb));
          sq2 = f2_fma(xa, xa, f2_fma(xb, xb, sq2));
        }
        float a0, a1, b0, b1;
        f2_unpack(sum2, a0, a1);
        f2_unpack(sq2, b0, b1);
        sum = a0 + a1;
        sq = b0 + b1;
      }
      // row statistics (sum, sum of squares) over the four column parts, in a fixed order: (p0 + p1) + (p2 + p3)
      float2* st2 = stats + (part >> 1) * 128 + row;
      if (part & 1) *st2 = make_float2(sum, sq);
      bar_sync(6 + quad, 128);  // the four warps that share these 32 rows
      if (!(part & 1)) {
        const float2 o = *st2;
        *st2 = make_float2(sum + o.x, sq + o.y);
      }
      tr(540);
      // Three boxes (h_out columns [64p, +32), [64p+32, +32) as fp32, u columns [64p, +64) as bf16; 16 KB each) go
      // through my part's slot to TMA stores: the LSU store path sustains only ~24 B/clk per SM here (its queue of
      // outstanding L2 writes is latency bound), TMA is not.  The slot is rewritten once the previous store has read it.
      // Three boxes per column part (h_out columns [64p, +32), [64p+32, +32) as fp32, u columns [64p, +64) as bf16; 16 KB
      // each) leave through TMA stores.  Two staging slots per part: my slot in G and K box p of the A buffer (dead once
      // the next tile's out_proj MMAs have retired).  With a single slot every box waited ~2.5 k cycles for the previous
      // store to finish reading it; now the third box reuses the first slot only after two stagings.
      const bool has_next = pr + pair_step < num_pairs;
      const uint32_t own = my_slot + row * 128;  // my row of a 128-row box, 16-byte chunks XOR-swizzled by (row & 7)
      const uint32_t y_slot = a_base + part * kUnitBytes;
      const uint32_t own_y = y_slot + row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) sts128(own + (((uint32_t)q ^ sw) << 4), v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      fence_proxy_async();
      bar_sync(2 + part, 128);
      if (storer) {
        tma_store_2d(&tmHout, my_slot, colA, tok0);
        bulk_commit();
      }
      bar_sync(6 + quad, 128);  // pair sums of all four parts are in place
      // the A buffer: free once out_proj of the NEXT tile has read the y tile that was loaded into it (after the last
      // tile of this CTA pair nothing is loaded, m is dead since the last fc1 group retired)
      if (has_next) mbar_wait(bar(Y_DEAD), (n + 1) & 1);
#pragma unroll
      for (int q = 0; q < 8; ++q) sts128(own_y + (((uint32_t)q ^ sw) << 4), v2[4 * q], v2[4 * q + 1], v2[4 * q + 2], v2[4 * q + 3]);
      fence_proxy_async();
      bar_sync(2 + part, 128);
      if (storer) {
        tma_store_2d(&tmHout, y_slot, colA + 32, tok0);
        bulk_commit();
      }
      tr(551);
      float2 sa = stats[row];
      {
        const float2 sb = stats[128 + row];
        sa.x += sb.x;
        sa.y += sb.y;
      }
      const float mean = sa.x * (1.0f / 256.0f);
      const float rstd = rsqrtf(fmaxf(sa.y * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-5f);
      const uint64_t rs2 = f2_pack(rstd, rstd), nm2 = f2_pack(-mean * rstd, -mean * rstd);
      {  // u: my 64 bf16 columns (128 B per row), normalised in registers while the h_out stores drain
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&x)[32] = c ? v2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float y0, y1;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(x[q * 8 + 2 * i]), __uint_as_float(x[q * 8 + 2 * i + 1])), rs2, nm2), y0, y1);
              x[q * 4 + i] = pack_bf16(y0, y1);  // (in place: index q*4+i <= q*8+2i)
            }
          }
        }
        if (storer) bulk_wait_read<1>();  // the first store (my slot) has been read; the second may still be reading
        bar_sync(2 + part, 128);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t (&x)[32] = c ? v2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(own + (((uint32_t)(c * 4 + q) ^ sw) << 4), x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
        }
        fence_proxy_async();
        bar_sync(2 + part, 128);
        if (storer) {
          tma_store_2d(&tmU, my_slot, colA, tok0);
          bulk_commit();
          if (has_next) {  // next tile's residual boxes: into each slot as soon as its store has been read
            const int ntok0 = (pr + pair_step) * 256 + (int)rank * 128;
            bulk_wait_read<1>();
            mbar_arrive_expect_tx(bar(R2_FULL + part), kUnitBytes);
            tma_load_2d(y_slot, &tmHin, bar(R2_FULL + part), colA + 32, ntok0);
            bulk_wait_read<0>();
            mbar_arrive_expect_tx(my_rfull, kUnitBytes);
            tma_load_2d(my_slot, &tmHin, my_rfull, colA, ntok0);
          } else {
            bulk_wait_read<0>();
          }
        }
      }
      tr(560);
      bar_sync(1, 512);  // every store has read its slot (G halves / D slots): the next tile may reuse them
    }
    if (storer) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still be signalling my barriers / the leader's MMAs reading my smem until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

int launch_block(dcb200_ctx* ctx, const CUtensorMap& tm_y, const CUtensorMap& tm_wo, const CUtensorMap& tm_w1,
                 const CUtensorMap& tm_w2, const CUtensorMap& tm_hin, const CUtensorMap& tm_hout, const CUtensorMap& tm_u,
                 const BlockParams& p) {
  const size_t smem = kABytes + kSlots * kUnitBytes + kGBytes + V_FLOATS * 4 + 2 * 128 * 8 + 56 * 8;
  auto kern = p.trace ? &block_kernel<true> : &block_kernel<false>;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(kern), smem));
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
  ProfScope prof(ctx, K_BLOCK);
  DCB_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_y, tm_wo, tm_w1, tm_w2, tm_hin, tm_hout, tm_u, p));
  ctx->launches++;
  return DCB200_OK;
}

}  // namespace dcb
