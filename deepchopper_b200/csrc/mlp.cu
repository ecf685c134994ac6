// Fused Hyena-block MLP (HF modeling_hyena.py HyenaMlp + residual + next LayerNorm):
//   h_out = fc2( gelu_tanh( fc1(m) + b1 ) ) + b2 + h_in ;  u = LN(h_out)   (next layer's LN1, or ln_f)
// The [T x 1024] hidden activation never leaves the SM: it is produced 128 columns at a time into a TMEM
// accumulator, GELU'd by the epilogue warps into a bf16 K-major tile in shared memory and consumed as the A operand
// of the second GEMM, whose [128 x 256] fp32 accumulator stays in TMEM for the whole tile.  HBM traffic per token:
// read m (512 B) + h_in (1 KB), write h_out (1 KB) + u (512 B); the unfused pair of GEMM kernels moved 7.6 KB.
//
// CTA pairs (cta_group::2): a cluster of two CTAs on one TPC owns 256 tokens; every tcgen05.mma is M = 256, with each
// CTA supplying its own 128 token rows of A and HALF of the weight tile's rows as B.  All weights (1 MB per layer)
// stream through shared memory once per token tile, so halving the per-SM weight bytes is what matters: a 96 KB ring
// can only keep ~50 GB/s per SM in flight against L2 latency, and one CTA alone would need twice that.
//
//   TMEM   [0,256) acc2 (fc2 accumulator)   [256,512) acc1 (fc1 accumulator of one 256-wide hidden group)
//   smem   m tile 64 KB | weight ring 6 x 16 KB | G 32 KB (gelu chunk = A of fc2) | 2 x 16 KB staging
//          The ring must cover the L2 round trip (~1 us) of the weight stream: three MMA groups in flight.  G is single-
//          buffered: GELU(j+1) is computed in registers while fc2(j) still reads G and written once fc2(j) retires,
//          which hides behind fc2(j) + fc1(j+2) on the tensor pipe.  Staging slots hold the h_out / u boxes on their way
//          to TMA stores; G doubles as two more slots once the tile's last fc2 has retired.  The fp32 residual never
//          waits on the critical path: its eight 32-column boxes stream through D0/D1 by TMA DURING the chunk loop and
//          are added straight into the fc2 accumulator in TMEM by two of the epilogue warp groups, in the window
//          between fc2(j-1) retiring and G(j) being published (the tensor pipe runs fc1(j+1) meanwhile).
//          Bias / LayerNorm vectors live in the kernel parameters (constant bank, warp-uniform reads): with 227 KB of
//          shared memory carved out there is no L1 left for __ldg.
//   warp 0  weight-ring producer (both CTAs; completion bytes are credited to the leader's barrier)
//   warp 1  MMA issuer (leader CTA only): fc1(0) | fc2(2P) fc1(P+1) fc2(2P+1) ...  all MMAs are M = 256, N = 256; the
//           hidden activation is consumed in 128-wide chunks j (K = 128 of fc2) through the single G buffer
//   warp 2  TMEM allocator
//   warp 3  m-tile loader + early residual loads (as soon as the G buffers of a tile are dead)
//   warps 4-19 epilogue: thread = accumulator row, four warps per TMEM lane quadrant split the columns (the GELU of
//           128 x 1024 values per tile is MUFU- and latency-bound: 16 warps keep the SFU pipes fed)
#include "common.cuh"
#include "gemm.h"
#include "mlp.h"
#include "ptx.cuh"

#include <string.h>

namespace dcb {

using namespace ptx;

namespace {

constexpr int kThreads = 640;  // 4 service warps + 16 epilogue warps
constexpr int kSlots = 6;
constexpr uint32_t kUnitBytes = 128 * 128;  // 128 rows x 64 bf16 (or 2 x 64 rows x 64 bf16)
constexpr uint32_t kABytes = 4 * kUnitBytes;
constexpr uint32_t kGBytes = 2 * kUnitBytes;  // one gelu chunk: 128 rows x 128 k
constexpr int kChunks = 8;                  // 1024 hidden / 128

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

__global__ void __launch_bounds__(kThreads, 1)
mlp_kernel(const __grid_constant__ CUtensorMap tmM, const __grid_constant__ CUtensorMap tmW1,
           const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmHin,
           const __grid_constant__ CUtensorMap tmHout, const __grid_constant__ CUtensorMap tmU, const MlpParams p) {
  // The 224 KB of operand tiles leave no room for alignment slack: the dynamic window itself must be 1024-byte
  // aligned (it is when the kernel has no static shared memory); trap loudly otherwise.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();
  const uint32_t a_base = smem_u32(smem);
  const uint32_t w_base = a_base + kABytes;
  const uint32_t g_base = w_base + kSlots * kUnitBytes;  // G (also staging slots s0, s1 at the end of a tile)
  const uint32_t r_base = g_base + kGBytes;              // dedicated staging slots D0, D1
  uint8_t* tail = smem + kABytes + kSlots * kUnitBytes + kGBytes + 2 * kUnitBytes;
  float2* stats = reinterpret_cast<float2*>(tail);  // [2 part pairs][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * 128 * 8);
  const uint32_t bar_base = smem_u32(bars);
  // Barriers.  "L" = only the leader's copy is used (waited on by the leader's MMA thread; the peer's threads and
  // TMA loads signal it through its shared::cluster address), "B" = both copies, signalled by multicast commits.
  enum { A_FULL = 0 /*L*/, A_EMPTY = 1 /*B*/, W_FULL = 2 /*L*/, W_EMPTY = W_FULL + kSlots /*B*/,
         T1_FULL = W_EMPTY + kSlots /*B*/, T1_EMPTY = T1_FULL + 1 /*L*/, G_FULL = T1_EMPTY + 1 /*L*/,
         G_EMPTY = G_FULL + 1 /*B*/, T2_FULL = G_EMPTY + 1 /*B*/, T2_EMPTY = T2_FULL + 1 /*L*/, R_FULL = T2_EMPTY + 1 /*local: D0, D1*/,
         N_BARS = R_FULL + 2 };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  auto lbar = [&](int i) { return mapa(bar(i), 0); };  // the leader's copy (shared::cluster address)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmM);
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
    for (int s = 0; s < 2; ++s) mbar_init(bar(R_FULL + s), 1);
    fence_barrier_init();
  }
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

  if (warp == 0) {
    // ===== weight-ring producer (both CTAs) =====
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      Tracer tr{traced ? p.trace + 2 * 2 * kTraceCap : nullptr, 0};
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
      // fc2 chunk j: my 128 rows (output features) of W2 x K = 128 -> two slots of 64 k
      auto load_fc2 = [&](int j) {
        for (int kb = 0; kb < 2; ++kb) {
          tr(320 + j);
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          tr(330 + j);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmW2, wfull[slot], j * 128 + kb * 64, (int)rank * 128);
          advance();
        }
      };
      for (int pr = pair0; pr < num_pairs; pr += pair_step) {
        load_fc1(0);
        for (int P = 0; P < kChunks / 2; ++P) {
          load_fc2(2 * P);
          if (P + 1 < kChunks / 2) load_fc1(P + 1);
          load_fc2(2 * P + 1);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader only) =====
    if (lane == 0 && leader) {
      constexpr uint32_t idesc2 = make_idesc_bf16(256, 256, false, false);
      int slot = 0;
      uint32_t wphase = 0;
      uint32_t n = 0;  // tile pairs done by this cluster
      Tracer tr{traced ? p.trace : nullptr, 0};
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
          mbar_wait_cluster(bar(T1_EMPTY), (P & 1) ^ 1);  // (4 uses per tile: parity of 4 n + P - 1)
          tr(110 + P);
          tc_fence_after();
          const uint32_t d = tmem_base + 256;
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tr(120 + P);
            tc_fence_after();
            const uint32_t a_addr = a_base + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
            umma_bf16_x4<2>(d, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc2, kb ? 1u : 0u);
            umma_commit_2sm(bar(W_EMPTY + slot), 3);
            advance();
          }
          umma_commit_2sm(bar(T1_FULL), 3);
          if (P == kChunks / 2 - 1) umma_commit_2sm(bar(A_EMPTY), 3);
        };
        auto fc2 = [&](int j) {
          tr(200 + j);
          mbar_wait_cluster(bar(G_FULL), j & 1);  // (8 uses per tile: parity of 8 n + j)
          tr(210 + j);
          if (j == 0) mbar_wait_cluster(bar(T2_EMPTY), (n & 1) ^ 1);
          tr(220 + j);
          tc_fence_after();
          for (int kb = 0; kb < 2; ++kb) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tr(230 + j);
            tc_fence_after();
            const uint32_t a_addr = g_base + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
            umma_bf16_x4<2>(tmem_base, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc2,
                            (j | kb) ? 1u : 0u);
            umma_commit_2sm(bar(W_EMPTY + slot), 3);
            advance();
          }
          umma_commit_2sm(bar(G_EMPTY), 3);
          if (j == kChunks - 1) umma_commit_2sm(bar(T2_FULL), 3);
        };
        tr(90);
        mbar_wait(bar(A_FULL), n & 1);
        tr(91);
        tc_fence_after();
        fc1(0);
        for (int P = 0; P < kChunks / 2; ++P) {
          fc2(2 * P);
          if (P + 1 < kChunks / 2) fc1(P + 1);  // acc1 is free once the epilogue holds its second half in registers
          fc2(2 * P + 1);
        }
      }
    }
  } else if (warp == 3) {
    // ===== m-tile loader (one tile ahead) + L2 prefetch of the tile's residual =====
    if (lane == 0) {
      const uint32_t afull = lbar(A_FULL);
      auto load_m = [&](int pr) {
        const int tok0 = pr * 256 + (int)rank * 128;
        if (leader) mbar_arrive_expect_tx(bar(A_FULL), 2 * kABytes);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_2sm(a_base + kb * kUnitBytes, &tmM, afull, kb * 64, tok0);
        for (int kb = 0; kb < 8; ++kb) tma_prefetch_2d(&tmHin, kb * 32, tok0);
        if (pr + pair_step < num_pairs)  // the tile after this one: its m load (issued late, at fc1(7)) will then hit L2
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_2d(&tmM, kb * 64, tok0 + pair_step * 256);
      };
      if (pair0 < num_pairs) load_m(pair0);
      uint32_t n = 0;
      for (int pr = pair0; pr + pair_step < num_pairs; pr += pair_step, ++n) {
        mbar_wait(bar(A_EMPTY), n & 1);  // fc1(7) of tile n has retired: the m buffer is free for tile n+1
        load_m(pr + pair_step);
      }
    }
  } else if (warp >= 4) {
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
    Tracer tr{(traced && warp == 4 && lane == 0) ? p.trace + 2 * kTraceCap : nullptr, 0};
    const uint64_t kC0 = f2_pack(0.7978845608f, 0.7978845608f), kC1 = f2_pack(0.0356774081f, 0.0356774081f);
    const uint64_t kHalf = f2_pack(0.5f, 0.5f);
    // Final-epilogue staging: part p owns the columns [64p, 64p+64) and ONE 16 KB slot (parts 0,1 the dedicated D0,D1,
    // parts 2,3 the halves of G, dead once the tile's last fc2 has retired) through which its two fp32 h_out boxes and
    // its bf16 u box go to TMA stores.
    const uint32_t my_slot = part < 2 ? r_base + part * kUnitBytes : g_base + (part - 2) * kUnitBytes;
    // Residual injection (parts 0,1 only): part p adds the boxes p, p+2, p+4, p+6 (32 fp32 columns each) through D[p].
    const uint32_t my_rfull = bar(R_FULL + (part & 1));
    uint32_t rphase = 0;
    for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
      const int tok0 = pr * 256 + (int)rank * 128;
      if (storer && part < 2) {  // first residual box of the tile -> D[part] (free since the end-of-tile barrier)
        mbar_arrive_expect_tx(my_rfull, kUnitBytes);
        tma_load_2d(my_slot, &tmHin, my_rfull, part * 32, tok0);
      }
      // ---- GELU chunks: acc1[s] -> bf16 K-major tile G[s]; my 32 of the chunk's 128 columns ---------------------
      for (int j = 0; j < kChunks; ++j) {
        const int s = j & 1;  // which 128-column half of acc1
        tr(400 + j);
        if (s == 0) mbar_wait(bar(T1_FULL), (j >> 1) & 1);  // fc1 group j/2 (parity of 4 n + j/2)
        tr(410 + j);
        tc_fence_after();
        tmem_ld32(tmem_base + lane_off + 256 + 128 * s + part * 32, v);
        const float* b1 = p.b1 + j * 128 + part * 32;  // kernel-parameter (constant bank) array, warp-uniform index
        tmem_ld_wait();
        // acc1 is free as soon as its second half sits in registers: fc1 of the next group may start
        tc_fence_before();
        __syncwarp();
        if (s == 1 && lane == 0) mbar_arrive_remote(t1_empty);
        const uint32_t grow = g_base + (part >> 1) * kUnitBytes + row * 128;
        uint32_t o[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float bj[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) bj[i] = b1[q * 8 + i];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // gelu_tanh on a pair: 0.5 x (1 + tanh(x (c0 + c1 x^2))), packed fp32x2 arithmetic
            const uint64_t x = f2_add(f2_pack(__uint_as_float(v[q * 8 + 2 * i]), __uint_as_float(v[q * 8 + 2 * i + 1])),
                                      f2_pack(bj[2 * i], bj[2 * i + 1]));
            const uint64_t in = f2_fma(f2_mul(x, x), kC1, kC0);
            float u0, u1, t0, t1;
            f2_unpack(f2_mul(x, in), u0, u1);
            asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
            const uint64_t hx = f2_mul(x, kHalf);
            float y0, y1;
            f2_unpack(f2_fma(hx, f2_pack(t0, t1), hx), y0, y1);
            o[q * 4 + i] = pack_bf16(y0, y1);
          }
        }
        tr(420 + j);
        mbar_wait(bar(G_EMPTY), (j & 1) ^ 1);  // fc2 of the previous chunk has finished reading G (parity of 8 n + j - 1)
        tr(430 + j);
        if (part < 2 && j >= 1 && j <= 4) {
          // acc2 is quiescent (fc2(j-1) retired, fc2(j) waits for my G_FULL arrival): acc2[:, box] += residual box
          const int col0 = (2 * (j - 1) + part) * 32;
          uint32_t a[32];
          mbar_wait(my_rfull, rphase);
          rphase ^= 1;
          tmem_ld32(tmem_base + lane_off + col0, a);
          tmem_ld_wait();
          const uint32_t rrow = my_slot + row * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            uint32_t r[4];
            lds128(rrow + (((uint32_t)q ^ sw) << 4), r);
#pragma unroll
            for (int i = 0; i < 4; ++i) a[4 * q + i] = __float_as_uint(__uint_as_float(a[4 * q + i]) + __uint_as_float(r[i]));
          }
          tmem_st32(tmem_base + lane_off + col0, a);
          tmem_st_wait();
          tc_fence_before();
          if (j < 4) {  // refill D[part] with my next box
            bar_sync(2 + part, 128);
            if (storer) {
              mbar_arrive_expect_tx(my_rfull, kUnitBytes);
              tma_load_2d(my_slot, &tmHin, my_rfull, col0 + 64, tok0);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(grow + (((uint32_t)((part & 1) * 4 + q) ^ sw) << 4), o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
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
      const int colA = part * 64;
      tmem_ld32(tmem_base + lane_off + colA, v);
      tmem_ld32(tmem_base + lane_off + colA + 32, v2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(t2_empty);  // acc2 drained
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xa = __uint_as_float(v[i]) + p.b2[colA + i];
        const float xb = __uint_as_float(v2[i]) + p.b2[colA + 32 + i];
        sum += xa + xb;
        sq = fmaf(xa, xa, fmaf(xb, xb, sq));
        v[i] = __float_as_uint(xa);
        v2[i] = __float_as_uint(xb);
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
      const uint32_t own = my_slot + row * 128;  // my row of the 128-row box, 16-byte chunks XOR-swizzled by (row & 7)
#pragma unroll
      for (int st = 0; st < 2; ++st) {
        const uint32_t (&x)[32] = st ? v2 : v;
        if (st == 1) bar_sync(2 + part, 128);  // (the storer arrives after the first store has read the slot)
#pragma unroll
        for (int q = 0; q < 8; ++q) sts128(own + (((uint32_t)q ^ sw) << 4), x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        fence_proxy_async();
        bar_sync(2 + part, 128);
        if (storer) {
          tma_store_2d(&tmHout, my_slot, colA + st * 32, tok0);
          bulk_commit();
        }
        if (st == 0) {
          bar_sync(6 + quad, 128);  // pair sums of all four parts are in place
        }
        if (storer) bulk_wait_read<0>();
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
      {  // u: my 64 bf16 columns (128 B per row), normalised in registers while the second h_out store drains
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&x)[32] = c ? v2 : v;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = colA + c * 32 + q * 8;
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf((__uint_as_float(x[q * 8 + i]) - mean) * rstd, p.ln_g[col + i], p.ln_b[col + i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[q * 4 + i] = pack_bf16(y[2 * i], y[2 * i + 1]);  // (in place: index q*4+i <= q*8+2i)
          }
        }
        bar_sync(2 + part, 128);  // slot free again
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
          bulk_wait_read<0>();
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

int launch_mlp(dcb200_ctx* ctx, const CUtensorMap& tm_m, const CUtensorMap& tm_w1, const CUtensorMap& tm_w2,
               const CUtensorMap& tm_hin, const CUtensorMap& tm_hout, const CUtensorMap& tm_u, const MlpParams& p) {
  const size_t smem = kABytes + kSlots * kUnitBytes + kGBytes + 2 * kUnitBytes + 2 * 128 * 8 + 40 * 8;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&mlp_kernel), smem));
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
  ProfScope prof(ctx, K_MLP);
  DCB_CUDA(cudaLaunchKernelEx(&cfg, mlp_kernel, tm_m, tm_w1, tm_w2, tm_hin, tm_hout, tm_u, p));
  ctx->launches++;
  return DCB200_OK;
}

}  // namespace dcb
