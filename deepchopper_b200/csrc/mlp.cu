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
//   TMEM   [0,256) acc2 (fc2 accumulator)   [256,384) [384,512) acc1 stages (fc1 chunk accumulators)
//   smem   m tile 64 KB | weight ring 4 x 16 KB | G 2 x 32 KB (gelu chunk = A of fc2) | 2 x 16 KB staging
//          staging slots hold TMA-loaded residual boxes that become TMA-stored h_out / u boxes in place; the G buffers
//          double as four more slots once the tile's last fc2 MMAs have retired
//   warp 0  weight-ring producer (both CTAs; completion bytes are credited to the leader's barrier)
//   warp 1  MMA issuer (leader CTA only): fc1(0) fc1(1) | fc2(j) fc1(j+2) ...  fc1 of the next chunk is queued before
//           fc2 so the tensor pipe stays busy while the epilogue warps GELU the current chunk
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
constexpr int kSlots = 4;
constexpr uint32_t kUnitBytes = 128 * 128;  // 128 rows x 64 bf16 (or 2 x 64 rows x 64 bf16)
constexpr uint32_t kABytes = 4 * kUnitBytes;
constexpr uint32_t kGBytes = 2 * kUnitBytes;
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
  const uint32_t g_base = w_base + kSlots * kUnitBytes;  // G[0] = slots 0,1 ; G[1] = slots 2,3
  const uint32_t r_base = g_base + 2 * kGBytes;          // dedicated staging slots D[0], D[1] (one per column half)
  uint8_t* tail = smem + kABytes + kSlots * kUnitBytes + 2 * kGBytes + 2 * kUnitBytes;
  float2* stats = reinterpret_cast<float2*>(tail);  // [2 part pairs][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 2 * 128 * 8);
  const uint32_t bar_base = smem_u32(bars);
  // Barriers.  "L" = only the leader's copy is used (waited on by the leader's MMA thread; the peer's threads and
  // TMA loads signal it through its shared::cluster address), "B" = both copies, signalled by multicast commits.
  enum { A_FULL = 0 /*L*/, A_EMPTY = 1 /*B*/, W_FULL = 2 /*L*/, W_EMPTY = W_FULL + kSlots /*B*/,
         T1_FULL = W_EMPTY + kSlots /*B*/, T1_EMPTY = T1_FULL + 2 /*L*/, G_FULL = T1_EMPTY + 2 /*L*/,
         G_EMPTY = G_FULL + 2 /*B*/, T2_FULL = G_EMPTY + 2 /*B*/, T2_EMPTY = T2_FULL + 1 /*L*/,
         R_FULL = T2_EMPTY + 1 /*local: s0..s3, D0a, D1a, D0b, D1b*/, GD0 = R_FULL + 8 /*B*/, N_BARS = GD0 + 1 };
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
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(T1_FULL + s), 1);
      mbar_init(bar(T1_EMPTY + s), 32);  // one arrival per epilogue warp of both CTAs
      mbar_init(bar(G_FULL + s), 32);
      mbar_init(bar(G_EMPTY + s), 1);
    }
    mbar_init(bar(T2_FULL), 1);
    mbar_init(bar(T2_EMPTY), 32);
    for (int s = 0; s < 8; ++s) mbar_init(bar(R_FULL + s), 1);
    mbar_init(bar(GD0), 1);
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
  const int pair0 = (int)cluster_id_x(), pair_step = (int)cluster_nclusters_x();

  if (warp == 0) {
    // ===== weight-ring producer (both CTAs) =====
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      uint32_t wfull[kSlots];
      for (int s = 0; s < kSlots; ++s) wfull[s] = lbar(W_FULL + s);
      auto advance = [&]() {
        if (++slot == kSlots) {
          slot = 0;
          phase ^= 1;
        }
      };
      // fc1 chunk j: my 64 rows (hidden units) of W1 x K = 256 -> two slots, each [kb even 8 KB][kb odd 8 KB]
      auto load_fc1 = [&](int j) {
        for (int h = 0; h < 2; ++h) {
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          const uint32_t dst = w_base + slot * kUnitBytes;
          tma_load_2d_2sm(dst, &tmW1, wfull[slot], (2 * h) * 64, j * 128 + (int)rank * 64);
          tma_load_2d_2sm(dst + 8192, &tmW1, wfull[slot], (2 * h + 1) * 64, j * 128 + (int)rank * 64);
          advance();
        }
      };
      // fc2 chunk j: my 128 rows (output features) of W2 x K = 128 -> two slots of 64 k
      auto load_fc2 = [&](int j) {
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(bar(W_EMPTY + slot), phase ^ 1);
          if (leader) mbar_arrive_expect_tx(bar(W_FULL + slot), 2 * kUnitBytes);
          tma_load_2d_2sm(w_base + slot * kUnitBytes, &tmW2, wfull[slot], j * 128 + kb * 64, (int)rank * 128);
          advance();
        }
      };
      for (int pr = pair0; pr < num_pairs; pr += pair_step) {
        load_fc1(0);
        load_fc1(1);
        for (int j = 0; j < kChunks; ++j) {
          load_fc2(j);
          if (j + 2 < kChunks) load_fc1(j + 2);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader only) =====
    if (lane == 0 && leader) {
      constexpr uint32_t idesc1 = make_idesc_bf16(256, 128, false, false);
      constexpr uint32_t idesc2 = make_idesc_bf16(256, 256, false, false);
      int slot = 0;
      uint32_t wphase = 0;
      uint32_t n = 0;  // tile pairs done by this cluster
      auto advance = [&]() {
        if (++slot == kSlots) {
          slot = 0;
          wphase ^= 1;
        }
      };
      for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
        // use counters: acc1 stage s / G buffer s are used by chunks j with j % 2 == s -> 4 uses per tile
        auto fc1 = [&](int j) {
          const int s = j & 1;
          const uint32_t use = 4 * n + (j >> 1);
          mbar_wait_cluster(bar(T1_EMPTY + s), (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + 256 + 128 * s;
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int kb = 2 * h + (kk >> 2), k = kk & 3;
              const uint32_t a_addr = a_base + kb * kUnitBytes + k * 32;
              const uint32_t b_addr = w_base + slot * kUnitBytes + (kk >> 2) * 8192 + k * 32;
              umma_bf16_2sm(d, make_desc_sw128(a_addr, 16, 1024), make_desc_sw128(b_addr, 16, 1024), idesc1,
                            (h | kk) ? 1u : 0u);
            }
            umma_commit_2sm(bar(W_EMPTY + slot), 3);
            advance();
          }
          umma_commit_2sm(bar(T1_FULL + s), 3);
          if (j == kChunks - 1) umma_commit_2sm(bar(A_EMPTY), 3);
        };
        auto fc2 = [&](int j) {
          const int s = j & 1;
          const uint32_t use = 4 * n + (j >> 1);
          mbar_wait_cluster(bar(G_FULL + s), use & 1);
          if (j == 0) mbar_wait_cluster(bar(T2_EMPTY), (n & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < 2; ++kb) {
            mbar_wait(bar(W_FULL + slot), wphase);
            tc_fence_after();
            const uint32_t a_addr = g_base + s * kGBytes + kb * kUnitBytes;
            const uint32_t b_addr = w_base + slot * kUnitBytes;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_2sm(tmem_base, make_desc_sw128(a_addr + k * 32, 16, 1024), make_desc_sw128(b_addr + k * 32, 16, 1024),
                            idesc2, (j | kb | k) ? 1u : 0u);
            umma_commit_2sm(bar(W_EMPTY + slot), 3);
            advance();
          }
          umma_commit_2sm(bar(G_EMPTY + s), 3);
          if (j == kChunks - 2) umma_commit_2sm(bar(GD0), 3);  // G[0] is dead for the rest of the tile
          if (j == kChunks - 1) umma_commit_2sm(bar(T2_FULL), 3);
        };
        mbar_wait(bar(A_FULL), n & 1);
        tc_fence_after();
        fc1(0);
        fc1(1);
        for (int j = 0; j < kChunks; ++j) {
          fc2(j);
          if (j + 2 < kChunks) fc1(j + 2);
        }
      }
    }
  } else if (warp == 3) {
    // ===== m-tile loader + early residual loads =====
    if (lane == 0) {
      uint32_t n = 0;
      const uint32_t afull = lbar(A_FULL);
      for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
        const int tok0 = pr * 256 + (int)rank * 128;
        mbar_wait(bar(A_EMPTY), (n & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(bar(A_FULL), 2 * kABytes);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d_2sm(a_base + kb * kUnitBytes, &tmM, afull, kb * 64, tok0);
        // pull this tile's residual (128 rows x 1 KB) into L2 now; the staged TMA loads below then hit L2
        for (int kb = 0; kb < 8; ++kb) tma_prefetch_2d(&tmHin, kb * 32, tok0);
        // Residual boxes staged through the G buffers once they are dead (see the epilogue's staging table);
        // (once-per-tile barriers: a parity wait cannot tell completions two apart, so G_EMPTY cannot be used here)
        mbar_wait(bar(GD0), n & 1);
        for (int h = 0; h < 2; ++h) {  // s0 <- box 4, s1 <- box 6 (step 0 of parts 2,3)
          mbar_arrive_expect_tx(bar(R_FULL + h), kUnitBytes);
          tma_load_2d(g_base + h * kUnitBytes, &tmHin, bar(R_FULL + h), (2 + h) * 64, tok0);
        }
        mbar_wait(bar(T2_FULL), n & 1);
        for (int h = 0; h < 2; ++h) {  // s2 <- box 1, s3 <- box 3 (step 1 of parts 0,1)
          mbar_arrive_expect_tx(bar(R_FULL + 2 + h), kUnitBytes);
          tma_load_2d(g_base + (2 + h) * kUnitBytes, &tmHin, bar(R_FULL + 2 + h), h * 64 + 32, tok0);
        }
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
    const uint32_t t1_empty[2] = {lbar(T1_EMPTY), lbar(T1_EMPTY + 1)};
    const uint32_t g_full[2] = {lbar(G_FULL), lbar(G_FULL + 1)};
    const uint32_t t2_empty = lbar(T2_EMPTY);
    uint32_t n = 0;
    uint32_t v[32];
    const uint64_t kC0 = f2_pack(0.7978845608f, 0.7978845608f), kC1 = f2_pack(0.0356774081f, 0.0356774081f);
    const uint64_t kHalf = f2_pack(0.5f, 0.5f);
    // Staging of the final epilogue.  Part p owns the 32-column fp32 boxes 2p (step 0) and 2p+1 (step 1):
    //   step 0:  p=0 D0a   p=1 D1a   p=2 s0   p=3 s1        (D* dedicated, s0,s1 = G[0], s2,s3 = G[1])
    //   step 1:  p=0 s2    p=1 s3    p=2 D0b  p=3 D1b       (D0b/D1b: refilled by the storers of parts 0/1 after step 0)
    // R_FULL indices: s0..s3 = 0..3, D0a,D1a = 4,5, D0b,D1b = 6,7 -- every one completes exactly once per tile.
    const uint32_t slot_addr[2] = {
        part < 2 ? r_base + part * kUnitBytes : g_base + (part - 2) * kUnitBytes,
        part < 2 ? g_base + (2 + part) * kUnitBytes : r_base + (part - 2) * kUnitBytes};
    const int slot_bar[2] = {part < 2 ? 4 + part : part - 2, part < 2 ? 2 + part : 6 + (part - 2)};
    if (storer && part < 2 && pair0 < num_pairs) {  // first tile: residual box 2*part -> D[part]
      mbar_arrive_expect_tx(bar(R_FULL + 4 + part), kUnitBytes);
      tma_load_2d(r_base + part * kUnitBytes, &tmHin, bar(R_FULL + 4 + part), part * 64, pair0 * 256 + (int)rank * 128);
    }
    for (int pr = pair0; pr < num_pairs; pr += pair_step, ++n) {
      const int tok0 = pr * 256 + (int)rank * 128;
      // ---- GELU chunks: acc1[s] -> bf16 K-major tile G[s]; my 32 of the chunk's 128 columns ---------------------
      for (int j = 0; j < kChunks; ++j) {
        const int s = j & 1;
        const uint32_t use = 4 * n + (j >> 1);
        mbar_wait(bar(T1_FULL + s), use & 1);
        tc_fence_after();
        tmem_ld32(tmem_base + lane_off + 256 + 128 * s + part * 32, v);
        const float* b1 = p.b1 + j * 128 + part * 32;
        float4 bias[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) bias[q] = __ldg(reinterpret_cast<const float4*>(b1 + q * 4));
        tmem_ld_wait();
        // the accumulator stage is free as soon as it sits in registers: fc1 of chunk j+2 may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(t1_empty[s]);
        mbar_wait(bar(G_EMPTY + s), (use & 1) ^ 1);  // fc2 of the previous chunk on this buffer has retired
        const uint32_t grow = g_base + s * kGBytes + (part >> 1) * kUnitBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float bj[8] = {bias[2 * q].x, bias[2 * q].y, bias[2 * q].z, bias[2 * q].w,
                               bias[2 * q + 1].x, bias[2 * q + 1].y, bias[2 * q + 1].z, bias[2 * q + 1].w};
          uint32_t o[4];
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
            o[i] = pack_bf16(y0, y1);
          }
          sts128(grow + (((uint32_t)((part & 1) * 4 + q) ^ sw) << 4), o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(g_full[s]);
      }
      // ---- final: x = acc2 + b2 + h_in -> h_out ; LayerNorm -> u ----------------------------------------------------
      mbar_wait(bar(T2_FULL), n & 1);
      tc_fence_after();
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int st = 0; st < 2; ++st) {
        const int col0 = part * 64 + st * 32;
        const uint32_t sbase = slot_addr[st];
        mbar_wait(bar(R_FULL + slot_bar[st]), n & 1);
        tmem_ld32(tmem_base + lane_off + col0, v);
        tmem_ld_wait();
        const uint32_t rrow = sbase + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t addr = rrow + (((uint32_t)q ^ sw) << 4);
          uint32_t r[4];
          lds128(addr, r);
          const float4 bb = __ldg(reinterpret_cast<const float4*>(p.b2 + col0 + q * 4));
          const float x0 = __uint_as_float(v[4 * q]) + bb.x + __uint_as_float(r[0]);
          const float x1 = __uint_as_float(v[4 * q + 1]) + bb.y + __uint_as_float(r[1]);
          const float x2 = __uint_as_float(v[4 * q + 2]) + bb.z + __uint_as_float(r[2]);
          const float x3 = __uint_as_float(v[4 * q + 3]) + bb.w + __uint_as_float(r[3]);
          sum += (x0 + x1) + (x2 + x3);
          sq = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, sq))));
          v[4 * q] = __float_as_uint(x0);
          v[4 * q + 1] = __float_as_uint(x1);
          v[4 * q + 2] = __float_as_uint(x2);
          v[4 * q + 3] = __float_as_uint(x3);
          sts128(addr, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        tmem_st32(tmem_base + lane_off + col0, v);  // x stays in TMEM for the normalisation pass
        fence_proxy_async();
        bar_sync(2 + part, 128);
        if (storer) {
          tma_store_2d(&tmHout, sbase, col0, tok0);
          bulk_commit();
          if (st == 0 && part < 2) {  // refill D[part] with the step-1 box of part + 2
            bulk_wait_read<0>();
            mbar_arrive_expect_tx(bar(R_FULL + 6 + part), kUnitBytes);
            tma_load_2d(sbase, &tmHin, bar(R_FULL + 6 + part), (part + 2) * 64 + 32, tok0);
          }
        }
      }
      // row statistics (sum, sum of squares) over the four column parts, in a fixed order: (p0 + p1) + (p2 + p3)
      float2* st2 = stats + (part >> 1) * 128 + row;
      if (part & 1) *st2 = make_float2(sum, sq);
      tmem_st_wait();
      bar_sync(6 + quad, 128);  // the four warps that share these 32 rows
      if (!(part & 1)) {
        const float2 o = *st2;
        *st2 = make_float2(sum + o.x, sq + o.y);
      }
      if (storer) bulk_wait_read<0>();
      bar_sync(1, 512);  // pair sums visible; every h_out store has finished reading its slot
      float2 sa = stats[row];
      {
        const float2 sb = stats[128 + row];
        sa.x += sb.x;
        sa.y += sb.y;
      }
      const float mean = sa.x * (1.0f / 256.0f);
      const float rstd = rsqrtf(fmaxf(sa.y * (1.0f / 256.0f) - mean * mean, 0.f) + 1e-5f);
      {  // u box `part`: 64 bf16 columns x 128 rows, staged in G slot `part`
        const int col0 = part * 64;
        const uint32_t ubase = g_base + part * kUnitBytes;
        const uint32_t urow = ubase + row * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          tmem_ld32(tmem_base + lane_off + col0 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = col0 + c * 32 + q * 8;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.ln_g + col));
            const float4 gb = __ldg(reinterpret_cast<const float4*>(p.ln_g + col + 4));
            const float4 ba = __ldg(reinterpret_cast<const float4*>(p.ln_b + col));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln_b + col + 4));
            const float gj[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float bj[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            float y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fmaf((__uint_as_float(v[q * 8 + i]) - mean) * rstd, gj[i], bj[i]);
            sts128(urow + (((uint32_t)(c * 4 + q) ^ sw) << 4), pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]),
                   pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(t2_empty);  // acc2 fully drained: fc2(0) of the next tile may overwrite it
        fence_proxy_async();
        bar_sync(2 + part, 128);
        if (storer) {
          tma_store_2d(&tmU, ubase, col0, tok0);
          bulk_commit();
          bulk_wait_read<0>();
          if (part < 2 && pr + pair_step < num_pairs) {  // next tile: residual box 2*part -> D[part]
            mbar_arrive_expect_tx(bar(R_FULL + 4 + part), kUnitBytes);
            tma_load_2d(r_base + part * kUnitBytes, &tmHin, bar(R_FULL + 4 + part), part * 64,
                        (pr + pair_step) * 256 + (int)rank * 128);
          }
        }
      }
      bar_sync(1, 512);  // all stores have finished reading the G buffers: the next tile's GELU may overwrite them
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
  const size_t smem = kABytes + kSlots * kUnitBytes + 2 * kGBytes + 2 * kUnitBytes + 2 * 128 * 8 + 40 * 8;
  static bool configured = false;
  if (!configured) {
    DCB_CUDA(cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
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
