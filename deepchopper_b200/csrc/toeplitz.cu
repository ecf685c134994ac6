// Tensor-core form of the Hyena long convolution, the product path below the crossover with the blocked FFT kernel
// (lconv.cu; any L <= 32768 works):
//   y[b,c,t] = gate[b,c,t] * sum_{s<=t} vv[b,c,s] * k'_c[t-s],   vv = sc(v) * sc(x1),  gate = sc(x0),  k'[0] = k[0] + D
//
// Per channel the causal convolution is a GEMM  Y[b, t] = sum_s V[b, s] K[s, t]  with a Toeplitz K.  One CTA owns a
// work item (channel c, 128 batch rows): A = V tiles streamed by TMA (K-major, 128B swizzle), D = 128 x 256 fp32
// accumulators in TMEM (two of them, so the epilogue of one output tile overlaps the MMAs of the next).
//
// The Toeplitz operand is never materialised as tiles.  With the accumulator columns in REVERSED token order
// (column n <-> token t_hi - n) the B operand entry for (column n, k-index s') is k'[Q - n - s'] - it depends on
// n + s' only.  In the no-swizzle K-major UMMA layout a core matrix is 8 rows x 16 bytes, and the core matrix at
// (n/8 = a, s'/8 = b) then depends on a + b only, so ONE array of (L/8 + 32) core matrices
//     E[i][r][c] = k'[8 i + 7 - r - c]           (128 bytes per i, stored by descending i, zero for negative taps)
// with LBO = SBO = 128 bytes serves every (output tile, input block) pair of the channel: the descriptor start
// address selects Q.  A 64-token K chunk touches a window of 40 consecutive core matrices (5 KB), which the producer
// bulk-copies from the (L2-resident, 16 L bytes per channel) table next to the 16 KB A tile of the same pipeline
// stage, so HBM traffic stays at the algorithmic minimum: read vv and gate, write y.  Causality (s > t) falls out of
// the zero taps.  Cost grows with L (64 KFLOP x (L/128 + 1)/2 per token).
//
//   warp 0   producer: per stage one A tile (3-D TMA box 64 tokens x 1 channel x 128 rows) + its E window
//   warp 1   MMA issuer (tcgen05.mma M=128, N=256 / 128 on the diagonal block, K=16)
//   warp 2   TMEM allocator
//   warp 3   gate-tile loader (TMA, same box shape)
//   warps 4-11 epilogue: tcgen05.ld, multiply by the gate tile IN PLACE in its swizzled smem slot, TMA store
#include "common.cuh"
#include "gemm.h"
#include "ptx.cuh"
#include "toeplitz.h"

namespace dcb {

using namespace ptx;

constexpr int kTzThreads = 384;
constexpr int kTzAStages = 7;   // 16 KB A tile (128 rows x 64 tokens) + 5 KB window of Toeplitz core matrices
constexpr uint32_t kTzWin = 40 * 128;           // bytes of one E window
constexpr uint32_t kTzStage = 128 * 128 + 6144; // stage pitch (keeps the A tiles 1024-byte aligned)
// 16 KB each.  A slot is released one box late (when the NEXT store has been issued and this one's read is done), so
// with 3 slots only one gate box could be in flight: 4 slots are worth 6-7 % at L = 1280 (HBM-bound regime); a fifth
// slot, or more A stages, change nothing.  Tried and dropped: L2 prefetch of the next item's first 512 tokens (+11 %);
// one store per TMEM lane quadrant behind 64-thread barriers instead of one per box behind a 256-thread one (no change).
constexpr int kTzGSlots = 4;
constexpr int kTzZeroChunks = 32;
constexpr uint32_t kTzBox = 128 * 128;  // bytes of one TMA box

// E table: [256 channels][cap/8 + 32 chunks][8 rows][8 cols] bf16, chunk position p <-> i = cap/8 - 1 - p
__global__ void __launch_bounds__(256) toeplitz_table_kernel(const float* __restrict__ k, int k_stride, int k_len,
                                                             const float* __restrict__ D, int cap,
                                                             __nv_bfloat16* __restrict__ E) {
  const int P = cap / 8 + kTzZeroChunks;
  const int c = blockIdx.y;
  const float* kc = k + (size_t)c * k_stride;
  __nv_bfloat16* dst = E + (size_t)c * P * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P * 64; idx += gridDim.x * blockDim.x) {
    const int p = idx >> 6, r = (idx >> 3) & 7, cc = idx & 7;
    const int i = cap / 8 - 1 - p;
    const int u = 8 * i + 7 - r - cc;
    float v = 0.f;
    if (u >= 0 && u < k_len) v = kc[u];
    if (u == 0) v += D[c];
    dst[idx] = __float2bfloat16_rn(v);
  }
}

struct ToepParams {
  const __nv_bfloat16* E;  // table base
  int L;                   // padded length (multiple of 128, <= cap)
  int cap;                 // table built for reads up to cap tokens
  int n_rt;                // ceil(B / 128)
  int n_parts;             // output-tile ranges per (channel, row tile): equal MMA work each (see launch_toeplitz_conv)
  int jb[9];               // part k covers the 256-token output tiles [jb[k], jb[k+1])
  int n_items;             // 256 * n_rt * n_parts
  long long* trace;        // optional timeline trace (DCB200_TRACE=toeplitz), or null
};

__device__ __forceinline__ uint32_t tz_pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <bool kTrace>
__global__ void __launch_bounds__(kTzThreads, 1)
toeplitz_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmG,
                const __grid_constant__ CUtensorMap tmY, const ToepParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_base = smem_u32(smem);
  const uint32_t g_base = a_base + kTzAStages * kTzStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTzAStages * kTzStage + kTzGSlots * kTzBox);
  // bars: afull[S] aempty[S] gfull[G] gempty[G] (4 unused) tfull[2] tempty[2]
  const uint32_t bar_base = smem_u32(bars);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (kTzAStages + s); };
  auto gfull = [&](int s) { return bar_base + 8u * (2 * kTzAStages + s); };
  auto gempty = [&](int s) { return bar_base + 8u * (2 * kTzAStages + kTzGSlots + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (2 * kTzAStages + 2 * kTzGSlots + 4 + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (2 * kTzAStages + 2 * kTzGSlots + 6 + s); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kTzAStages + 2 * kTzGSlots + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmV);
    prefetch_tmap(&tmG);
    prefetch_tmap(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kTzAStages; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < kTzGSlots; ++s) {
      mbar_init(gfull(s), 1);
      mbar_init(gempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int L = p.L;
  const int n_tiles = (L + 255) / 256;

  if (warp == 0) {
    // ===== producer: A tile + E window per stage (whole warp runs the loop, the elected lane issues) =====
    {
      const uint32_t el = elect_one() ? 1u : 0u;
      int stage = 0;
      uint32_t phase = 0;
      TracerT<kTrace> tr{(kTrace && p.trace && blockIdx.x == 0 && el) ? p.trace + 2 * 2 * kTraceCap : nullptr, 0};
      const int kP = p.cap / 8 + kTzZeroChunks;  // chunks per channel in the table
      for (int o = blockIdx.x; o < p.n_items; o += gridDim.x) {
        const int part = o % p.n_parts, cr = o / p.n_parts;
        const int c = cr / p.n_rt, rt = cr % p.n_rt;
        const __nv_bfloat16* e_ch = p.E + (size_t)c * kP * 64;
        for (int J = p.jb[part]; J < p.jb[part + 1]; ++J) {
          const int t_hi = min(256 * J + 255, L - 1);
          const int nkc = 2 * (t_hi / 128 + 1);  // 64-token chunks of the input blocks 0..t_hi/128
          const int q0 = (t_hi - 7) / 8;
          for (int kc = 0; kc < nkc; ++kc) {
            tr(300);
            mbar_wait(aempty(stage), phase ^ 1);
            tr(310);
            mbar_arrive_expect_tx_e(afull(stage), kTzBox + kTzWin, el);
            const uint32_t dst = a_base + stage * kTzStage;
            tma_load_3d_e(dst, &tmV, afull(stage), kc * 64, c, rt * 128, el);
            // window: core matrices i = q, q-1, ..., q-39 (table position p = cap/8 - 1 - i)
            const int q = q0 - 8 * kc;
            bulk_load_1d_e(dst + kTzBox, e_ch + (size_t)(p.cap / 8 - 1 - q) * 64, kTzWin, afull(stage), el);
            if (++stage == kTzAStages) {
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
      TracerT<kTrace> tr{(kTrace && p.trace && blockIdx.x == 0 && el) ? p.trace : nullptr, 0};
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int o = blockIdx.x; o < p.n_items; o += gridDim.x) {
        const int part = o % p.n_parts;
        for (int J = p.jb[part]; J < p.jb[part + 1]; ++J) {
          const int t_hi = min(256 * J + 255, L - 1);
          const int ntile = t_hi - 256 * J + 1;  // 256 or 128
          const int ilast = t_hi / 128;
          tr(90);
          mbar_wait(tempty(acc), acc_phase ^ 1);
          tr(100);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          const uint32_t idesc_full = make_idesc_bf16(128, ntile, false, false);
          const uint32_t idesc_diag = make_idesc_bf16(128, 128, false, false);
          for (int i = 0; i <= ilast; ++i) {
            const uint32_t idesc = (i == ilast) ? idesc_diag : idesc_full;
            for (int hh = 0; hh < 2; ++hh) {
              tr(110);
              mbar_wait(afull(stage), phase);
              tr(120);
              tc_fence_after();
              const uint32_t a_addr = a_base + stage * kTzStage;
              // A: +32 B per K = 16 slice inside the swizzled tile; B: the window slides by two core matrices (256 B)
              umma_bf16_x4_e<1>(d_tmem, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_nosw(a_addr + kTzBox, 128, 128), 16, idesc,
                              (i | hh) ? 1u : 0u, el);
              umma_commit_e(aempty(stage), el);
              if (++stage == kTzAStages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          umma_commit_e(tfull(acc), el);
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== gate tile loader =====
    if (lane == 0) {
      int slot = 0;
      uint32_t gphase = 0;
      for (int o = blockIdx.x; o < p.n_items; o += gridDim.x) {
        const int part = o % p.n_parts, cr = o / p.n_parts;
        const int c = cr / p.n_rt, rt = cr % p.n_rt;
        for (int t0 = 256 * p.jb[part]; t0 < min(L, 256 * p.jb[part + 1]); t0 += 64) {
          mbar_wait(gempty(slot), gphase ^ 1);
          mbar_arrive_expect_tx(gfull(slot), kTzBox);
          tma_load_3d(g_base + slot * kTzBox, &tmG, gfull(slot), t0, c, rt * 128);
          if (++slot == kTzGSlots) {
            slot = 0;
            gphase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;  // which 32 tokens of a 64-token box
    const int row = quad * 32 + lane;
    const bool storer = (threadIdx.x == 128);
    int slot = 0, acc = 0, prev_slot = -1;
    uint32_t gphase = 0, acc_phase = 0;
    uint32_t v[32];
    TracerT<kTrace> tr{(kTrace && p.trace && blockIdx.x == 0 && warp == 4 && lane == 0) ? p.trace + 2 * kTraceCap : nullptr, 0};
    for (int o = blockIdx.x; o < p.n_items; o += gridDim.x) {
      const int part = o % p.n_parts, cr = o / p.n_parts;
      const int c = cr / p.n_rt, rt = cr % p.n_rt;
      for (int J = p.jb[part]; J < p.jb[part + 1]; ++J) {
        const int t_hi = min(256 * J + 255, L - 1);
        const int ntile = t_hi - 256 * J + 1;
        tr(400);
        mbar_wait(tfull(acc), acc_phase);
        tr(410);
        tc_fence_after();
        const uint32_t t_row = tmem_base + acc * 256 + ((uint32_t)(quad * 32) << 16);
        for (int x = 0; x < ntile / 64; ++x) {
          mbar_wait(gfull(slot), gphase);
          const int c0 = ntile - 64 * x - 32 * half - 32;  // accumulator columns of my 32 tokens (reversed order)
          tmem_ld32(t_row + c0, v);
          tmem_ld_wait();
          const uint32_t rowaddr = g_base + slot * kTzBox + row * 128;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t addr = rowaddr + (uint32_t)(((4 * half + jj) ^ (row & 7)) * 16);
            uint32_t g0, g1, g2, g3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(g0), "=r"(g1), "=r"(g2), "=r"(g3) : "r"(addr) : "memory");
            const int m = 31 - 8 * jj;  // token 8 jj + i of my span <-> v[m - i]
            const uint32_t w0 = tz_pack(__uint_as_float(v[m]) * __uint_as_float(g0 << 16),
                                        __uint_as_float(v[m - 1]) * __uint_as_float(g0 & 0xffff0000u));
            const uint32_t w1 = tz_pack(__uint_as_float(v[m - 2]) * __uint_as_float(g1 << 16),
                                        __uint_as_float(v[m - 3]) * __uint_as_float(g1 & 0xffff0000u));
            const uint32_t w2 = tz_pack(__uint_as_float(v[m - 4]) * __uint_as_float(g2 << 16),
                                        __uint_as_float(v[m - 5]) * __uint_as_float(g2 & 0xffff0000u));
            const uint32_t w3 = tz_pack(__uint_as_float(v[m - 6]) * __uint_as_float(g3 << 16),
                                        __uint_as_float(v[m - 7]) * __uint_as_float(g3 & 0xffff0000u));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
          }
          fence_proxy_async();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (storer) {
            tma_store_3d(&tmY, g_base + slot * kTzBox, 256 * J + 64 * x, c, rt * 128);
            bulk_commit();
            if (prev_slot >= 0) {
              bulk_wait_read<1>();  // the previous store has finished reading its slot
              mbar_arrive(gempty(prev_slot));
            }
            prev_slot = slot;
          }
          if (++slot == kTzGSlots) {
            slot = 0;
            gphase ^= 1;
          }
        }
        tc_fence_before();
        mbar_arrive(tempty(acc));
        tr(480);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    if (storer) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_toeplitz_table(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, int cap,
                          __nv_bfloat16* E) {
  dim3 grid(8, 256);
  toeplitz_table_kernel<<<grid, 256, 0, ctx->stream>>>(k, k_stride, k_len, D, cap, E);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

size_t toeplitz_table_bytes(int cap) { return (size_t)256 * (cap / 8 + kTzZeroChunks) * 128; }

int launch_toeplitz_conv(dcb200_ctx* ctx, const __nv_bfloat16* E, int cap, const CUtensorMap& tm_vv,
                         const CUtensorMap& tm_gate, const CUtensorMap& tm_y, int B, int L) {
  if (L % 128 != 0 || L > cap || L <= 0) {
    set_error("toeplitz conv: L=%d must be a multiple of 128 and <= %d", L, cap);
    return DCB200_EINVAL;
  }
  ToepParams p;
  p.E = E;
  p.L = L;
  p.cap = cap;
  p.n_rt = (B + 127) / 128;
  // Work items are (channel, 128-row tile, range of output tiles).  A long read's batch has few row tiles (one for
  // 16-32 kb reads): 256 items over 148 SMs are 1.73 waves, i.e. 14 % of the machine idle in the second one.  The output
  // tiles of an item are therefore dealt to up to 8 parts of equal MMA work (tile J costs ~4 J + 4 chunks), enough
  // parts for >= 8 waves.
  const int n_tiles = (L + 255) / 256;
  int parts = 1;
  const int base_items = 256 * p.n_rt;
  // (only where the kernel is MMA-bound, L >= 2048: a part re-reads the input blocks of the parts before it, which
  // costs the HBM-bound short reads 30 %)
  if (L >= 2048 && base_items < 8 * ctx->sm_count) parts = (8 * ctx->sm_count + base_items - 1) / base_items;
  if (parts > 8) parts = 8;
  if (parts > n_tiles) parts = n_tiles;
  p.n_parts = parts;
  {
    const double total = 2.0 * n_tiles * n_tiles + 2.0 * n_tiles;
    int J = 0;
    p.jb[0] = 0;
    for (int k = 1; k < parts; ++k) {
      while (J < n_tiles && 2.0 * J * J + 2.0 * J < total * k / parts) ++J;
      if (J <= p.jb[k - 1]) J = p.jb[k - 1] + 1;  // every part gets at least one tile
      if (J > n_tiles - (parts - k)) J = n_tiles - (parts - k);
      p.jb[k] = J;
    }
    p.jb[parts] = n_tiles;
    for (int k = parts + 1; k < 9; ++k) p.jb[k] = n_tiles;
  }
  p.n_items = base_items * parts;
  const size_t want = (size_t)kTzAStages * kTzStage + (size_t)kTzGSlots * kTzBox + 1024 + 512;
  p.trace = nullptr;
  {
    if (ctx->trace_kind == TRACE_TOEPLITZ && !ctx->traced_once) {
      ctx->traced_once = true;
      DevBuf& bt = ctx->buf("trace");
      DCB_CHECK(bt.reserve(3 * 4096 * 2 * 8));
      DCB_CUDA(cudaMemsetAsync(bt.p, 0, 3 * 4096 * 2 * 8, ctx->stream));
      p.trace = bt.as<long long>();
    }
  }
  auto kern = p.trace ? &toeplitz_kernel<true> : &toeplitz_kernel<false>;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(kern), want));
  const int grid = p.n_items < ctx->sm_count ? p.n_items : ctx->sm_count;
  ProfScope prof(ctx, K_TOEP);
  kern<<<grid, kTzThreads, want, ctx->stream>>>(tm_vv, tm_gate, tm_y, p);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
