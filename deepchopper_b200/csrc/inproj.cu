// Hyena operator front end, fused (HF modeling_hyena.py HyenaOperator.forward up to the long conv):
//   z = in_linear(u)  (256 -> 768)          zc = short depthwise conv (3 taps, causal) of z
//   x0, x1, v = split(zc)                   vv = v * x1 ,  gate = x0          -> both bf16 channel-major [B,256,L]
// replacing gemm_kernel<INPROJ> + shortconv_gate_kernel: z ([B,768,L] bf16, 1.5 KB per token written and read back)
// never exists.  HBM traffic per token: read u (512 B), write vv + gate (1 KB).
//
// One persistent CTA per SM (a CTA-pair variant with M = 256 MMAs measured 5 % slower: the kernel is bound by waits
// around the MMA series, not by the MMA rate: with a full weight ring a 64-wide K stage issues every ~275 cycles, the
// nominal rate; the waits are the ring -- 112 KB in flight per ~3 k cycles of loaded L2 latency, 37 B/clk per SM, ~5.5
// KB/clk over the chip, i.e. the L2 output limit: 3 KB of weights per token -- and the first box of the next u tile).  A work unit is (128-token tile, 128-channel group): three tcgen05 MMA series
//   x1 = W_in[256+c..] . u^T ,  v = W_in[512+c..] . u^T ,  x0 = W_in[c..] . u^T        (M = 128 channels, K = 256)
// with N = 144 tokens: the tile plus the 16 tokens before it, so the causal 3-tap convolution (which needs z[t-1],
// z[t-2]) has its halo in the same accumulator and tiles stay independent.  The accumulator rows are channels, so an
// epilogue thread owns one channel and walks consecutive tokens: the convolution is register arithmetic.
//
//   TMEM   three 144-column regions X1 | V | X0 (stride 160), each with its own full/empty barrier: the MMA warp is
//          at most one series ahead of the epilogue (VV epilogue runs under the x0 series, G epilogue under the next x1)
//   smem   u tile 4 x 18 KB (K boxes of 144 token rows, refilled box by box as the tile's last series retires)
//          W ring 4 x 16 KB | 4 x 16 KB output staging (TMA-stored boxes 64 tokens x 128 channels) | per-channel consts
//   warp 0 producer (u boxes + W ring)   warp 1 MMA issuer   warp 2 TMEM allocator   warps 4-11 epilogue
#include "common.cuh"
#include "gemm.h"
#include "inproj.h"
#include "ptx.cuh"

#include <string.h>

namespace dcb {

using namespace ptx;

namespace {

constexpr int kThreads = 384;
constexpr int kN = 144;                       // tokens per MMA: 16 halo + 128
constexpr int kHalo = 16;
constexpr uint32_t kUBox = kN * 128;          // 18 KB: 144 token rows x 64 k
constexpr uint32_t kWBox = 128 * 128;         // 16 KB: 128 channel rows x 64 k
constexpr int kWStages = 7;   // the ring must cover ~1.5 k cycles of L2 latency at ~270 cycles per stage
constexpr int kCluster = 2;                   // CTAs that share every weight box through a multicast TMA load
constexpr uint32_t kWSlice = kWBox / kCluster;  // the rows of a box that one CTA fetches for the whole cluster
constexpr uint32_t kOutBox = 128 * 128;       // 16 KB: 128 channel rows x 64 tokens (bf16)
constexpr int kRegionStride = 160;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// two columns of 32-bit: the convolution halo
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}

// Causal 3-tap conv of 32 consecutive tokens, IN PLACE on the raw accumulator values r[0..31] (h0, h1 = raw values of
// the two tokens before).  The in_linear bias b is folded into the constant: w0 (a0+b) + w1 (a1+b) + w2 (a2+b) + cb =
// w0 a0 + w1 a1 + w2 a2 + (cb + b (w0+w1+w2)); descending order keeps the taps it still needs intact.
__device__ __forceinline__ void sconv32_inplace(uint32_t (&r)[32], uint32_t h0, uint32_t h1, float w0, float w1, float w2,
                                                float cbf) {
#pragma unroll
  for (int i = 31; i >= 2; --i)
    r[i] = __float_as_uint(fmaf(w0, __uint_as_float(r[i - 2]), fmaf(w1, __uint_as_float(r[i - 1]), fmaf(w2, __uint_as_float(r[i]), cbf))));
  r[1] = __float_as_uint(fmaf(w0, __uint_as_float(h1), fmaf(w1, __uint_as_float(r[0]), fmaf(w2, __uint_as_float(r[1]), cbf))));
  r[0] = __float_as_uint(fmaf(w0, __uint_as_float(h0), fmaf(w1, __uint_as_float(h1), fmaf(w2, __uint_as_float(r[0]), cbf))));
}

}  // namespace

template <bool kTrace>
__global__ void __launch_bounds__(kThreads, 1)
inproj_conv_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ CUtensorMap tmVV, const __grid_constant__ CUtensorMap tmGate,
                   const InprojParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t u_base = smem_u32(smem);                       // 4 x 18 KB (each 1024-aligned: 18 KB = 18 x 1024)
  const uint32_t w_base = u_base + 4 * kUBox;
  const uint32_t o_base = w_base + kWStages * kWBox;            // 2 output boxes (one per token half; VV then G)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kUBox + kWStages * kWBox + 2 * kOutBox);
  const uint32_t bar_base = smem_u32(bars);
  enum { U_FULL = 0, U_EMPTY = 4, W_FULL = 8, W_EMPTY = W_FULL + kWStages, R_FULL = W_EMPTY + kWStages, R_EMPTY = R_FULL + 3,
         N_BARS = R_EMPTY + 3 };
  auto bar = [&](int i) { return bar_base + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + N_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmU);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmVV);
    prefetch_tmap(&tmGate);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar(U_FULL + s), 1);
      mbar_init(bar(U_EMPTY + s), 1);
    }
    for (int s = 0; s < kWStages; ++s) {
      mbar_init(bar(W_FULL + s), 1);
      mbar_init(bar(W_EMPTY + s), kCluster);  // every CTA of the cluster has read the stage
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(bar(R_FULL + s), 1);
      mbar_init(bar(R_EMPTY + s), 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_ptr_smem), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peers' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int num_tiles = p.num_tiles;
  // The CTAs of a cluster walk DIFFERENT token tiles through the same (series, channel group, k box) sequence in
  // lockstep: every weight box is read from L2 once per cluster (each CTA fetches 1/kCluster of its rows and multicasts
  // them).  Weights were 2/3 of what this kernel pulls through L2 (3 KB of the 4.5 KB per token).  A cluster whose last
  // round has fewer tiles than CTAs repeats the last tile (identical bytes are stored twice).
  const uint32_t crank = cluster_ctarank();
  const int n_rounds = (num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int first = (int)(blockIdx.x / kCluster) * kCluster + (int)crank;   // == blockIdx.x
  auto tile_of = [&](int round) {
    const int o = first + round * (int)gridDim.x;
    return o < num_tiles ? o : num_tiles - 1;
  };
  constexpr uint16_t kMask = (1u << kCluster) - 1;
  const bool traced = kTrace && p.trace != nullptr && blockIdx.x == 0;

  // series order within a unit (token tile, channel group g): x1, v, x0  ->  TMEM region = series index
  // W_in row offset of series s for group g: x1 -> 256, v -> 512, x0 -> 0   (+ 128 g)
  if (warp == 0) {
    // ===== weight producer =====  (the u boxes have their own loader, warp 3: waiting here for a u box of the previous
    // tile to retire kept the weight ring from running ahead across the tile boundary, ~3 k cycles per tile)
    if (lane == 0) {
      int stage = 0;
      uint32_t wphase = 0;
      for (int rd = 0; rd < n_rounds; ++rd) {
        for (int g = 0; g < 2; ++g) {
          for (int s = 0; s < 3; ++s) {
            const int wrow = (s == 0 ? 256 : (s == 1 ? 512 : 0)) + 128 * g;
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(bar(W_EMPTY + stage), wphase ^ 1);
              mbar_arrive_expect_tx(bar(W_FULL + stage), kWBox);
              tma_load_2d_mc(w_base + stage * kWBox + crank * kWSlice, &tmW, bar(W_FULL + stage), kb * 64,
                             wrow + (int)crank * (128 / kCluster), kMask);
              if (++stage == kWStages) {
                stage = 0;
                wphase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, the elected lane issues (see ptx.cuh elect_one) =====
    {
      const uint32_t el = elect_one() ? 1u : 0u;
      TracerT<kTrace> tr{(traced && el) ? p.trace : nullptr, 0};
      constexpr uint32_t idesc = make_idesc_bf16(128, kN, false, false);
      int stage = 0;
      uint32_t wphase = 0, n = 0;
      uint32_t use = 0;  // units started (each region is used once per unit)
      for (int rd = 0; rd < n_rounds; ++rd, ++n) {
        for (int g = 0; g < 2; ++g, ++use) {
          for (int s = 0; s < 3; ++s) {
            tr(90 + s);
            mbar_wait(bar(R_EMPTY + s), (use & 1) ^ 1);
            tr(100 + s);
            tc_fence_after();
            const uint32_t d = tmem_base + s * kRegionStride;
            for (int kb = 0; kb < 4; ++kb) {
              if (g == 0 && s == 0) mbar_wait(bar(U_FULL + kb), n & 1);
              mbar_wait(bar(W_FULL + stage), wphase);
              tr(120 + s);
              tc_fence_after();
              const uint32_t a_addr = w_base + stage * kWBox;
              const uint32_t b_addr = u_base + kb * kUBox;
              umma_bf16_x4_e<1>(d, make_desc_sw128(a_addr, 16, 1024), 2, make_desc_sw128(b_addr, 16, 1024), 2, idesc, kb ? 1u : 0u, el);
              umma_commit_mc_e(bar(W_EMPTY + stage), kMask, el);
              if (g == 1 && s == 2) umma_commit_e(bar(U_EMPTY + kb), el);  // last series of the tile: u box kb may be refilled
              if (++stage == kWStages) {
                stage = 0;
                wphase ^= 1;
              }
            }
            umma_commit_e(bar(R_FULL + s), el);
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===== u-tile loader: box kb of the next tile as soon as the tile's last series has retired its k block kb =====
    if (lane == 0) {
      // u comes from HBM (the previous kernel wrote 0.5 GB of it): an L2 prefetch one tile ahead turns the 4-5 k cycle
      // wait of the tile's first series into an L2 hit
      uint32_t n = 0;
      for (int rd = 0; rd < n_rounds; ++rd, ++n) {
        const int tok0 = tile_of(rd) * 128;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(U_EMPTY + kb), (n & 1) ^ 1);
          mbar_arrive_expect_tx(bar(U_FULL + kb), kUBox);
          tma_load_2d(u_base + kb * kUBox, &tmU, bar(U_FULL + kb), kb * 64, tok0 - kHalo);
        }
        if (rd + 1 < n_rounds) {
          const int nxt = tile_of(rd + 1) * 128;
          for (int kb = 0; kb < 4; ++kb) tma_prefetch_2d(&tmU, kb * 64, nxt - kHalo);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;      // tokens [64 half, 64 half + 64) of the tile
    const int row = quad * 32 + lane;      // channel within the group
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const bool storer = (quad == 0 && lane == 0);  // one per half
    uint32_t use = 0;
    TracerT<kTrace> tr{(traced && warp == 4 && lane == 0) ? p.trace + 2 * kTraceCap : nullptr, 0};
    for (int rd = 0; rd < n_rounds; ++rd) {
      const int tok0 = tile_of(rd) * 128;
      const int b = tok0 / p.L, l0 = tok0 % p.L;
      const bool row_start = (l0 == 0);  // conv zero padding: the halo columns then hold the previous read's tail
      for (int g = 0; g < 2; ++g, ++use) {
        const int ch = g * 128 + row;  // my channel within each third
        // ---- VV: sc(x1) * sc(v) ----------------------------------------------------------------------------------
        {
          // per-channel constants straight from global memory (15 scalars per thread and unit, requested before the wait
          // for the accumulators so their latency is hidden; a shared-memory table cost a weight-ring stage)
          const int i1 = 256 + ch, iv = 512 + ch;
          const float b1 = __ldg(p.b_in + i1), w10 = __ldg(p.short_w + 3 * i1), w11 = __ldg(p.short_w + 3 * i1 + 1),
                      w12 = __ldg(p.short_w + 3 * i1 + 2), cb1 = fmaf(b1, (w10 + w11) + w12, __ldg(p.short_b + i1));
          const float bv = __ldg(p.b_in + iv), wv0 = __ldg(p.short_w + 3 * iv), wv1 = __ldg(p.short_w + 3 * iv + 1),
                      wv2 = __ldg(p.short_w + 3 * iv + 2), cbv = fmaf(bv, (wv0 + wv1) + wv2, __ldg(p.short_b + iv));
          tr(400);
          mbar_wait(bar(R_FULL + 0), use & 1);
          mbar_wait(bar(R_FULL + 1), use & 1);
          tr(410);
          tc_fence_after();
          const uint32_t t_x1 = tmem_base + lane_off + 0 * kRegionStride + kHalo + 64 * half;
          const uint32_t t_v = tmem_base + lane_off + 1 * kRegionStride + kHalo + 64 * half;
          const uint32_t obox = o_base + half * kOutBox + row * 128;
          // wait until the previous unit's store from this box has finished reading it
          if (storer) bulk_wait_read<0>();
          bar_sync(2 + half, 128);
          tr(420);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t ra[32], rb[32], ha0, ha1, hb0, hb1;
            tmem_ld32(t_x1 + 32 * sub, ra);
            tmem_ld2(t_x1 + 32 * sub - 2, ha0, ha1);
            tmem_ld32(t_v + 32 * sub, rb);
            tmem_ld2(t_v + 32 * sub - 2, hb0, hb1);
            tmem_ld_wait();
            // at the start of a read the conv input is zero-padded: make the halo's (raw + bias) vanish
            const bool zero_halo = row_start && half == 0 && sub == 0;
            if (zero_halo) {
              ha0 = ha1 = __float_as_uint(-b1);
              hb0 = hb1 = __float_as_uint(-bv);
            }
            sconv32_inplace(ra, ha0, ha1, w10, w11, w12, cb1);
            sconv32_inplace(rb, hb0, hb1, wv0, wv1, wv2, cbv);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i)
                o[i] = pack_bf16(__uint_as_float(ra[8 * q + 2 * i]) * __uint_as_float(rb[8 * q + 2 * i]),
                                 __uint_as_float(ra[8 * q + 2 * i + 1]) * __uint_as_float(rb[8 * q + 2 * i + 1]));
              sts128(obox + (((uint32_t)(sub * 4 + q) ^ sw) << 4), o[0], o[1], o[2], o[3]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar(R_EMPTY + 0));
            mbar_arrive(bar(R_EMPTY + 1));
          }
          tr(430);
          fence_proxy_async();
          bar_sync(2 + half, 128);
          if (storer) {
            tma_store_3d(&tmVV, o_base + half * kOutBox, l0 + 64 * half, g * 128, b);
            bulk_commit();
          }
        }
        // ---- G: sc(x0) ---------------------------------------------------------------------------------------------
        {
          const float b0 = __ldg(p.b_in + ch), w0 = __ldg(p.short_w + 3 * ch), w1 = __ldg(p.short_w + 3 * ch + 1),
                      w2 = __ldg(p.short_w + 3 * ch + 2), cb0 = fmaf(b0, (w0 + w1) + w2, __ldg(p.short_b + ch));
          tr(440);
          mbar_wait(bar(R_FULL + 2), use & 1);
          tr(450);
          tc_fence_after();
          const uint32_t t_x0 = tmem_base + lane_off + 2 * kRegionStride + kHalo + 64 * half;
          const uint32_t obox = o_base + half * kOutBox + row * 128;
          // The accumulator goes to registers and the region back to the MMA warp BEFORE the wait for the staging box
          // (the previous unit's TMA store takes ~1.7 k cycles to read it): x0 of the next unit was stalling on that.
          uint32_t ra[32], rb[32], ha0, ha1;
          tmem_ld32(t_x0, ra);
          tmem_ld2(t_x0 - 2, ha0, ha1);
          tmem_ld32(t_x0 + 32, rb);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(R_EMPTY + 2));
          tr(460);
          if (storer) bulk_wait_read<0>();
          bar_sync(2 + half, 128);
          tr(470);
          if (row_start && half == 0) ha0 = ha1 = __float_as_uint(-b0);
          const uint32_t hb0 = ra[30], hb1 = ra[31];  // raw values: the halo of the second 32 tokens
          sconv32_inplace(ra, ha0, ha1, w0, w1, w2, cb0);
          sconv32_inplace(rb, hb0, hb1, w0, w1, w2, cb0);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t (&r)[32] = sub ? rb : ra;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              sts128(obox + (((uint32_t)(sub * 4 + q) ^ sw) << 4),
                     pack_bf16(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1])),
                     pack_bf16(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3])),
                     pack_bf16(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5])),
                     pack_bf16(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7])));
          }
          fence_proxy_async();
          bar_sync(2 + half, 128);
          if (storer) {
            tma_store_3d(&tmGate, o_base + half * kOutBox, l0 + 64 * half, g * 128, b);
            bulk_commit();
          }
          tr(480);
        }
      }
    }
    if (storer) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // a peer may still be multicasting into my ring / signalling my barriers until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_inproj_conv(dcb200_ctx* ctx, const CUtensorMap& tm_u, const CUtensorMap& tm_w, const CUtensorMap& tm_vv,
                       const CUtensorMap& tm_gate, const InprojParams& p) {
  const size_t smem = 4 * kUBox + kWStages * kWBox + 2 * kOutBox + 40 * 8 + 1024;
  auto kern = p.trace ? &inproj_conv_kernel<true> : &inproj_conv_kernel<false>;
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(kern), smem));
  int clusters = ctx->sm_count / kCluster;
  const int want = (p.num_tiles + kCluster - 1) / kCluster;
  if (want < clusters) clusters = want;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(kCluster * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof(ctx, K_INPROJ);
  DCB_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_u, tm_w, tm_vv, tm_gate, p));
  ctx->launches++;
  return DCB200_OK;
}

}  // namespace dcb
