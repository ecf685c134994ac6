// O(L log L) Hyena long convolution for long reads (BASELINE configs[3]): thread-level building blocks.
//
//   y[b,c,t] = gate[b,c,t] * sum_{s<=t} vv[b,c,s] * k'_c[t-s],      k'[0] = k[0] + D
//
// The reference does rfft / irfft of size 2L through cuFFT on fp32 tensors (fftconv in the HF modeling_hyena.py,
// restated in SURVEY.md Appendix A; call site deepchopper/models/llm/hyena.py:34-41).  Here one CTA owns one
// (batch row, channel) sequence and walks it in blocks of P = 8192 tokens (overlap-save block convolution):
//
//   y_i = first P outputs of  sum_{j<=i} circ_{2P}(c_{i-j}) [v_j ; 0],      c_d = k'[dP .. dP+P) | 0 | k'[dP-P+1 .. dP)
//
// Every block is ONE real FFT of 2P points done as a complex FFT of N = P points in shared memory (fp32, 64 KB):
//   * z[n] = v[2n] + i v[2n+1]; the upper half of z is zero padding, so the first radix-2 DIF step is free and is
//     applied while loading (half A = z, half B = z W_N^n); each half then takes three radix-16 passes (4096 = 16^3)
//     and stays in digit-reversed order;
//   * one pointwise pass untangles the spectrum of the real sequence from Z[k], Z[N-k], multiplies it by the cached
//     filter spectrum of block distance 0, adds the products of the EARLIER blocks' spectra (kept in an L2-resident
//     scratch of the CTA) with the filter spectra of distance i-j, and re-tangles the result;
//   * three mirrored radix-16 DIT passes, and the last radix-2 step is applied while storing (times the gate, bf16).
// HBM sees vv, gate and y once; the filter spectra (64 KB per channel and distance, shared by all rows) and the scratch
// live in L2.  The same functions compile for the host (tests/native/lconv_check.cu emulates a CTA thread by thread).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define LC_HD __host__ __device__ __forceinline__
#else
#error "lconv_core.cuh needs nvcc (host emulation is compiled with nvcc too)"
#endif

namespace dcb {
namespace lc {

constexpr int kLogP = 13;
constexpr int kP = 1 << kLogP;    // tokens per block = complex FFT points N
constexpr int kH = kP / 2;        // 4096: one radix-2 half = 16^3
constexpr int kSlots = kH;        // float4 slots of one block spectrum (pairs k, N-k)
constexpr int kThreads = 256;     // = kH / 16 butterflies per half and pass
constexpr int kMaxBlocks = 4;     // 4 x 8192 = the model's 32768 tokens
constexpr int kXFloat2 = kP + kP / 16;  // shared-memory array with one pad slot per 16 elements

// twiddle table (float2): T1[n] = W_8192^n, T2[k] = W_4096^k, T3[slot] = -i W_16384^{k(slot)}, 4096 entries each
constexpr int kTwT1 = 0, kTwT2 = kH, kTwT3 = 2 * kH, kTwTotal = 3 * kH;

LC_HD int padi(int i) { return i + (i >> 4); }
LC_HD int rev3(int p) { return ((p & 0xF) << 8) | (p & 0xF0) | ((p >> 8) & 0xF); }

// Complex arithmetic (scalar fp32; measured: Blackwell's packed FADD2 / FMUL2 / FFMA2 forms shorten the instruction
// stream by a quarter but not the run time -- the fp32 pipes, not the issue slots, are what the passes fill).
LC_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
LC_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
LC_HD float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
LC_HD float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
LC_HD float2 cfma(float2 a, float2 b, float2 c) {  // a * b + c
  return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
LC_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// Global accesses with an L2 eviction policy (createpolicy, lconv.cu): the per-CTA scratch of block spectra is re-read
// by the later blocks of the same sequence and should stay in L2 (evict_last), the activations stream through once
// (evict_first); the filter spectra take the default policy (shared by the rows in flight, then dead).
// `pol` is ignored by the host emulation.
LC_HD float4 ld_f4_hint(const float4* p, uint64_t pol) {
#if defined(__CUDA_ARCH__)
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
#else
  (void)pol;
  return *p;
#endif
}
LC_HD void st_f4_hint(float4* p, float4 v, uint64_t pol) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(pol)
               : "memory");
#else
  (void)pol;
  *p = v;
#endif
}
LC_HD uint32_t ld_u32_hint(const uint32_t* p, uint64_t pol) {
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
#else
  (void)pol;
  return *p;
#endif
}
LC_HD void st_u32_hint(uint32_t* p, uint32_t v, uint64_t pol) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
#else
  (void)pol;
  *p = v;
#endif
}

LC_HD float bf16_bits_to_f32(uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b << 16);
#else
  union { uint32_t u; float f; } x;
  x.u = b << 16;
  return x.f;
#endif
}
LC_HD uint32_t f32_to_bf16_bits(float f) {  // round to nearest even (finite inputs)
#if defined(__CUDA_ARCH__)
  uint32_t u = __float_as_uint(f);
#else
  union { uint32_t u; float f; } x;
  x.f = f;
  uint32_t u = x.u;
#endif
  u += 0x7FFFu + ((u >> 16) & 1u);
  return u >> 16;
}

// slot -> (position of Z[k], position of Z[N-k]) in the shared array, and whether it is the self-paired special slot
LC_HD void slot_positions(int slot, int& pa, int& pb, bool& special) {
  special = false;
  if (slot < kH / 2) {  // odd frequencies k = 2 rev3(p) + 1 live in half B; N - k sits at the mirrored position
    pa = kH + slot;
    pb = kH + (kH - 1 - slot);
  } else {              // even frequencies k = 2 k', k' = rev3(p) in [0, 2048): positions with bit 3 clear
    const int i = slot - kH / 2;
    const int p = ((i >> 3) << 4) | (i & 7);
    pa = p;
    if (i == 0) {
      special = true;   // k = 0 (with k = N folded in) and k = N/2
      pb = 8;           // rev3(2048)
    } else {
      pb = rev3(kH - rev3(p));
    }
  }
}
// frequency index k (of the 2P-point real transform) of a slot
LC_HD int slot_freq(int slot) {
  if (slot < kH / 2) return 2 * rev3(slot) + 1;
  const int i = slot - kH / 2;
  return 2 * rev3(((i >> 3) << 4) | (i & 7));
}

// multiply by exp(-+ 2 pi i k16 / 16), k16 a compile-time constant after unrolling
template <bool INV> LC_HD float2 mul_w16(float2 d, int k16) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  if (k16 == 0) return d;
  if (k16 == 4) return INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
  float c, s;
  switch (k16) {
    case 1: c = C1; s = S1; break;
    case 2: c = R2; s = R2; break;
    case 3: c = S1; s = C1; break;
    case 5: c = -S1; s = C1; break;
    case 6: c = -R2; s = R2; break;
    default: c = -C1; s = S1; break;
  }
  if (!INV) s = -s;
  return make_float2(fmaf(d.x, c, -d.y * s), fmaf(d.x, s, d.y * c));
}

LC_HD constexpr int brev4(int q) { return ((q & 1) << 3) | ((q & 2) << 1) | ((q & 4) >> 1) | ((q & 8) >> 3); }

// 16-point DFT in registers (radix-2 decimation in frequency); output q lands in a[brev4(q)]
template <bool INV> LC_HD void dft16(float2 (&a)[16]) {
#pragma unroll
  for (int len = 16; len >= 2; len >>= 1) {
    const int half = len >> 1;
#pragma unroll
    for (int blk = 0; blk < 16; blk += len) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float2 u = a[blk + j], v = a[blk + j + half];
        a[blk + j] = cadd(u, v);
        a[blk + j + half] = mul_w16<INV>(csub(u, v), j * (16 / len));
      }
    }
  }
}

// w[q] = w1^q, q = 1..15, from w1 and w4 = w1^4 (both straight from the table: every power is at most 3 complex
// multiplications away from a correctly rounded value)
LC_HD void twiddle_powers(float2 (&w)[16], float2 w1, float2 w4) {
  w[0] = make_float2(1.f, 0.f);
  w[1] = w1;
  w[2] = cmul(w1, w1);
  w[3] = cmul(w[1], w[2]);
  w[4] = w4;
#pragma unroll
  for (int q = 5; q < 8; ++q) w[q] = cmul(w[4], w[q - 4]);
  w[8] = cmul(w4, w4);
#pragma unroll
  for (int q = 9; q < 16; ++q) w[q] = cmul(w[8], w[q - 8]);
}

// The per-thread twiddle bases of the passes: they depend on the thread index only, so the convolution kernel loads
// them once (10 registers) instead of fetching table entries at the head of every pass of every block.
struct ThreadTw {
  float2 wB;          // W_8192^tid: the radix-2 step folded into pass 0 of half B
  float2 p0a, p0b;    // pass 0: W_4096^tid and its 4th power
  float2 p1a, p1b;    // pass 1: W_256^j and its 4th power, j = tid % 16
};
LC_HD ThreadTw load_thread_tw(const float2* __restrict__ tw, int tid) {
  ThreadTw t;
  const int j = tid & 15;
  t.wB = tw[kTwT1 + tid];
  t.p0a = tw[kTwT2 + tid];
  t.p0b = tw[kTwT2 + 4 * tid];
  t.p1a = tw[kTwT2 + 16 * j];
  t.p1b = tw[kTwT2 + 64 * j];
  return t;
}

// One radix-16 pass of thread `tid` over both halves.  PASS 0: n = 4096 (s = 256), 1: n = 256 (s = 16), 2: n = 16 (s = 1);
// (w1, w4) = (W_n^j, W_n^4j) with j = tid % s (unused for PASS 2).
// The 16 elements of a butterfly sit at padi(base + m s) = padi(base) + m ps with ps = s + s / 16 (s a multiple of 16)
// or ps = 1 (s = 1: base is a multiple of 16), and half B starts at padi(kH) = kH + kH / 16: one base register and
// immediate offsets per access.
template <int PASS, bool INV> LC_HD void radix16_pass(float2* X, float2 w1, float2 w4, int tid) {
  constexpr int s = PASS == 0 ? 256 : (PASS == 1 ? 16 : 1);
  constexpr int ps = s >= 16 ? s + s / 16 : s;
  constexpr int n = 16 * s;
  const int blk = tid / s, j = tid % s;
  float2* Xb = X + padi(blk * n + j);
  float2 w[16];
  if (s > 1) twiddle_powers(w, w1, w4);
  // (not unrolled: the two halves run the same code, and the kernel's instruction footprint matters -- the CTAs of an SM
  // are in different phases and share its instruction cache)
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float2* Xh = Xb + half * (kH + kH / 16);
    float2 a[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m] = Xh[m * ps];
    if (!INV) {
      dft16<false>(a);
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        float2 v = a[brev4(q)];
        if (s > 1 && q > 0) v = cmul(v, w[q]);
        Xh[q * ps] = v;
      }
    } else {
      if (s > 1) {
#pragma unroll
        for (int q = 1; q < 16; ++q) a[q] = cmulc(a[q], w[q]);
      }
      dft16<true>(a);
#pragma unroll
      for (int m = 0; m < 16; ++m) Xh[brev4(m) * ps] = a[m];
    }
  }
}

// multiply by exp(-+ 2 pi i m / 32), m in [0, 16) a compile-time constant after unrolling
template <bool INV> LC_HD float2 mul_w32(float2 d, int m) {
  if (m & 1) {
    const float C[8] = {0.98078528040323043f, 0.83146961230254524f, 0.55557023301960218f, 0.19509032201612825f,
                        -0.19509032201612825f, -0.55557023301960218f, -0.83146961230254524f, -0.98078528040323043f};
    const float S[8] = {0.19509032201612825f, 0.55557023301960218f, 0.83146961230254524f, 0.98078528040323043f,
                        0.98078528040323043f, 0.83146961230254524f, 0.55557023301960218f, 0.19509032201612825f};
    const float c = C[m >> 1], sn = INV ? S[m >> 1] : -S[m >> 1];
    return make_float2(fmaf(d.x, c, -d.y * sn), fmaf(d.x, sn, d.y * c));
  }
  return mul_w16<INV>(d, m >> 1);
}

LC_HD float2 unpack_bf16x2(uint32_t w) { return make_float2(bf16_bits_to_f32(w & 0xFFFFu), bf16_bits_to_f32(w >> 16)); }

// Forward pass 0 fused with the load of a block: zraw[m] = the bf16 pair (tokens 2n, 2n+1) of n = tid + 256 m, i.e.
// z[n] = zraw.lo + i zraw.hi (zero beyond the read).  The upper half of z is zero padding, so the radix-2 DIF step is
//   half A input = z[n],   half B input = z[n] W_8192^n = z[n] W_32^m W_8192^tid
// and W_8192^tid, constant over the butterfly, moves behind the 16-point DFT (into the output twiddle).
LC_HD void fwd_pass0_fused(float2* X, const uint32_t (&zraw)[16], const ThreadTw& t, int tid) {
  float2* XA = X + padi(tid);
  float2* XB = XA + (kH + kH / 16);
  float2 w[16];
  twiddle_powers(w, t.p0a, t.p0b);
  {
    float2 a[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m] = unpack_bf16x2(zraw[m]);
    dft16<false>(a);
#pragma unroll
    for (int q = 0; q < 16; ++q) XA[q * 272] = q ? cmul(a[brev4(q)], w[q]) : a[0];
  }
  {
    float2 a[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m] = mul_w32<false>(unpack_bf16x2(zraw[m]), m);
    dft16<false>(a);
#pragma unroll
    for (int q = 0; q < 16; ++q) XB[q * 272] = cmul(a[brev4(q)], q ? cmul(w[q], t.wB) : t.wB);
  }
}

// Inverse pass 0 fused with the store of a block: the mirror of fwd_pass0_fused.  Half A's result is parked in its own
// shared-memory positions (same thread, no barrier) while half B is transformed; then
//   z'[n] = A'[n] + conj(W_8192^n) B'[n],   y[2n] = Re z' * gate[2n],  y[2n+1] = Im z' * gate[2n+1]   (bf16 pair out[m])
// `load_gate(m)` fetches the gate pair of n = tid + 256 m; it is called after the twiddles of half B have been applied so
// that the loads are in flight under the 16-point DFT without holding registers earlier.
template <class GateFn> LC_HD void inv_pass0_fused(float2* X, const ThreadTw& t, int tid, GateFn load_gate, uint32_t (&out)[16]) {
  float2* XA = X + padi(tid);
  float2* XB = XA + (kH + kH / 16);
  float2 w[16];
  twiddle_powers(w, t.p0a, t.p0b);
  {
    float2 a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = q ? cmulc(XA[q * 272], w[q]) : XA[0];
    dft16<true>(a);
#pragma unroll
    for (int m = 0; m < 16; ++m) XA[brev4(m) * 272] = a[m];
  }
  float2 b[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) b[q] = cmulc(XB[q * 272], q ? cmul(w[q], t.wB) : t.wB);
  uint32_t g[16];
#pragma unroll
  for (int m = 0; m < 16; ++m) g[m] = load_gate(m);
  dft16<true>(b);
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 zz = cadd(XA[m * 272], mul_w32<true>(b[brev4(m)], m));
    const float2 gg = unpack_bf16x2(g[m]);
    out[m] = f32_to_bf16_bits(zz.x * gg.x) | (f32_to_bf16_bits(zz.y * gg.y) << 16);
  }
}

// (A, B) = (Z[k], Z[N-k]) -> twice the spectrum of the real sequence at k and N-k (w = -i W_2N^k); with (A, B) =
// (Y[k], Y[N-k]) and conj(w) it is the inverse step (twice Z'[k], Z'[N-k])
LC_HD void tangle(float2 A, float2 B, float2 w, float2& outk, float2& outm) {
  const float2 Bc = cconj(B);
  const float2 E = cadd(A, Bc), D = csub(A, Bc);
  const float2 t = cmul(w, D);
  outk = cadd(E, t);
  outm = cconj(csub(E, t));
}

// Filter-spectrum prologue: the full 2P-point real sequence c (fp32) -> halves A and B (general radix-2 step)
LC_HD void prologue_store_full(float2* X, const float2* __restrict__ T1, int n, float2 zlo, float2 zhi) {
  X[padi(n)] = cadd(zlo, zhi);
  X[padi(kH + n)] = cmul(csub(zlo, zhi), T1[n]);
}

// Pointwise pass of block i for G slots of one thread (slot = (it0 + g) * kThreads + tid): all global loads of the group
// (twiddle, filter spectrum, then per earlier block its spectrum and the filter spectrum of that distance) are issued
// before the arithmetic that needs them.  K: filter spectra of this channel, [nbK][kSlots] float4 (distance-major);
// S: this CTA's scratch, [kMaxBlocks-1][kSlots] float4 (spectra of the earlier blocks of the current sequence).
template <int G>
LC_HD void pointwise_group(float2* X, const float2* __restrict__ T3, const float4* __restrict__ K, float4* S, int i, int nb,
                           int it0, int tid, uint64_t pol_keep) {
  int pa[G], pb[G];
  bool special[G];
  float2 w[G];
  float4 k0[G], xs[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int slot = (it0 + g) * kThreads + tid;
    slot_positions(slot, pa[g], pb[g], special[g]);
    w[g] = T3[slot];
    k0[g] = K[slot];
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float2 A = X[padi(pa[g])], Bv = X[padi(pb[g])];
    if (!special[g]) {
      float2 xk, xm;
      tangle(A, Bv, w[g], xk, xm);
      xs[g] = make_float4(xk.x, xk.y, xm.x, xm.y);
    } else {
      // k = 0: X[0] = Re + Im, X[N] = Re - Im (both real); k = N/2: X = conj(Z).  All doubled like the general slots.
      xs[g] = make_float4(2.f * (A.x + A.y), 2.f * (A.x - A.y), 2.f * Bv.x, -2.f * Bv.y);
    }
  }
  if (i + 1 < nb) {
#pragma unroll
    for (int g = 0; g < G; ++g) st_f4_hint(S + (size_t)i * kSlots + (it0 + g) * kThreads + tid, xs[g], pol_keep);
  }
  float2 yk[G], ym[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    ym[g] = cmul(make_float2(xs[g].z, xs[g].w), make_float2(k0[g].z, k0[g].w));
    yk[g] = special[g] ? make_float2(xs[g].x * k0[g].x, xs[g].y * k0[g].y)
                       : cmul(make_float2(xs[g].x, xs[g].y), make_float2(k0[g].x, k0[g].y));
  }
  for (int j = 0; j < i; ++j) {
    float4 sj[G], kd[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int slot = (it0 + g) * kThreads + tid;
      sj[g] = ld_f4_hint(S + (size_t)j * kSlots + slot, pol_keep);
      kd[g] = K[(size_t)(i - j) * kSlots + slot];
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      ym[g] = cfma(make_float2(sj[g].z, sj[g].w), make_float2(kd[g].z, kd[g].w), ym[g]);
      yk[g] = special[g] ? make_float2(fmaf(sj[g].x, kd[g].x, yk[g].x), fmaf(sj[g].y, kd[g].y, yk[g].y))
                         : cfma(make_float2(sj[g].x, sj[g].y), make_float2(kd[g].x, kd[g].y), yk[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (!special[g]) {
      float2 za, zb;
      tangle(yk[g], ym[g], cconj(w[g]), za, zb);
      X[padi(pa[g])] = za;
      X[padi(pb[g])] = zb;
    } else {
      X[padi(pa[g])] = make_float2(yk[g].x + yk[g].y, yk[g].x - yk[g].y);
      X[padi(pb[g])] = make_float2(2.f * ym[g].x, -2.f * ym[g].y);
    }
  }
}

// Filter-spectrum epilogue for one slot: the doubled spectrum scaled by `scale` (= 1 / (8 N)), in slot layout
LC_HD float4 spectrum_slot(const float2* X, const float2* __restrict__ T3, int slot, float scale) {
  int pa, pb;
  bool special;
  slot_positions(slot, pa, pb, special);
  const float2 A = X[padi(pa)], Bv = X[padi(pb)];
  if (!special) {
    float2 xk, xm;
    tangle(A, Bv, T3[slot], xk, xm);
    return make_float4(xk.x * scale, xk.y * scale, xm.x * scale, xm.y * scale);
  }
  return make_float4(2.f * (A.x + A.y) * scale, 2.f * (A.x - A.y) * scale, 2.f * Bv.x * scale, -2.f * Bv.y * scale);
}

}  // namespace lc
}  // namespace dcb
