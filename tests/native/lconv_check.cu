// CPU emulation of lconv.cu's kernels (one CTA, thread by thread, a barrier = the end of a loop over tid) against a
// direct O(L^2) causal convolution in double.  Test infrastructure: built and run by tests/test_lconv_host.py with
//   nvcc -O2 -o <tmp>/lconv_check tests/native/lconv_check.cu && <tmp>/lconv_check
// It validates the index algebra of the blocked real-FFT convolution (slot layout, digit-reversed pairing, block
// distances, scaling) without a GPU; the GPU tests then compare the same code, compiled for sm_100a, with the oracle.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../deepchopper_b200/csrc/lconv_core.cuh"

using namespace dcb::lc;

static std::vector<float2> make_tw() {
  std::vector<float2> tw(kTwTotal);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < kTwTotal; ++i) {
    double c, s;
    if (i < kTwT2) {
      c = cos(-2.0 * PI * i / kP);
      s = sin(-2.0 * PI * i / kP);
    } else if (i < kTwT3) {
      c = cos(-2.0 * PI * (i - kTwT2) / kH);
      s = sin(-2.0 * PI * (i - kTwT2) / kH);
    } else {
      const double a = -2.0 * PI * slot_freq(i - kTwT3) / (2.0 * kP);
      c = sin(a);
      s = -cos(a);
    }
    tw[i] = make_float2((float)c, (float)s);
  }
  return tw;
}

static void spectra(const std::vector<float>& k, float D, const std::vector<float2>& tw, int nbK, std::vector<float4>& K) {
  std::vector<float2> X(kXFloat2);
  K.assign((size_t)nbK * kSlots, make_float4(0, 0, 0, 0));
  for (int d = 0; d < nbK; ++d) {
    auto tap = [&](int m) -> float {
      int u;
      if (m < kP) u = d * kP + m;
      else if (m == kP) return 0.f;
      else u = d * kP + m - 2 * kP;
      float v = (u >= 0 && u < (int)k.size()) ? k[u] : 0.f;
      if (u == 0) v += D;
      return v;
    };
    for (int tid = 0; tid < kThreads; ++tid)
      for (int n = tid; n < kH; n += kThreads)
        prologue_store_full(X.data(), tw.data() + kTwT1, n, make_float2(tap(2 * n), tap(2 * n + 1)),
                            make_float2(tap(2 * (n + kH)), tap(2 * (n + kH) + 1)));
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      radix16_pass<0, false>(X.data(), t.p0a, t.p0b, tid);
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      radix16_pass<1, false>(X.data(), t.p1a, t.p1b, tid);
    }
    for (int tid = 0; tid < kThreads; ++tid) radix16_pass<2, false>(X.data(), make_float2(1, 0), make_float2(1, 0), tid);
    for (int slot = 0; slot < kSlots; ++slot)
      K[(size_t)d * kSlots + slot] = spectrum_slot(X.data(), tw.data() + kTwT3, slot, 1.0f / (8.0f * kP));
  }
}

static void run_row(const std::vector<uint16_t>& vv, const std::vector<uint16_t>& gate, std::vector<uint16_t>& y, int L,
                    const std::vector<float4>& K, const std::vector<float2>& tw) {
  std::vector<float2> X(kXFloat2);
  std::vector<float4> S((size_t)(kMaxBlocks - 1) * kSlots);
  const int nb = (L + kP - 1) / kP;
  const float2* T3 = tw.data() + kTwT3;
  const uint32_t* v32 = reinterpret_cast<const uint32_t*>(vv.data());
  const uint32_t* g32 = reinterpret_cast<const uint32_t*>(gate.data());
  uint32_t* y32 = reinterpret_cast<uint32_t*>(y.data());
  for (int i = 0; i < nb; ++i) {
    const int n0 = i * kH;
    const int lim = L / 2 - n0;
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      uint32_t zraw[16];
      for (int m = 0; m < 16; ++m) {
        const int n = tid + 256 * m;
        zraw[m] = n < lim ? v32[n0 + n] : 0u;
      }
      fwd_pass0_fused(X.data(), zraw, t, tid);
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      radix16_pass<1, false>(X.data(), t.p1a, t.p1b, tid);
    }
    for (int tid = 0; tid < kThreads; ++tid) radix16_pass<2, false>(X.data(), make_float2(1, 0), make_float2(1, 0), tid);
    for (int tid = 0; tid < kThreads; ++tid)
      for (int it = 0; it < kSlots / kThreads; it += 4) pointwise_group<4>(X.data(), T3, K.data(), S.data(), i, nb, it, tid, 0);
    for (int tid = 0; tid < kThreads; ++tid) radix16_pass<2, true>(X.data(), make_float2(1, 0), make_float2(1, 0), tid);
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      radix16_pass<1, true>(X.data(), t.p1a, t.p1b, tid);
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      const ThreadTw t = load_thread_tw(tw.data(), tid);
      uint32_t out[16];
      inv_pass0_fused(X.data(), t, tid, [&](int m) -> uint32_t {
        const int n = tid + 256 * m;
        return n < lim ? g32[n0 + n] : 0u;
      }, out);
      for (int m = 0; m < 16; ++m) {
        const int n = tid + 256 * m;
        if (n < lim) y32[n0 + n] = out[m];
      }
    }
  }
}

int main() {
  const std::vector<float2> tw = make_tw();
  // every position of the shared array belongs to exactly one slot
  {
    std::vector<int> seen(kP, 0);
    for (int s = 0; s < kSlots; ++s) {
      int pa, pb;
      bool sp;
      slot_positions(s, pa, pb, sp);
      seen[pa]++;
      seen[pb]++;
    }
    for (int p = 0; p < kP; ++p)
      if (seen[p] != 1) {
        printf("FAIL: position %d covered %d times\n", p, seen[p]);
        return 1;
      }
  }
  srand(1234);
  auto rnd = []() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; };
  const int Lmax = 32770;
  std::vector<float> k(Lmax);
  for (int t = 0; t < Lmax; ++t) k[t] = rnd() * (expf(-(float)t / 3000.f) + 0.05f);
  const float D = 0.37f;
  std::vector<float4> K;
  spectra(k, D, tw, kMaxBlocks, K);
  const int Ls[] = {128, 4224, 8192, 8320, 16384, 19968, 32768};
  int bad = 0;
  for (int L : Ls) {
    std::vector<uint16_t> vv(L), gate(L), y(L, 0);
    std::vector<double> vd(L), gd(L);
    for (int t = 0; t < L; ++t) {
      vv[t] = (uint16_t)f32_to_bf16_bits(rnd());
      gate[t] = (uint16_t)f32_to_bf16_bits(rnd() + 1.5f);
      vd[t] = bf16_bits_to_f32(vv[t]);
      gd[t] = bf16_bits_to_f32(gate[t]);
    }
    run_row(vv, gate, y, L, K, tw);
    double maxerr = 0, rms = 0;
    for (int t = 0; t < L; ++t) {
      double acc = D * vd[t];
      for (int s = 0; s <= t; ++s) acc += vd[s] * (double)k[t - s];
      const double want = acc * gd[t];
      const double got = bf16_bits_to_f32(y[t]);
      const double err = fabs(got - want) - fabs(want) * 0.004;  // bf16 output rounding
      if (err > maxerr) maxerr = err;
      rms += want * want;
    }
    rms = sqrt(rms / L);
    printf("L=%d: max excess error %.3g (signal rms %.3g)\n", L, maxerr, rms);
    if (!(maxerr < 1e-3 * rms + 1e-4)) {
      printf("FAIL at L=%d\n", L);
      bad = 1;
    }
  }
  if (!bad) printf("OK\n");
  return bad;
}
