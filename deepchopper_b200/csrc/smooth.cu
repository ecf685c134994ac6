// Smoothing / interval / chop-coordinate kernels.
//
// Replace, per read, the reference's CPU chain
//   majority_voting        src/smooth/utils.rs:48-97
//   get_label_region       src/utils.rs:671-695         (start==0 sentinel quirk kept: position 0 is masked)
//   smooth_and_select_intervals  src/smooth/predict.rs:186-209
//   generate_unmaped_intervals + _split_records_by_remove_internal + chop-type gate
//                          src/output/split.rs:60-136,171-201,260-292
//   process_chunk gating   src/bin/predict.rs:141-164
//
// Labels are packed to a bit stream (32 bases per word); the majority vote over the default 21-wide window is a
// bit-sliced carry-save adder tree on those words (all 32 positions of a word at once).  Two kernels, bit-exact with
// each other and with the oracle:
//   smooth_chop_kernel  one WARP per read, a word per lane, 1024 bases per step, neighbours by shuffle, run boundaries
//                       with shifts / ballots, ordered emission by warp prefix sum.  Takes int8 labels or fp32 logits,
//                       any read length, optional smoothed-label output; used for small launches (a batch inside
//                       predict), the logits form, majority_voting, and reads beyond 32768 bases.
//   smooth_tile_kernel  one THREAD per word, a CTA per 64 reads (int8 labels, big launches): see its header below.
// HBM traffic is the 1 byte per base of the labels (8 B/base when reading fp32 logits) plus <= a few dozen bytes of
// coordinates per read.
#include "common.cuh"

namespace dcb {

// 4 int8 labels -> bits 28..31 (bit 28 + i = labels[i] == 1), exact for any byte values
__device__ __forceinline__ uint32_t pack4_top(uint32_t x) {
  const uint32_t t = x ^ 0x01010101u;  // a zero byte <=> label == 1
  // bit 7 of every zero byte: (t & 0x7f) + 0x7f sets bit 7 unless the low seven bits are zero (no carry leaves a byte),
  // or-ing t adds the bytes >= 0x80
  const uint32_t z = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
  return z * 0x00204081u;  // bits 7, 15, 23, 31 -> 28, 29, 30, 31 (the other partial products land below bit 24)
}
__device__ __forceinline__ uint32_t pack16(const uint4 v) {
  // 16 int8 labels -> 16 bits (bit i = labels[i] == 1)
  return (pack4_top(v.x) >> 28) | ((pack4_top(v.y) >> 24) & 0xf0u) | ((pack4_top(v.z) >> 20) & 0xf00u) |
         ((pack4_top(v.w) >> 16) & 0xf000u);
}
// 32 int8 labels -> 32 bits.  When every byte is 0 or 1 (what the classifier writes; checked here, not assumed) a byte IS
// its bit and a dot product with (1, 2, 4, 8 | 16, 32, 64, 128) gathers eight of them: 13 instructions instead of 65.
__device__ __forceinline__ uint32_t pack32(const uint4 v0, const uint4 v1) {
  const uint32_t other = ((v0.x | v0.y | v0.z | v0.w) | (v1.x | v1.y | v1.z | v1.w)) & 0xfefefefeu;
  if (other == 0u) {
    const uint32_t b0 = __dp4a(v0.y, 0x80402010u, __dp4a(v0.x, 0x08040201u, 0u));
    const uint32_t b1 = __dp4a(v0.w, 0x80402010u, __dp4a(v0.z, 0x08040201u, 0u));
    const uint32_t b2 = __dp4a(v1.y, 0x80402010u, __dp4a(v1.x, 0x08040201u, 0u));
    const uint32_t b3 = __dp4a(v1.w, 0x80402010u, __dp4a(v1.z, 0x08040201u, 0u));
    return b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
  }
  return pack16(v0) | (pack16(v1) << 16);
}

#define DCB_FA(a, b, c, s, cy)        \
  {                                   \
    uint32_t _a = (a), _b = (b), _c = (c); \
    s = _a ^ _b ^ _c;                 \
    cy = (_a & _b) | (_c & (_a ^ _b)); \
  }
#define DCB_HA(a, b, s, cy)  \
  {                          \
    uint32_t _a = (a), _b = (b); \
    s = _a ^ _b;             \
    cy = _a & _b;            \
  }

// Majority over the full 21-wide window for the 32 positions of word c (p/n: previous/next word).
// Bit-sliced population count: 16 full adders + 2 half adders, then count >= 11.
__host__ __device__ __forceinline__ uint32_t shl_in(uint32_t lo, uint32_t hi, int s) {  // (hi<<s)|(lo>>(32-s)), 0<s<32
  return (hi << s) | (lo >> (32 - s));
}
__host__ __device__ __forceinline__ uint32_t shr_in(uint32_t lo, uint32_t hi, int s) {  // (lo>>s)|(hi<<(32-s)), 0<s<32
  return (lo >> s) | (hi << (32 - s));
}

__host__ __device__ inline uint32_t majority21(uint32_t p, uint32_t c, uint32_t n) {
  uint32_t v[21];
#pragma unroll
  for (int d = 1; d <= 10; ++d) {
    v[10 - d] = shl_in(p, c, d);  // label at position b-d
    v[10 + d] = shr_in(c, n, d);  // label at position b+d
  }
  v[10] = c;
  uint32_t o[7], t[10], f[5], e[2];
  // ones: 21 -> 7 sums + 7 twos
#pragma unroll
  for (int i = 0; i < 7; ++i) DCB_FA(v[3 * i], v[3 * i + 1], v[3 * i + 2], o[i], t[i]);
  uint32_t o7, o8, b0;
  DCB_FA(o[0], o[1], o[2], o7, t[7]);
  DCB_FA(o[3], o[4], o[5], o8, t[8]);
  DCB_FA(o7, o8, o[6], b0, t[9]);
  // twos: 10
  uint32_t t10, t11, t12, t13, b1;
  DCB_FA(t[0], t[1], t[2], t10, f[0]);
  DCB_FA(t[3], t[4], t[5], t11, f[1]);
  DCB_FA(t[6], t[7], t[8], t12, f[2]);
  DCB_FA(t10, t11, t12, t13, f[3]);
  DCB_HA(t13, t[9], b1, f[4]);
  // fours: 5
  uint32_t f5, b2;
  DCB_FA(f[0], f[1], f[2], f5, e[0]);
  DCB_FA(f5, f[3], f[4], b2, e[1]);
  // eights: 2
  uint32_t b3, b4;
  DCB_HA(e[0], e[1], b3, b4);
  // count >= 11  (11 = 0b01011)
  return b4 | (b3 & (b2 | (b1 & b0)));
}

// Generic half-window h (1..31): per-position popcount of the window bits.
__device__ __forceinline__ uint32_t majority_generic(uint32_t p, uint32_t c, uint32_t n, int h) {
  const int W = 2 * h + 1;
  const uint64_t mask = (W >= 64) ? ~0ull : ((1ull << W) - 1ull);
  const uint64_t lo = ((uint64_t)c << 32) | p;  // positions -32..31
  const uint64_t hi = ((uint64_t)n << 32) | c;  // positions 0..63
  uint32_t out = 0;
#pragma unroll 4
  for (int b = 0; b < 32; ++b) {
    int off = 32 + b - h;  // bit offset of window start inside the 96-bit p:c:n, >= 1
    uint64_t win = (off < 32) ? ((lo >> off) | (off ? (uint64_t)n << (64 - off) : 0ull)) : (hi >> (off - 32));
    // for off<32 the window may need bits of n above position 63-off: handled by the OR above
    int c1 = __popcll(win & mask);
    out |= (uint32_t)(c1 > h) << b;
  }
  return out;
}

struct SmoothArgs {
  const int8_t* labels;
  const float* logits;
  int64_t total;  // labels: bytes, logits: tokens
  const int64_t* starts;
  const int32_t* lens;
  const int32_t* qual_lens;
  int64_t R;
  dcb200_chop_params p;
  int32_t* n_adapter;
  int32_t* adapter_iv;
  int32_t* n_keep;
  int32_t* keep_iv;
  uint8_t* action;
  int8_t* smoothed;  // optional: majority_voting output, layout of labels
  // the warp-per-read kernel as the second launch behind the tile kernel: it visits the reads sel[0 .. *sel_count) only
  // (those the tile kernel listed because they are longer than its shared arrays hold); null = all reads
  int32_t* sel;
  int32_t* sel_count;
};

template <bool LOGITS>
__device__ __forceinline__ uint32_t label_at(const SmoothArgs& a, int64_t g) {
  if (g < 0 || g >= a.total) return 0;
  if (LOGITS) {
    float2 l = reinterpret_cast<const float2*>(a.logits)[g];
    return l.y > l.x;
  }
  return a.labels[g] == 1;
}

// 32 consecutive "aligned" elements starting at element index g0 -> bits.
template <bool LOGITS>
__device__ __forceinline__ uint32_t load_word(const SmoothArgs& a, int64_t g0, bool wanted) {
  if (!wanted || g0 >= a.total || g0 + 32 <= 0) return 0;
  if (g0 >= 0 && g0 + 32 <= a.total) {
    if (LOGITS) {
      const float4* q = reinterpret_cast<const float4*>(a.logits + 2 * g0);
      uint32_t r = 0;
      if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float4 v = __ldg(q + i);
          r |= (uint32_t)(v.y > v.x) << (2 * i);
          r |= (uint32_t)(v.w > v.z) << (2 * i + 1);
        }
      } else {
        const float2* q2 = reinterpret_cast<const float2*>(a.logits + 2 * g0);
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          float2 v = __ldg(q2 + i);
          r |= (uint32_t)(v.y > v.x) << i;
        }
      }
      return r;
    } else {
      const uint4* q = reinterpret_cast<const uint4*>(a.labels + g0);  // g0 chosen so the address is 32B aligned
      uint4 v0 = __ldg(q), v1 = __ldg(q + 1);
      return pack32(v0, v1);
    }
  }
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r |= label_at<LOGITS>(a, g0 + i) << i;
  return r;
}

// The int8 path's word load split in two so that the 32 bytes can be in flight while other work runs: issue() starts the
// two 16-byte loads (or resolves a word at the edge of the buffer byte by byte), word() packs them.
struct RawWord {
  uint4 v0, v1;
  uint32_t slow;
  bool fast;
};
__device__ __forceinline__ RawWord issue_word(const SmoothArgs& a, int64_t g0, bool wanted) {
  RawWord r;
  r.fast = false;
  r.slow = 0;
  r.v0 = r.v1 = make_uint4(0u, 0u, 0u, 0u);
  if (!wanted || g0 >= a.total || g0 + 32 <= 0) return r;
  if (g0 >= 0 && g0 + 32 <= a.total) {
    const uint4* q = reinterpret_cast<const uint4*>(a.labels + g0);  // g0 chosen so the address is 32B aligned
    r.v0 = __ldg(q);
    r.v1 = __ldg(q + 1);
    r.fast = true;
    return r;
  }
  for (int i = 0; i < 32; ++i) r.slow |= label_at<false>(a, g0 + i) << i;
  return r;
}
__device__ __forceinline__ uint32_t finish_word(const RawWord& r) {
  return r.fast ? pack32(r.v0, r.v1) : r.slow;
}

__device__ __forceinline__ uint32_t mask_below(int n, int word) {  // bits of word with position < n
  const int d = n - word * 32;  // read positions are 31-bit (lens are int32): no 64-bit arithmetic in the hot loop
  if (d >= 32) return 0xffffffffu;
  if (d <= 0) return 0u;
  return (1u << d) - 1u;
}

// tie -> keep original; else majority value.  c1 ones among size.
__device__ __forceinline__ uint32_t vote(int c1, int size, uint32_t orig) {
  int c0 = size - c1;
  return c1 == c0 ? orig : (c1 > c0 ? 1u : 0u);
}

// Per-read tail of both kernels (one thread): the approved-interval rule, the kept intervals and the chop decision.
// `ad`: the read's first `total` adapter intervals (the global table, or a shared-memory copy of its head).
__device__ __forceinline__ void finish_read(const SmoothArgs& a, int64_t r, int n, bool skip, int total,
                                            const volatile int32_t* ad) {
  const int approved = a.p.approved_interval_number;
  if (total > approved) total = 0;  // src/smooth/predict.rs:204-206
  int nk = 0;
  uint8_t act = DCB200_ACTION_PASSTHROUGH;
  const bool qual_ok = !a.qual_lens || a.qual_lens[r] == n;  // src/bin/predict.rs:160-164
  if (!skip && total > 0 && total <= a.p.max_process_intervals && qual_ok) {
    if (a.p.output_chopped_seqs) {
      act = DCB200_ACTION_ADAPTERS;
    } else {
      // generate_unmaped_intervals, src/output/split.rs:260-292
      const int mc = a.p.min_read_length_after_chop;
      int before = 0, cur = 0, first_len = -1;
      int32_t* keep = a.keep_iv + r * (approved + 1) * 2;
      for (int i = 0; i < total; ++i) {
        const int s = ad[2 * i], e = ad[2 * i + 1];
        if (cur < s) {
          ++before;
          if (s - cur >= mc) {
            keep[2 * nk] = cur;
            keep[2 * nk + 1] = s;
            if (nk == 0) first_len = s - cur;
            ++nk;
          }
        }
        cur = e;
      }
      if (cur < n - 1) {
        ++before;
        if (n - 1 - cur >= mc) {
          keep[2 * nk] = cur;
          keep[2 * nk + 1] = n - 1;
          if (nk == 0) first_len = n - 1 - cur;
          ++nk;
        }
      }
      const bool terminal = before == 1;  // src/output/split.rs:185-189
      const int ct = a.p.chop_type;
      if ((ct == DCB200_CHOP_TERMINAL && !terminal) || (ct == DCB200_CHOP_INTERNAL && terminal) ||
          (nk > 0 && first_len == n)) {
        nk = 0;  // rebuilt, uncut record: src/output/split.rs:191-201
        act = DCB200_ACTION_UNCHOPPED;
      } else {
        act = terminal ? DCB200_ACTION_CHOP_T : DCB200_ACTION_CHOP_I;
      }
    }
  }
  a.n_adapter[r] = skip ? 0 : total;
  a.n_keep[r] = nk;
  a.action[r] = act;
}

template <bool LOGITS, int HFIX>
__global__ void __launch_bounds__(256) smooth_chop_kernel(const SmoothArgs a) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  int window = a.p.smooth_window_size;
  if ((window & 1) == 0) window += 1;  // src/smooth/utils.rs:50-54
  const int h = HFIX >= 0 ? HFIX : window / 2;
  const int W = 2 * h + 1;
  const int approved = a.p.approved_interval_number;

  // Software pipeline across reads (int8 labels): while read r is processed, the first 1 KB of read r + nwarps is already
  // in flight and the (start, length) of read r + 2 nwarps is being fetched -- a warp's dependent chain
  // metadata -> labels -> votes -> coordinates would otherwise leave one 1 KB load per warp in flight.
  int64_t start_n = 0, n_n = 0, start_nn = 0, n_nn = 0;
  RawWord pre;
  pre.fast = false;
  pre.slow = 0;
  pre.v0 = pre.v1 = make_uint4(0u, 0u, 0u, 0u);
  auto plan_read = [&](int64_t st, int64_t nn, int& mis_o, int64_t& abase_o) -> bool {  // false: nothing to load
    mis_o = 0;
    abase_o = st;
    if (!LOGITS) {
      mis_o = (int)((reinterpret_cast<uintptr_t>(a.labels) + (uintptr_t)st) & 31);
      abase_o = st - mis_o;
    }
    return nn > 0 && ((a.smoothed != nullptr) || nn >= a.p.min_read_length);
  };
  const int64_t Rn = a.sel ? (int64_t)*a.sel_count : a.R;            // reads this launch visits
  auto rid = [&](int64_t i) -> int64_t { return a.sel ? (int64_t)a.sel[i] : i; };
  if (warp_global < Rn) {
    start_n = a.starts[rid(warp_global)];
    n_n = a.lens[rid(warp_global)];
    if (!LOGITS) {
      int mis0;
      int64_t ab0;
      const bool live = plan_read(start_n, n_n, mis0, ab0);
      pre = issue_word(a, ab0 + 32 * (int64_t)lane, live && lane <= n_n / 32 + 1);
    }
    if (warp_global + nwarps < Rn) {
      start_nn = a.starts[rid(warp_global + nwarps)];
      n_nn = a.lens[rid(warp_global + nwarps)];
    }
  }
  for (int64_t ri = warp_global; ri < Rn; ri += nwarps) {
    const int64_t r = rid(ri);
    const int64_t start = start_n;
    const int n = (int)n_n;
    const RawWord cur = pre;
    start_n = start_nn;
    n_n = n_nn;
    if (!LOGITS && ri + nwarps < Rn) {
      int mis1;
      int64_t ab1;
      const bool live = plan_read(start_n, n_n, mis1, ab1);
      pre = issue_word(a, ab1 + 32 * (int64_t)lane, live && lane <= n_n / 32 + 1);
    }
    if (ri + 2 * nwarps < Rn) {
      start_nn = a.starts[rid(ri + 2 * nwarps)];
      n_nn = a.lens[rid(ri + 2 * nwarps)];
    }
    int total = 0;
    const bool skip = (!a.smoothed) && (n < a.p.min_read_length);  // src/bin/predict.rs:146-148
    if (n > 0 && !skip) {
      // alignment of the packed stream: label bytes are fetched as 32B-aligned groups
      int mis = 0;
      int64_t abase = start;
      if (!LOGITS) {
        mis = (int)((reinterpret_cast<uintptr_t>(a.labels) + (uintptr_t)start) & 31);
        abase = start - mis;
      }
      const int NW = n / 32 + 1;  // words 0..n/32 cover positions 0..n (position n closes a trailing run)
      const int nchunks = (NW + 31) / 32;

      // right-edge window (i + h + 1 > n): [max(0,n-W), n), identical for all such positions; its count of ones is
      // taken from the packed words of the chunk(s) that hold right-edge positions (no extra byte loads)
      const int sizeR = (int)(n < W ? n : W);
      int cR = 0;
      const int redge = n - h > 0 ? n - h : 0;  // first right-edge position

      uint32_t A0 = LOGITS ? load_word<LOGITS>(a, abase + 32 * (int64_t)lane, lane <= NW) : finish_word(cur);
      uint32_t prev_word = 0;      // raw word k-1 for lane 0
      uint32_t prev_S_last = 0;    // smoothed bit of position 32k-1 for lane 0
      int open_start = -1;         // most recent run start seen in earlier chunks

      for (int c = 0; c < nchunks; ++c) {
        const int k = 32 * c + lane;
        uint32_t A1 = load_word<LOGITS>(a, abase + 32 * (int64_t)(k + 32), (k + 32) <= NW);
        // raw read-relative words
        uint32_t up = __shfl_down_sync(0xffffffffu, A0, 1);
        const uint32_t A1_0 = __shfl_sync(0xffffffffu, A1, 0);
        const uint32_t A1_1 = __shfl_sync(0xffffffffu, A1, 1);
        if (lane == 31) up = A1_0;
        uint32_t w = (mis ? shr_in(A0, up, mis) : A0) & mask_below(n, k);
        uint32_t wn0 = (mis ? shr_in(A1_0, A1_1, mis) : A1_0) & mask_below(n, 32 * (c + 1));
        uint32_t wnext = __shfl_down_sync(0xffffffffu, w, 1);
        if (lane == 31) wnext = wn0;
        uint32_t wprev = __shfl_up_sync(0xffffffffu, w, 1);
        if (lane == 0) wprev = prev_word;
        prev_word = __shfl_sync(0xffffffffu, w, 31);

        // ---- majority vote -------------------------------------------------------------------
        uint32_t S;
        if (h == 0) S = w;
        else if (HFIX == 10) S = majority21(wprev, w, wnext);
        else S = majority_generic(wprev, w, wnext, h);
        // left edge: positions i < h with i+h+1 <= n use the clipped window [0, i+h+1)
        if (c == 0 && h > 0) {
          const uint32_t w0 = __shfl_sync(0xffffffffu, w, 0);
          const uint32_t w1 = __shfl_sync(0xffffffffu, w, 1);
          const uint64_t X = ((uint64_t)w1 << 32) | w0;
          const int size = lane + h + 1;
          uint32_t bit = 0;
          const bool is_left = lane < h && size <= n;
          if (is_left) bit = vote(__popcll(X & ((1ull << size) - 1ull)), size, (w0 >> lane) & 1u);
          const uint32_t lmask = __ballot_sync(0xffffffffu, is_left);
          const uint32_t lbits = __ballot_sync(0xffffffffu, bit != 0);
          if (lane == 0) S = (S & ~lmask) | (lbits & lmask);
        }
        // right edge: positions >= max(0, n-h) share one window
        if (h > 0 && 1024 * (c + 1) > redge) {  // (warp-uniform) this chunk holds right-edge positions
          const int lo_pos = n - sizeR;         // window = [lo_pos, n): at most 21 bits, within words 32c-1 .. 32(c+1)
          int cnt = __popc(w & ~mask_below(lo_pos, k));
          if (lane == 0 && c > 0) cnt += __popc(wprev & ~mask_below(lo_pos, k - 1));
          if (lane == 31) cnt += __popc(wn0 & ~mask_below(lo_pos, k + 1));
          cR = __reduce_add_sync(0xffffffffu, cnt);
        }
        {
          const int lo = 32 * k;
          if (lo + 32 > redge && h > 0) {
            uint32_t em = mask_below(n, k);
            if (redge > lo) em &= ~((1u << (redge - lo)) - 1u);
            const int c0 = sizeR - cR;
            const uint32_t val = cR == c0 ? w : (cR > c0 ? 0xffffffffu : 0u);
            S = (S & ~em) | (val & em);
          }
        }
        S &= mask_below(n, k);
        if (a.smoothed) {
          const int lo = 32 * k;
          for (int b = 0; b < 32 && lo + b < n; ++b) a.smoothed[start + lo + b] = (int8_t)((S >> b) & 1u);
        }
        if (a.smoothed) {  // majority_voting mode: no interval pass
          A0 = A1;
          continue;
        }
        if (k == 0) S &= ~1u;  // src/utils.rs:677-684: `start == 0` is the "no open run" sentinel

        // ---- runs ------------------------------------------------------------------------------
        if (!__any_sync(0xffffffffu, S != 0u) && prev_S_last == 0u) {  // nothing starts, nothing ends in this chunk
          A0 = A1;
          continue;
        }
        uint32_t cin = __shfl_up_sync(0xffffffffu, S, 1) >> 31;
        if (lane == 0) cin = prev_S_last;
        prev_S_last = __shfl_sync(0xffffffffu, S, 31) >> 31;
        const uint32_t Sprev = (S << 1) | cin;
        const uint32_t st = S & ~Sprev;
        uint32_t en = ~S & Sprev;
        const int kbase = 32 * k;
        int scan = st ? kbase + 31 - __clz(st) : -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, scan, d);
          if (lane >= d) scan = max(scan, t);
        }
        int excl = __shfl_up_sync(0xffffffffu, scan, 1);
        if (lane == 0) excl = -1;
        excl = max(excl, open_start);
        open_start = max(open_start, __shfl_sync(0xffffffffu, scan, 31));

        int cnt_local = 0;
        for (uint32_t e = en; e; e &= e - 1) {
          const int b = __ffs(e) - 1;
          const uint32_t below = st & ((1u << b) - 1u);
          const int s = below ? kbase + 31 - __clz(below) : excl;
          cnt_local += (kbase + b - s) >= a.p.min_interval_size;
        }
        if (__any_sync(0xffffffffu, cnt_local)) {
          int incl = cnt_local;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
          }
          int slot = total + incl - cnt_local;
          total += __shfl_sync(0xffffffffu, incl, 31);
          for (uint32_t e = en; e; e &= e - 1) {
            const int b = __ffs(e) - 1;
            const uint32_t below = st & ((1u << b) - 1u);
            const int s = below ? kbase + 31 - __clz(below) : excl;
            if ((kbase + b - s) >= a.p.min_interval_size) {
              if (slot < approved) {
                a.adapter_iv[(r * approved + slot) * 2 + 0] = s;
                a.adapter_iv[(r * approved + slot) * 2 + 1] = kbase + b;
              }
              ++slot;
            }
          }
        }
        A0 = A1;
      }
    }
    if (a.smoothed) continue;
    __syncwarp();
    // ---- per-read decision (lane 0) ------------------------------------------------------------
    if (lane == 0) finish_read(a, r, n, skip, total, a.adapter_iv + r * approved * 2);
    __syncwarp();
  }
}

// ---- tile kernel: int8 labels, windows up to 63, no smoothed-label output ------------------------------------------
// The warp-per-read kernel above spends its second 1024-base step on 6 of 32 lanes for a typical 1.2 kb read (38
// lane-words) and pays every per-step cost twice.  Here a CTA owns a tile of 64 consecutive reads and its threads own
// WORDS (32 bases) of the tile, whatever read they belong to:
//   P1  thread per aligned raw word: two 16-byte loads, bytes -> bits                              -> A[] (shared)
//   P2  thread per word: funnel shift by the read's misalignment, mask beyond the read             -> Wd[]
//       (every read gets one extra all-zero word behind it: the "next word" of the read's last word and the
//        "previous word" of the following read's first one; there is a zero word in front of the first read too)
//   P3  thread per word: 21-wide (or generic) majority from Wd[k-1], Wd[k], Wd[k+1]                 -> S[] (over A[])
//   P4  thread per READ: the clipped windows at the read's two edges, then the scalar run scan over its S words,
//       interval emission in order and the chop decision (the reference's loops, on 32-base words)
// A tile whose words do not fit the shared arrays is processed in sub-batches of whole reads.
// (measured at 2 M reads: 32 reads / 128 threads 1.59 ms, 64 / 128 1.62, 64 / 512 1.76, 128 / 256 1.48, cap 2048 1.63;
//  64 reads, 256 threads, cap 3072: 1.36 ms)
constexpr int kTileReads = 64;      // multiple of 32
constexpr int kTileThreads = 256;   // >= kTileReads
constexpr int kTileRW = kTileReads / 32;
constexpr int kTileCap = 3072;                 // a sub-batch = the reads whose first slot falls into one window of kTileCap slots
constexpr int kTileWords = kTileCap + 1032;    // + the rest of one maximal read (32768 bases: 1025 words + separator) + 2 guards
constexpr int kTileMaxLen = 32768;             // longest read the tile kernel takes (1025 words + separator)
constexpr int kIvHead = 4;                     // (the default max_process_intervals: more intervals than that pass the read through)

template <int HFIX>
__global__ void __launch_bounds__(kTileThreads, 6) smooth_tile_kernel(const SmoothArgs a) {
  __shared__ __align__(16) uint32_t A[kTileWords];  // raw aligned words, later the smoothed words S
  __shared__ uint32_t Wd[kTileWords];    // read-relative label words (P1 leaves (read in tile << 11) | word index here)
  __shared__ int64_t r_start[kTileReads];
  __shared__ int r_n[kTileReads], r_slots[kTileReads], r_pexcl[kTileReads + 1], r_mis[kTileReads];
  __shared__ int iv_head[kTileReads][2 * kIvHead];  // the first intervals of every read: the decision reads them back
  __shared__ int blk_first[(kTileWords + 31) / 32];  // the read that owns the first slot of every block of 32 slots
  __shared__ int warp_tot[kTileRW];
  const int tid = threadIdx.x;
  int window = a.p.smooth_window_size;
  if ((window & 1) == 0) window += 1;  // src/smooth/utils.rs:50-54
  const int h = HFIX >= 0 ? HFIX : window / 2;
  const int W = 2 * h + 1;
  const int approved = a.p.approved_interval_number;
  const int64_t n_tiles = (a.R + kTileReads - 1) / kTileReads;
  const uintptr_t lab0 = reinterpret_cast<uintptr_t>(a.labels);

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTileReads;
    const int nr = (int)min((int64_t)kTileReads, a.R - r0);
    // ---- read table: slots per read (n / 32 + 1 words + separator; 0 for reads that are not processed) and their
    //      exclusive prefix over the tile -------------------------------------------------------------------------------
    int slots = 0, incl = 0;
    if (tid < kTileReads) {
      if (tid < nr) {
        const int n = a.lens[r0 + tid];
        const int64_t st = a.starts[r0 + tid];
        r_start[tid] = st;
        r_n[tid] = n;
        r_mis[tid] = (int)((lab0 + (uintptr_t)st) & 31);  // label bytes are fetched as 32-byte aligned groups
        // (src/bin/predict.rs:146-148: short reads pass through; reads beyond the model's window do not fit the shared
        //  arrays: launch_smooth sends them through the warp-per-read kernel)
        if (n > 0 && n >= a.p.min_read_length && n <= kTileMaxLen) slots = n / 32 + 2;
      }
      incl = slots;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((tid & 31) >= d) incl += t;
      }
      if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < kTileReads) {
      int pe = incl - slots;
      for (int wq = 0; wq < (tid >> 5); ++wq) pe += warp_tot[wq];
      r_slots[tid] = slots;
      r_pexcl[tid] = pe;
      if (tid == kTileReads - 1) r_pexcl[kTileReads] = pe + slots;
    }
    __syncthreads();
    const int all_slots = r_pexcl[kTileReads];
    // ---- sub-batches: the reads whose first slot lies in [w0, w0 + cap); the last of them may reach beyond ----------
    for (int w0 = 0; w0 < all_slots; w0 += kTileCap) {
      // first / last member (reads with slots; a slotless read shares its prefix with the read after it and is skipped
      // by "largest read whose prefix is <= the slot" as long as the search stops at the last member)
      int first = -1, last = -1;
      {
        const bool mine = tid < kTileReads && r_slots[tid] > 0 && r_pexcl[tid] >= w0 && r_pexcl[tid] < w0 + kTileCap;
        const unsigned m = __ballot_sync(0xffffffffu, mine);
        if (tid < kTileReads && (tid & 31) == 0) warp_tot[tid >> 5] = (int)m;
        __syncthreads();
#pragma unroll
        for (int wq = 0; wq < kTileRW; ++wq) {
          const unsigned mw = (unsigned)warp_tot[wq];
          if (mw) {
            if (first < 0) first = 32 * wq + __ffs(mw) - 1;
            last = 32 * wq + 31 - __clz(mw);
          }
        }
      }
      if (first < 0) {  // (the previous sub-batch's last read covered this whole window)
        __syncthreads();
        continue;
      }
      const int base = r_pexcl[first];
      const int n_slots = r_pexcl[last] + r_slots[last] - base;  // <= cap + 1027
      // slot i (0-based) lives at index i + 1 of the shared arrays: index 0 is the zero word in front of the first read
      if (tid >= first && tid <= last && r_slots[tid] > 0) {
        const int o0 = r_pexcl[tid] - base;
        for (int b = (o0 + 31) >> 5; 32 * b < o0 + r_slots[tid]; ++b) blk_first[b] = tid;
      }
      __syncthreads();
      // ---- P1: raw aligned words (also for the separator slot: the last word's funnel shift may need it) ----------
      for (int i = tid; i < n_slots; i += kTileThreads) {
        int rr = blk_first[i >> 5];  // owner of slot 32 (i / 32); walk to the last read whose prefix is <= i
        while (rr < last && r_pexcl[rr + 1] - base <= i) ++rr;
        const int k = i - (r_pexcl[rr] - base);
        Wd[i + 1] = (uint32_t)((rr << 11) | k);
        A[i + 1] = load_word<false>(a, r_start[rr] - r_mis[rr] + 32 * (int64_t)k, true);
      }
      if (tid == 0) {
        A[0] = 0u;
        A[n_slots + 1] = 0u;
        Wd[0] = 0u;
        Wd[n_slots + 1] = 0u;
      }
      __syncthreads();
      // ---- P2: read-relative words -------------------------------------------------------------------------------
      for (int i = tid; i < n_slots; i += kTileThreads) {
        const uint32_t m = Wd[i + 1];
        const int rr = (int)(m >> 11), k = (int)(m & 2047u);
        const int n = r_n[rr];
        uint32_t w = 0u;
        if (k <= n / 32) {
          const int mis = r_mis[rr];
          const uint32_t a0 = A[i + 1];
          w = (mis ? shr_in(a0, A[i + 2], mis) : a0) & mask_below(n, k);
        }
        Wd[i + 1] = w;
      }
      __syncthreads();
      // ---- P3: majority vote with full windows (a separator's own value is never read) -----------------------------
      for (int i = tid; i < n_slots; i += kTileThreads) {
        const uint32_t w = Wd[i + 1], wp = Wd[i], wn = Wd[i + 2];
        uint32_t S;
        if (h == 0 || (w | wp | wn) == 0u) S = w;  // (no ones within reach: what real predictions mostly look like)
        else if (HFIX == 10) S = majority21(wp, w, wn);
        else S = majority_generic(wp, w, wn, h);
        A[i + 1] = S;
      }
      __syncthreads();
      // ---- P4: one thread per read: edges, runs, intervals, decision -------------------------------------------------
      // (meanwhile two of the six idle warps pull the label bytes of this CTA's next tile into L2, one request per line)
      if (w0 == 0 && tid >= kTileThreads - kTileReads && tile + gridDim.x < n_tiles) {
        const int64_t rn = (tile + gridDim.x) * kTileReads + (tid - (kTileThreads - kTileReads));
        if (rn < a.R) {
          const int n = a.lens[rn];
          if (n >= a.p.min_read_length && n <= kTileMaxLen) {
            const int8_t* q = a.labels + a.starts[rn];
            for (int o = 0; o < n; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
          }
        }
      }
      if (tid >= first && tid <= last && r_slots[tid] > 0) {
        const int64_t r = r0 + tid;
        const int n = r_n[tid];
        const int off = r_pexcl[tid] - base + 1;
        const int NW = n / 32 + 1;
        uint32_t* S = A + off;
        const uint32_t* w = Wd + off;
        if (h > 0) {
          // left edge: positions i < h with i + h + 1 <= n use the clipped window [0, i + h + 1)
          const uint64_t X = ((uint64_t)w[1] << 32) | w[0];
          uint32_t s0 = S[0];
          for (int b = 0; b < h && b + h + 1 <= n; ++b) {
            const int size = b + h + 1;
            const uint32_t bit = vote(__popcll(X & ((1ull << size) - 1ull)), size, (w[0] >> b) & 1u);
            s0 = (s0 & ~(1u << b)) | (bit << b);
          }
          S[0] = s0;
          // right edge: positions >= max(0, n - h) share the window [max(0, n - W), n)
          const int sizeR = n < W ? n : W;
          const int redge = n - h > 0 ? n - h : 0;
          const int lo_pos = n - sizeR;
          int cR = 0;
          for (int kw = lo_pos / 32; kw <= (n - 1) / 32; ++kw) cR += __popc(w[kw] & ~mask_below(lo_pos, kw));
          const int c0 = sizeR - cR;
          for (int kr = redge / 32; kr <= (n - 1) / 32; ++kr) {
            uint32_t em = mask_below(n, kr);
            if (redge > 32 * kr) em &= ~((1u << (redge - 32 * kr)) - 1u);
            const uint32_t val = cR == c0 ? w[kr] : (cR > c0 ? 0xffffffffu : 0u);
            S[kr] = (S[kr] & ~em) | (val & em);
          }
        }
        S[NW - 1] &= mask_below(n, NW - 1);
        S[0] &= ~1u;  // src/utils.rs:677-684: `start == 0` is the "no open run" sentinel
        int total = 0, open_start = -1;
        uint32_t cin = 0u;
        for (int k = 0; k < NW; ++k) {
          // four all-zero words with no run open: nothing starts or ends (most of a read)
          if (cin == 0u && ((off + k) & 3) == 0 && k + 4 <= NW) {
            const uint4 q = *reinterpret_cast<const uint4*>(S + k);
            if ((q.x | q.y | q.z | q.w) == 0u) {
              k += 3;
              continue;
            }
          }
          const uint32_t Sk = S[k];
          const uint32_t Sprev = (Sk << 1) | cin;
          const uint32_t st = Sk & ~Sprev;
          cin = Sk >> 31;
          for (uint32_t e = ~Sk & Sprev; e; e &= e - 1) {
            const int b = __ffs(e) - 1;
            const uint32_t below = st & ((1u << b) - 1u);
            const int s = below ? 32 * k + 31 - __clz(below) : open_start;
            if (32 * k + b - s >= a.p.min_interval_size) {
              if (total < approved) {
                a.adapter_iv[(r * approved + total) * 2 + 0] = s;
                a.adapter_iv[(r * approved + total) * 2 + 1] = 32 * k + b;
              }
              if (total < kIvHead) {
                iv_head[tid][2 * total] = s;
                iv_head[tid][2 * total + 1] = 32 * k + b;
              }
              ++total;
            }
          }
          if (st) open_start = 32 * k + 31 - __clz(st);
        }
        finish_read(a, r, n, false, total, total <= kIvHead ? iv_head[tid] : a.adapter_iv + r * approved * 2);
      }
      __syncthreads();
    }
    // reads without label words: shorter than min_read_length (src/bin/predict.rs:146-148) or empty
    if (tid < nr && r_slots[tid] == 0) {
      if (r_n[tid] <= kTileMaxLen) finish_read(a, r0 + tid, r_n[tid], r_n[tid] < a.p.min_read_length, 0, nullptr);
      else a.sel[atomicAdd(a.sel_count, 1)] = (int32_t)(r0 + tid);  // left to the warp-per-read kernel (launch_smooth)
    }
    __syncthreads();
  }
}

// Windows wider than 63: literal per-position recount into a scratch label buffer (then the fast
// kernel runs with window 1 on it).
__global__ void majority_naive_kernel(const int8_t* labels, const float* logits, int64_t total, const int64_t* starts,
                                      const int32_t* lens, int64_t R, int window, int8_t* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  if ((window & 1) == 0) window += 1;
  const int h = window / 2;
  for (int64_t r = warp_global; r < R; r += nwarps) {
    const int64_t start = starts[r], n = lens[r];
    for (int64_t i = lane; i < n; i += 32) {
      int64_t s = i - h < 0 ? 0 : i - h;
      int64_t e = i + h + 1 < n ? i + h + 1 : n;
      if (e == n && e - s < window) s = e - window < 0 ? 0 : e - window;
      int c1 = 0;
      for (int64_t j = s; j < e; ++j) {
        int64_t g = start + j;
        c1 += logits ? (logits[2 * g + 1] > logits[2 * g]) : (labels[g] == 1);
      }
      const int c0 = (int)(e - s) - c1;
      const int64_t g = start + i;
      const int orig = logits ? (logits[2 * g + 1] > logits[2 * g]) : (labels[g] == 1);
      out[g] = (int8_t)(c1 == c0 ? orig : (c1 > c0));
    }
  }
}

static int launch_smooth(dcb200_ctx* ctx, SmoothArgs a) {
  if (a.R == 0) return DCB200_OK;
  int window = a.p.smooth_window_size;
  if ((window & 1) == 0) window += 1;
  const int threads = 256;
  int64_t blocks64 = (a.R + 7) / 8;
  const int64_t cap = (int64_t)ctx->sm_count * 8 * 4;  // 8 resident 256-thread CTAs per SM, a few waves; grid-stride beyond
  int blocks = (int)(blocks64 < cap ? blocks64 : cap);
  const bool logits = a.logits != nullptr;
  ProfScope prof(ctx, K_SMOOTH);
  if (window > 63) {
    // rare parameterisation: recount literally, then run the interval pass with window 1
    dcb::DevBuf& scratch = ctx->buf("smooth_scratch");
    DCB_CHECK(scratch.reserve((size_t)a.total));
    majority_naive_kernel<<<blocks, threads, 0, ctx->stream>>>(a.labels, a.logits, a.total, a.starts, a.lens, a.R, window,
                                                             a.smoothed ? a.smoothed : scratch.as<int8_t>());
    DCB_LAUNCH_CHECK(ctx);
    if (a.smoothed) return DCB200_OK;
    a.labels = scratch.as<int8_t>();
    a.logits = nullptr;
    a.p.smooth_window_size = 1;
    smooth_chop_kernel<false, -1><<<blocks, threads, 0, ctx->stream>>>(a);
    DCB_LAUNCH_CHECK(ctx);
    return DCB200_OK;
  }
  // int8 labels -> coordinates: the tile kernel (thread per 32-base word) once there is at least a tile of 64 reads per SM;
  // below that (the ~800-read batches inside predict, a batch of 128 long reads) a CTA per 64 reads leaves the machine
  // empty and the warp-per-read kernel is 3-4 x faster (15 vs 54 us at 800 reads).  Option smooth_warp_kernel: 1 = never
  // the tile kernel, 2 = always (tests).
  const bool tile = ctx->smooth_warp_kernel == 2 || (ctx->smooth_warp_kernel == 0 && a.R >= (int64_t)kTileReads * ctx->sm_count);
  if (!logits && !a.smoothed && tile) {
    DCB_ARG(a.R <= INT_MAX);
    dcb::DevBuf& sel = ctx->buf("smooth_sel");
    DCB_CHECK(sel.reserve((size_t)(a.R + 1) * 4));
    a.sel_count = sel.as<int32_t>();
    a.sel = a.sel_count + 1;
    DCB_CUDA(cudaMemsetAsync(a.sel_count, 0, 4, ctx->stream));
    const int64_t tiles = (a.R + kTileReads - 1) / kTileReads;
    const int64_t tcap = (int64_t)ctx->sm_count * 24;  // ~6 resident CTAs per SM (36 KB of shared memory each), a few waves
    const int tblocks = (int)(tiles < tcap ? tiles : tcap);
    if (window == 21) smooth_tile_kernel<10><<<tblocks, kTileThreads, 0, ctx->stream>>>(a);
    else smooth_tile_kernel<-1><<<tblocks, kTileThreads, 0, ctx->stream>>>(a);
    DCB_LAUNCH_CHECK(ctx);
    // reads longer than the model's window (possible through the smoothing-only entry points) were listed, not processed:
    // the warp-per-read kernel, which takes any length, visits exactly those (normally none: it reads the count and exits)
    if (window == 21) smooth_chop_kernel<false, 10><<<ctx->sm_count, threads, 0, ctx->stream>>>(a);
    else smooth_chop_kernel<false, -1><<<ctx->sm_count, threads, 0, ctx->stream>>>(a);
    DCB_LAUNCH_CHECK(ctx);
    return DCB200_OK;
  }
  if (window == 21) {
    if (logits) smooth_chop_kernel<true, 10><<<blocks, threads, 0, ctx->stream>>>(a);
    else smooth_chop_kernel<false, 10><<<blocks, threads, 0, ctx->stream>>>(a);
  } else {
    if (logits) smooth_chop_kernel<true, -1><<<blocks, threads, 0, ctx->stream>>>(a);
    else smooth_chop_kernel<false, -1><<<blocks, threads, 0, ctx->stream>>>(a);
  }
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int smooth_chop_device(dcb200_ctx* ctx, const int8_t* labels, const float* logits, int64_t total, const int64_t* starts,
                       const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                       int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action,
                       int8_t* smoothed) {
  SmoothArgs a;
  a.labels = labels;
  a.logits = logits;
  a.total = total;
  a.starts = starts;
  a.lens = lens;
  a.qual_lens = qual_lens;
  a.R = R;
  a.p = *p;
  a.n_adapter = n_adapter;
  a.adapter_iv = adapter_iv;
  a.n_keep = n_keep;
  a.keep_iv = keep_iv;
  a.action = action;
  a.smoothed = smoothed;
  a.sel = nullptr;
  a.sel_count = nullptr;
  return launch_smooth(ctx, a);
}

}  // namespace dcb
