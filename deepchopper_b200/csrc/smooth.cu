// Segmented warp-level smoothing / interval / chop-coordinate kernel.
//
// Replaces, per read, the reference's CPU chain
//   majority_voting        src/smooth/utils.rs:48-97
//   get_label_region       src/utils.rs:671-695         (start==0 sentinel quirk kept: position 0 is masked)
//   smooth_and_select_intervals  src/smooth/predict.rs:186-209
//   generate_unmaped_intervals + _split_records_by_remove_internal + chop-type gate
//                          src/output/split.rs:60-136,171-201,260-292
//   process_chunk gating   src/bin/predict.rs:141-164
//
// One warp owns one read.  Labels are packed to a bit stream (32 bases per lane-word, 1024 bases per
// warp step); the majority vote over the default 21-wide window is a bit-sliced carry-save adder tree
// on those words (all 32 positions of a lane at once), run boundaries are found with shifts/ballots,
// and intervals are emitted in order with a warp prefix sum.  HBM traffic is the 1 byte per base of
// the labels (8 B/base when reading fp32 logits) plus <= a few dozen bytes of coordinates per read.
#include "common.cuh"

namespace dcb {

__device__ __forceinline__ uint32_t pack16(const uint4 v) {
  // 16 int8 labels -> 16 bits (bit i = labels[i] == 1)
  uint32_t r = 0;
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t e = __vcmpeq4(w[i], 0x01010101u) & 0x01010101u;
    r |= ((e * 0x01020408u) >> 24) << (4 * i);
  }
  return r;
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
      return pack16(v0) | (pack16(v1) << 16);
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
  return r.fast ? (pack16(r.v0) | (pack16(r.v1) << 16)) : r.slow;
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
  if (warp_global < a.R) {
    start_n = a.starts[warp_global];
    n_n = a.lens[warp_global];
    if (!LOGITS) {
      int mis0;
      int64_t ab0;
      const bool live = plan_read(start_n, n_n, mis0, ab0);
      pre = issue_word(a, ab0 + 32 * (int64_t)lane, live && lane <= n_n / 32 + 1);
    }
    if (warp_global + nwarps < a.R) {
      start_nn = a.starts[warp_global + nwarps];
      n_nn = a.lens[warp_global + nwarps];
    }
  }
  for (int64_t r = warp_global; r < a.R; r += nwarps) {
    const int64_t start = start_n;
    const int n = (int)n_n;
    const RawWord cur = pre;
    start_n = start_nn;
    n_n = n_nn;
    if (!LOGITS && r + nwarps < a.R) {
      int mis1;
      int64_t ab1;
      const bool live = plan_read(start_n, n_n, mis1, ab1);
      pre = issue_word(a, ab1 + 32 * (int64_t)lane, live && lane <= n_n / 32 + 1);
    }
    if (r + 2 * nwarps < a.R) {
      start_nn = a.starts[r + 2 * nwarps];
      n_nn = a.lens[r + 2 * nwarps];
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
    if (lane == 0) {
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
          const volatile int32_t* ad = a.adapter_iv + r * approved * 2;
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
          if (cur < (int)n - 1) {
            ++before;
            if ((int)n - 1 - cur >= mc) {
              keep[2 * nk] = cur;
              keep[2 * nk + 1] = (int)n - 1;
              if (nk == 0) first_len = (int)n - 1 - cur;
              ++nk;
            }
          }
          const bool terminal = before == 1;  // src/output/split.rs:185-189
          const int ct = a.p.chop_type;
          if ((ct == DCB200_CHOP_TERMINAL && !terminal) || (ct == DCB200_CHOP_INTERNAL && terminal) ||
              (nk > 0 && first_len == (int)n)) {
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
    __syncwarp();
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
  return launch_smooth(ctx, a);
}

}  // namespace dcb
