/* CPU ORACLE (test infrastructure, NOT the product): plain-C restatement of DeepChopper's
 * smooth / interval / chop-coordinate pass.  Each function cites the reference file:line it
 * follows (paths relative to /root/reference).  Built by oracle/Makefile into
 * oracle/_build/libdcref.so and used only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg.
 *
 * Parity: pinned by the reference's Rust unit-test vectors (tests/test_oracle_smooth.py) and
 * cross-checked against the pure-Python restatement oracle/smooth_ref.py.
 *
 * The algorithmic shape is kept literal (per-position window recount, O(n*W)), exactly what
 * src/smooth/utils.rs:62-95 does, minus the per-base HashMap/Vec allocations.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DCREF_ACTION_PASSTHROUGH 0
#define DCREF_ACTION_CHOP_T 1
#define DCREF_ACTION_CHOP_I 2
#define DCREF_ACTION_ADAPTERS 3
#define DCREF_ACTION_UNCHOPPED 4
#define DCREF_CHOP_TERMINAL 0
#define DCREF_CHOP_INTERNAL 1
#define DCREF_CHOP_ALL 2

/* src/smooth/utils.rs:48-97 (binary labels) */
void dcref_majority_voting(const int8_t* labels, int64_t n, int window, int8_t* out) {
  if (window % 2 == 0) window += 1;               /* :50-54 */
  int64_t half = window / 2;
  for (int64_t i = 0; i < n; ++i) {
    int64_t start = i - half < 0 ? 0 : i - half;   /* saturating_sub :65 */
    int64_t end = i + half + 1 < n ? i + half + 1 : n;
    if (end == n && (end - start) < window) {      /* :69-71 */
      start = end - window < 0 ? 0 : end - window;
    }
    int64_t c1 = 0, c0 = 0;
    for (int64_t j = start; j < end; ++j) {        /* literal recount :73-83 */
      if (labels[j] == 1) c1++; else c0++;
    }
    if (c1 == c0) out[i] = labels[i];               /* two classes, equal counts :86-91 */
    else out[i] = c1 > c0 ? 1 : 0;                  /* max_by_key :93-96 */
    /* note: labels other than 1 are counted as class 0; the reference's behaviour for >2 classes
       is hash-order dependent and never occurs (argmax over two classes). */
  }
}

/* src/utils.rs:671-695 incl. the start==0 sentinel quirk. returns number of regions (may exceed
 * cap; only the first cap are stored). */
int64_t dcref_get_label_region(const int8_t* labels, int64_t n, int64_t* out, int64_t cap) {
  int64_t cnt = 0, start = 0, end = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (labels[i] == 1) {
      if (start == 0) start = i;
      end = i;
    } else if (start != 0) {
      if (cnt < cap) { out[2 * cnt] = start; out[2 * cnt + 1] = end + 1; }
      cnt++;
      start = 0; end = 0;
    }
  }
  if (start != 0) {
    if (cnt < cap) { out[2 * cnt] = start; out[2 * cnt + 1] = end + 1; }
    cnt++;
  }
  return cnt;
}

/* one read: src/smooth/predict.rs:186-209 + src/bin/predict.rs:141-187 + src/output/split.rs:60-136,171-201,260-292 */
static void dcref_one(const int8_t* labels, int64_t n, int64_t qual_len, int window, int min_interval,
                      int approved, int max_process, int min_after_chop, int min_read_len, int chop_type,
                      int ocq, int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv,
                      uint8_t* action, int8_t* scratch) {
  *n_adapter = 0; *n_keep = 0; *action = DCREF_ACTION_PASSTHROUGH;
  if (n < min_read_len) return;                                 /* bin/predict.rs:146-148 */
  dcref_majority_voting(labels, n, window, scratch);
  /* regions, filtered by min length (smooth/predict.rs:194-203) */
  int64_t cnt = 0, start = 0, end = 0;
  for (int64_t i = 0; i <= n; ++i) {
    int is1 = (i < n) && scratch[i] == 1;
    if (is1) {
      if (start == 0) start = i;
      end = i;
    } else if (start != 0) {
      if (end + 1 - start >= min_interval) {
        if (cnt < approved) { adapter_iv[2 * cnt] = (int32_t)start; adapter_iv[2 * cnt + 1] = (int32_t)(end + 1); }
        cnt++;
      }
      start = 0; end = 0;
    }
  }
  if (cnt > approved) cnt = 0;                                  /* smooth/predict.rs:204-206 */
  *n_adapter = (int32_t)cnt;
  if (cnt > max_process || cnt == 0) return;                    /* bin/predict.rs:156-158 */
  if (qual_len != n) return;                                    /* bin/predict.rs:160-164 */
  if (ocq) { *action = DCREF_ACTION_ADAPTERS; return; }         /* split.rs:138-169 */
  /* generate_unmaped_intervals split.rs:260-292 (intervals already sorted) */
  int64_t before = 0, nk = 0, cur = 0, first_len = -1;
  for (int64_t k = 0; k < cnt; ++k) {
    int64_t s = adapter_iv[2 * k], e = adapter_iv[2 * k + 1];
    if (cur < s) {
      before++;
      if (s - cur >= min_after_chop) { keep_iv[2 * nk] = (int32_t)cur; keep_iv[2 * nk + 1] = (int32_t)s; if (nk == 0) first_len = s - cur; nk++; }
    }
    cur = e;
  }
  if (cur < n - 1) {
    before++;
    if (n - 1 - cur >= min_after_chop) { keep_iv[2 * nk] = (int32_t)cur; keep_iv[2 * nk + 1] = (int32_t)(n - 1); if (nk == 0) first_len = n - 1 - cur; nk++; }
  }
  int terminal = before == 1;                                   /* split.rs:185-189 */
  if ((chop_type == DCREF_CHOP_TERMINAL && !terminal) || (chop_type == DCREF_CHOP_INTERNAL && terminal) ||
      (nk > 0 && first_len == n)) {                             /* split.rs:191-201 */
    *action = DCREF_ACTION_UNCHOPPED;
    return;
  }
  *n_keep = (int32_t)nk;
  *action = terminal ? DCREF_ACTION_CHOP_T : DCREF_ACTION_CHOP_I;
}

/* Batched driver with the same argument meaning / output layout as include/dcb200.h:dcb200_smooth_chop. */
int dcref_smooth_chop(const int8_t* labels, const int64_t* starts, const int32_t* lens, const int32_t* qual_lens,
                      int64_t R, int window, int min_interval, int approved, int max_process, int min_after_chop,
                      int min_read_len, int chop_type, int ocq, int32_t* n_adapter, int32_t* adapter_iv,
                      int32_t* n_keep, int32_t* keep_iv, uint8_t* action, int threads) {
  int64_t maxlen = 0;
  for (int64_t r = 0; r < R; ++r) if (lens[r] > maxlen) maxlen = lens[r];
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    int8_t* scratch = (int8_t*)malloc((size_t)maxlen + 1);
#pragma omp for schedule(dynamic, 64)
    for (int64_t r = 0; r < R; ++r) {
      int64_t n = lens[r];
      dcref_one(labels + starts[r], n, qual_lens ? qual_lens[r] : n, window, min_interval, approved, max_process,
                min_after_chop, min_read_len, chop_type, ocq, n_adapter + r, adapter_iv + 2 * (int64_t)approved * r,
                n_keep + r, keep_iv + 2 * (int64_t)(approved + 1) * r, action + r, scratch);
    }
    free(scratch);
  }
  return 0;
}

/* FASTQ -> token / L2-normalised quality encoding of one read into a left-padded row.
 * deepchopper/models/llm/tokenizer.py:145-178 (ids + [SEP]=1, quals = normalize(cat(q,[0])), eps 1e-12)
 * and :64-84 (LEFT pad: ids 4, quals 0).  seq chars: A7 C8 G9 T10 N11 (lower-case folded, U->T per
 * needletail normalize, src/python.rs:272-275), anything else N(11). */
void dcref_encode_read(const uint8_t* seq, const uint8_t* qual, int32_t len, int32_t Lpad, uint8_t* tok, float* q) {
  int32_t pad = Lpad - (len + 1);
  for (int32_t i = 0; i < pad; ++i) { tok[i] = 4; q[i] = 0.0f; }
  float ss = 0.0f;
  /* torch's vector_norm on a contiguous float tensor accumulates in float with a pairwise/vectorised
     order; we accumulate in double and round once -- differences are below 1 ulp of the norm and
     are covered by the stated tolerance in tests (1e-6 relative). */
  double acc = 0.0;
  for (int32_t i = 0; i < len; ++i) { double v = (double)((int)qual[i] - 33); acc += v * v; }
  ss = (float)acc;
  float nrm = __builtin_sqrtf(ss);
  if (nrm < 1e-12f) nrm = 1e-12f;
  for (int32_t i = 0; i < len; ++i) {
    uint8_t c = seq[i];
    if (c >= 'a' && c <= 'z') c -= 32;              /* pyfastx uppercase=True, only_fq.py:34 */
    uint8_t t;
    switch (c) {
      case 'A': t = 7; break; case 'C': t = 8; break; case 'G': t = 9; break;
      case 'T': case 'U': t = 10; break;            /* needletail normalize: U -> T */
      case 'N': t = 11; break;
      case '-': case '.': case '~': t = 6; break;   /* normalize -> '-', not in vocab -> [UNK] */
      default: t = 11;                              /* other -> N */
    }
    tok[pad + i] = t;
    q[pad + i] = (float)((int)qual[i] - 33) / nrm;
  }
  tok[pad + len] = 1; q[pad + len] = 0.0f;
}
