/* dcb200.h -- C ABI of the B200-native DeepChopper predict hot path.
 *
 * One shared library (libdcb200.so, sm_100a only) behind plain pointers and sizes: no torch, no
 * C++ types.  Every entry point returns 0 on success or a negative DCB200_E* code;
 * dcb200_last_error() returns a thread-local, human-readable message for the last failure.
 * A dcb200_ctx owns one device + one stream (and its workspaces); it is not thread-safe, use
 * one ctx per thread/GPU.  The caller owns every buffer passed in.
 *
 * "device" pointers must be device-accessible (cudaMalloc or pinned/UVA host memory);
 * "host" pointers are ordinary host memory (the *_host entry points copy inside the call).
 *
 * Reference interfaces replaced (paths relative to the reference checkout):
 *   dcb200_encode_batch     deepchopper/data/only_fq.py:21-85 (parse_fastq_file), src/python.rs:25-35 (encode_qual),
 *                           src/python.rs:272-275 (normalize_seq), deepchopper/models/llm/tokenizer.py:145-178
 *                           (tokenize_and_align_labels_and_quals_ids) and :34-93 (collator, LEFT pad)
 *   dcb200_weights_create   deepchopper/models/dc_hg.py:70-163 (from_checkpoint / from_pretrained state dict)
 *   dcb200_forward          deepchopper/models/basic_module.py:90-100,197-207 -> deepchopper/models/llm/hyena.py:29-41
 *                           (HyenaDNA backbone, HF remote code) -> deepchopper/models/llm/head.py:94-102
 *   dcb200_majority_voting  src/smooth/utils.rs:48-97  (PyO3: src/python.rs:815-818)
 *   dcb200_smooth_chop      src/smooth/predict.rs:186-209 (smooth_and_select_intervals) == src/utils.rs:699-721
 *                           (smooth_label_region, PyO3 src/python.rs:674-690), src/utils.rs:671-695 (get_label_region),
 *                           src/output/split.rs:60-136,171-226,260-320 and the gating of src/bin/predict.rs:137-187
 *   dcb200_smooth_chop_logits  the same after src/smooth/predict.rs:263-317 (argmax(2) + drop target == -100)
 *   dcb200_chop_write_bgzf  src/bin/predict.rs:266-364 (write loop), src/output/split.rs:60-226, src/output/writefq.rs
 *   dcb200_read_file_inflate  src/output/writefq.rs:84-193 (plain / gzip / bgzip readers), deepchopper/data/only_fq.py:21-85
 *   dcb200_predict_batch_host  the whole per-batch hot loop of `deepchopper predict` + the interval step of
 *                           `deepchopper chop` (cli.py:66-152, src/bin/predict.rs:130-192) on host buffers
 */
#ifndef DCB200_H_
#define DCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCB200_OK 0
#define DCB200_EINVAL (-1)   /* bad argument */
#define DCB200_ECUDA (-2)    /* CUDA runtime / driver error (message has the CUDA string) */
#define DCB200_ENOMEM (-3)
#define DCB200_ENODEV (-4)   /* no sm_100 device: this library has no CPU fallback */
#define DCB200_EWEIGHT (-5)  /* missing / mis-shaped tensor in the state dict */

/* per-read action of the chop step (src/bin/predict.rs:141-187) */
#define DCB200_ACTION_PASSTHROUGH 0 /* emit the FASTQ record verbatim */
#define DCB200_ACTION_CHOP_T 1      /* emit keep_iv pieces named "{id}|s:e|T" */
#define DCB200_ACTION_CHOP_I 2      /* emit keep_iv pieces named "{id}|s:e|I" */
#define DCB200_ACTION_ADAPTERS 3    /* --ocq: emit adapter_iv pieces named "{id}|s:e" */
#define DCB200_ACTION_UNCHOPPED 4   /* chop-type mismatch (src/output/split.rs:191-201): emit "@{id}" (no description)
                                       with the prediction-decoded sequence and the FASTQ quality, uncut */

#define DCB200_CHOP_TERMINAL 0
#define DCB200_CHOP_INTERNAL 1
#define DCB200_CHOP_ALL 2

/* token ids of the HyenaDNA character tokenizer (src/smooth/utils.rs:6-25) */
#define DCB200_TOK_SEP 1
#define DCB200_TOK_PAD 4
#define DCB200_TOK_UNK 6
#define DCB200_TOK_A 7
#define DCB200_TOK_C 8
#define DCB200_TOK_G 9
#define DCB200_TOK_T 10
#define DCB200_TOK_N 11

typedef struct dcb200_ctx dcb200_ctx;
typedef struct dcb200_weights dcb200_weights;

/* clap defaults of deepchopper-chop (src/bin/predict.rs:19-78) + src/default.rs */
typedef struct dcb200_chop_params {
  int32_t smooth_window_size;         /* 21 */
  int32_t min_interval_size;          /* 13 */
  int32_t approved_interval_number;   /* 20 */
  int32_t max_process_intervals;      /* 4 */
  int32_t min_read_length_after_chop; /* 20 */
  int32_t min_read_length;            /* 150 = MIN_READ_LEN; 0 disables the gate (PyO3 smooth_label_region) */
  int32_t chop_type;                  /* DCB200_CHOP_* */
  int32_t output_chopped_seqs;        /* --ocq */
} dcb200_chop_params;

const char* dcb200_last_error(void);
int dcb200_version(void);
void dcb200_chop_params_default(dcb200_chop_params* p);

/* ctx: stream == NULL -> the ctx creates (and owns) a non-blocking stream; otherwise the given
 * cudaStream_t is used (e.g. torch.cuda.current_stream().cuda_stream). */
int dcb200_ctx_create(int device, void* stream, dcb200_ctx** out);
int dcb200_ctx_destroy(dcb200_ctx* ctx);
int dcb200_ctx_sync(dcb200_ctx* ctx);
void* dcb200_ctx_stream(dcb200_ctx* ctx);
/* number of kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t dcb200_ctx_launch_count(dcb200_ctx* ctx);
/* ctx options.  "fft_min_len": batches whose padded length is at least this take the blocked shared-memory FFT
 * long convolution (the reference's fftconv, SURVEY Appendix A / deepchopper/models/llm/hyena.py:34-41), shorter
 * ones the tensor-core Toeplitz kernel.  The default value (6144) is special: it selects the measured cost model,
 * which looks at the batch's rows and length (FFT from ~6.7 k tokens with full 128-row tiles, from ~3.4 k tokens
 * for a 16-row batch).  0 = always FFT, a huge value = never.
 * "smooth_warp_kernel": which kernel dcb200_smooth_chop takes for int8 labels: 0 (default) = the tile kernel (thread per
 * 32-base word) when the launch holds at least 64 reads per SM, else the warp-per-read kernel (the one the logits and
 * smoothed-label forms always use); 1 = always warp-per-read; 2 = always tile.  Both are bit-exact, tests compare them.
 * dcb200_ctx_get_option returns -1 for an unknown name. */
int dcb200_ctx_set_option(dcb200_ctx* ctx, const char* name, int64_t value);
int64_t dcb200_ctx_get_option(dcb200_ctx* ctx, const char* name);

/* ---- FASTQ -> tokens + L2-normalised quality --------------------------------------------------
 * bytes: device-accessible FASTQ text (or any byte buffer holding seq and quality strings);
 * seq_off/qual_off[R]: byte offsets of each read's sequence / quality string, len[R]: bases kept
 * (already truncated to max_tokens-1 by the caller, tokenizer.py:154-163); Lpad >= max(len)+1 is the
 * collated batch length (the reference pads to the batch maximum); Lrow >= Lpad, Lrow % 4 == 0, is the
 * row stride of tok/qual.  Row r = [PAD(4) x (Lpad-len-1)][bases][SEP(1)][PAD x (Lrow-Lpad)], qual 0
 * at pads and SEP.  (Device form: the caller guarantees seq_off[r] + len[r] and qual_off[r] + len[r] lie inside `bytes`;
 * nothing outside [string, string + len) is read.  The host form dcb200_predict_batch_host checks every offset against
 * n_bytes.)  The left pads are what the reference feeds the model; the right filler only
 * rounds the row up to the kernels' tile size -- the model is causal, so it cannot influence
 * columns < Lpad. */
int dcb200_encode_batch(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                        const int32_t* len, int32_t R, int32_t Lpad, int32_t Lrow, uint8_t* tok, float* qual);
/* Several of the reference's batches in ONE launch: row r is collated as in ITS batch, i.e. left-padded to
 * lpad_rows[r] (device array; len[r] + 1 <= lpad_rows[r] <= Lpad), and right-filled to Lrow like every row.  The reference
 * pads a FASTQ-order batch of 12-16 reads to that batch's maximum (deepchopper/data/only_fq.py:198-202,
 * deepchopper/models/llm/tokenizer.py:34-93) and the pads are semantic, so a row must keep its own batch's pad count; the
 * model is causal and rows are independent, so rows of many such batches can share a launch (full 128-row tiles instead
 * of 16 rows) without changing what any of them sees in columns < lpad_rows[r]. */
int dcb200_encode_batch_rows(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                             const int32_t* len, const int32_t* lpad_rows, int32_t R, int32_t Lpad, int32_t Lrow,
                             uint8_t* tok, float* qual);

/* ---- weights ------------------------------------------------------------------------------------
 * State dict as parallel arrays of names / host fp32 pointers / element counts.  Names are the
 * reference's keys ("net.backbone.backbone.layers.0.mixer.in_linear.weight", ...; any prefix before
 * "embeddings." / "layers." / "ln_f." / "head." is ignored).  Converts to the device layouts and
 * evaluates the implicit long-conv filters k[layer][256][max_seq_len] once. */
int dcb200_weights_create(dcb200_ctx* ctx, const char* const* names, const float* const* data, const int64_t* numel,
                          int32_t n_tensors, dcb200_weights** out);
int dcb200_weights_destroy(dcb200_weights* w);

/* ---- model forward ------------------------------------------------------------------------------
 * tok [B,L] u8, qual [B,L] f32 (device).  L must be a multiple of 128 and <= 32768.
 * logits [B,L,2] f32 and/or labels [B,L] u8 (label = logit1 > logit0) may be NULL. */
int dcb200_forward(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* tok, const float* qual, int32_t B,
                   int32_t L, float* logits, uint8_t* labels);

/* ---- smoothing / intervals / chop coordinates ------------------------------------------------------
 * labels: device int8 buffer of labels_bytes bytes; read r occupies labels[starts[r] .. +lens[r]).
 * qual_lens (device, may be NULL): FASTQ quality length per read; != lens[r] means the prediction
 * was truncated -> passthrough (src/bin/predict.rs:160-164).
 * Outputs (device): n_adapter[R], adapter_iv[R][approved][2], n_keep[R], keep_iv[R][approved+1][2],
 * action[R].  Only the first n_* entries of a row are written. */
int dcb200_smooth_chop(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                       const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                       int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action);

/* Same, reading fp32 logits [n_tokens][2] instead of labels: label = logit1 > logit0 (argmax, ties
 * -> class 0); starts/lens index tokens. */
int dcb200_smooth_chop_logits(dcb200_ctx* ctx, const float* logits, int64_t n_tokens, const int64_t* starts,
                              const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                              int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv,
                              uint8_t* action);

/* majority_voting over R reads; out has the layout of labels (only read positions are written).  Label domain: a
 * position counts as 1 iff its byte equals 1, anything else as 0 (the reference votes over arbitrary label values; on
 * this path labels are the argmax of two classes). */
int dcb200_majority_voting(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                           const int32_t* lens, int64_t R, int32_t window, int8_t* out);

/* Host-buffer convenience forms: copy in, run, copy out, synchronise. */
int dcb200_smooth_chop_host(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                            const int32_t* lens, const int32_t* qual_lens, int64_t R, const dcb200_chop_params* p,
                            int32_t* n_adapter, int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv,
                            uint8_t* action);
int dcb200_majority_voting_host(dcb200_ctx* ctx, const int8_t* labels, int64_t labels_bytes, const int64_t* starts,
                                const int32_t* lens, int64_t R, int32_t window, int8_t* out);

/* ---- whole hot loop on host buffers ----------------------------------------------------------------
 * One batch of R reads out of a (pinned) host FASTQ buffer: H2D -> encode -> forward -> smooth/chop
 * -> D2H, synchronised on return.  Lpad is the collated batch length (any value >= max(len)+1; rows
 * are rounded up to a multiple of 128 internally).  logits_out [R,Lpad,2] / labels_out [R,Lpad] (host) may be NULL;
 * interval outputs (host) as in dcb200_smooth_chop.  qual_lens may be NULL. */
int dcb200_predict_batch_host(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* bytes, int64_t n_bytes,
                              const int64_t* seq_off, const int64_t* qual_off, const int32_t* len,
                              const int32_t* qual_lens, int32_t R, int32_t Lpad, const dcb200_chop_params* p,
                              float* logits_out, uint8_t* labels_out, int32_t* n_adapter, int32_t* adapter_iv,
                              int32_t* n_keep, int32_t* keep_iv, uint8_t* action);
/* The same with several of the reference's FASTQ-order batches in one call (see dcb200_encode_batch_rows): row r is
 * left-padded to lpad_rows[r] (host array, len[r] + 1 <= lpad_rows[r] <= Lpad), the collated length of ITS batch.
 * logits_out / labels_out keep the row width Lpad; row r's read sits in columns [lpad_rows[r] - 1 - len[r], lpad_rows[r] - 1),
 * its SEP in column lpad_rows[r] - 1, columns beyond are filler. */
int dcb200_predict_batch_host_rows(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* bytes, int64_t n_bytes,
                                   const int64_t* seq_off, const int64_t* qual_off, const int32_t* len,
                                   const int32_t* lpad_rows, const int32_t* qual_lens, int32_t R, int32_t Lpad,
                                   const dcb200_chop_params* p, float* logits_out, uint8_t* labels_out, int32_t* n_adapter,
                                   int32_t* adapter_iv, int32_t* n_keep, int32_t* keep_iv, uint8_t* action);

/* ---- chop output: record assembly + BGZF on host threads ---------------------------------------------
 * Replaces the write loop of `deepchopper-chop` (src/bin/predict.rs:266-364), the record naming / slicing of
 * src/output/split.rs:60-226 and the bgzf writer of src/output/writefq.rs.  No GPU work: the per-read decisions are the
 * outputs of dcb200_smooth_chop* (action, adapter_iv [R,adapter_stride,2], keep_iv [R,keep_stride,2]).
 * Records are visited in FASTQ order; has_pred[r] == 0 drops the record (no prediction for it).  pseq[r] / pseq_len[r]
 * is the sequence decoded from the prediction batch (src/smooth/predict.rs:301: tokens -> ACGTN), used for every
 * action but PASSTHROUGH, which copies the FASTQ record verbatim (header with description, sequence, quality).
 * threads <= 0: all host cores; level 1-9 = zlib levels, 0 = Huffman-only deflate (about 8x the speed of level 6 for
 * ~5 % more bytes on FASTQ text), anything else = 6.  The file is a valid BGZF (htslib) stream whose decompressed
 * bytes do not depend on `threads`. */
typedef struct dcb200_fastq_index {
  const uint8_t* fastq;      /* the whole FASTQ text */
  const int64_t* name_off;   /* [R] offset of the byte after '@' */
  const int32_t* name_len;   /* [R] id length (up to the first blank) */
  const int32_t* head_len;   /* [R] full header line length after '@' */
  const int64_t* seq_off;    /* [R] */
  const int32_t* seq_len;    /* [R] */
  const int64_t* qual_off;   /* [R] */
  const int32_t* qual_len;   /* [R] */
} dcb200_fastq_index;
int dcb200_chop_write_bgzf(const dcb200_fastq_index* ix, int64_t R, const uint8_t* has_pred,
                           const uint8_t* const* pseq, const int32_t* pseq_len, const uint8_t* action,
                           const int32_t* n_adapter, const int32_t* adapter_iv, int32_t adapter_stride,
                           const int32_t* n_keep, const int32_t* keep_iv, int32_t keep_stride, const char* path,
                           int32_t threads, int32_t level, int64_t* n_records, int64_t* n_text_bytes);

/* The same for one CHUNK of records of a FASTQ that is streamed (src/bin/predict.rs:275-316 walks the FASTQ in chunks of
 * 10 000 records): DCB200_WRITE_APPEND appends to `path` instead of truncating it, DCB200_WRITE_NO_EOF leaves out the
 * 28-byte BGZF end-of-file block (every call but the last).  BGZF members concatenate, so the parts form one valid file
 * whose decompressed bytes equal the one-call output.  R = 0 with neither flag writes an empty BGZF file. */
#define DCB200_WRITE_APPEND 1
#define DCB200_WRITE_NO_EOF 2
int dcb200_chop_write_bgzf_part(const dcb200_fastq_index* ix, int64_t R, const uint8_t* has_pred,
                                const uint8_t* const* pseq, const int32_t* pseq_len, const uint8_t* action,
                                const int32_t* n_adapter, const int32_t* adapter_iv, int32_t adapter_stride,
                                const int32_t* n_keep, const int32_t* keep_iv, int32_t keep_stride, const char* path,
                                int32_t threads, int32_t level, int32_t flags, int64_t* n_records, int64_t* n_text_bytes);

/* ---- FASTQ ingest, file level: plain / gzip / BGZF file -> one host buffer ----------------------------
 * Replaces the compression sniffing and single-threaded readers of src/output/writefq.rs:84-193 and the file access
 * of deepchopper/data/only_fq.py:21-85 (pyfastx).  BGZF blocks (bgzip, and what dcb200_chop_write_bgzf writes) are
 * inflated on `threads` host threads (<= 0: all cores) with their CRCs checked; a plain gzip stream by one thread; any
 * other file is returned as is.  *out is malloc'ed and released with dcb200_free; *kind (may be NULL) = 0 plain,
 * 1 gzip, 2 BGZF. */
int dcb200_read_file_inflate(const char* path, int32_t threads, uint8_t** out, int64_t* out_len, int32_t* kind);
void dcb200_free(void* p);

/* Record index of a FASTQ text held in memory (4-line records), built on `threads` host threads (<= 0: all cores):
 * replaces the per-record Python loop of deepchopper/data/only_fq.py:21-85 (pyfastx iteration + validation: '@' / '+'
 * lines, equal and non-zero sequence / quality lengths) and the record splitting of noodles in src/output/writefq.rs.
 * The id (name) is the header up to the first blank; '\r' before '\n' is stripped; trailing empty lines are ignored.
 * The seven arrays are malloc'ed ([n_records] each) and released one by one with dcb200_free; on error nothing is
 * allocated and dcb200_last_error() names the first offending record.  The result feeds dcb200_encode_batch /
 * dcb200_predict_batch_host (seq_off, qual_off, len) and dcb200_chop_write_bgzf (dcb200_fastq_index). */
typedef struct dcb200_fastq_index_arrays {
  int64_t n_records;
  int64_t* name_off;
  int32_t* name_len;
  int32_t* head_len;
  int64_t* seq_off;
  int32_t* seq_len;
  int64_t* qual_off;
  int32_t* qual_len;
} dcb200_fastq_index_arrays;
int dcb200_index_fastq(const uint8_t* fastq, int64_t n_bytes, int32_t threads, dcb200_fastq_index_arrays* out);

/* ---- diagnostics (used by the parity tests to localise a mismatch) ---------------------------------
 * dcb200_forward_debug == dcb200_forward that stops after `stop_stage` kernels of the forward chain
 * (0 embed+LN1, then per layer l: 1+5l in_proj, 2+5l conv, 3+5l out_proj, 4+5l fc1, 5+5l fc2; 21 head1,
 * 22 head2 = full).  dcb200_ctx_read_workspace copies a named internal activation buffer to the host:
 * "act_hA" fp32 [T,256] residual stream, "act_u" bf16 [T,256], "act_vv" / "act_gate" / "act_y" bf16 [B,256,L],
 * "act_g" bf16 [T,1024].  Synchronises the stream. */
int dcb200_forward_debug(dcb200_ctx* ctx, const dcb200_weights* w, const uint8_t* tok, const float* qual, int32_t B,
                         int32_t L, float* logits, uint8_t* labels, int32_t stop_stage);
int dcb200_ctx_read_workspace(dcb200_ctx* ctx, const char* name, void* host_dst, int64_t bytes);

/* Per-kernel device timing with CUDA events on the ctx stream (bench.py's roofline figures).
 * dcb200_ctx_profile(ctx, 1) starts bracketing every kernel launch with an event pair;
 * dcb200_ctx_profile_read sums elapsed milliseconds and launch counts per kernel kind
 * (index -> dcb200_kernel_kind_name) into ms[n] / counts[n], optionally resetting the sums. */
int dcb200_ctx_profile(dcb200_ctx* ctx, int enable);
int dcb200_ctx_profile_read(dcb200_ctx* ctx, double* ms, int64_t* counts, int32_t n, int32_t reset);
const char* dcb200_kernel_kind_name(int32_t kind);

#ifdef __cplusplus
}
#endif
#endif /* DCB200_H_ */
