// FASTQ bytes -> left-padded token ids + L2-normalised quality.
//
// Replaces the per-read Python/PyO3 chain  parse_fastq_file (deepchopper/data/only_fq.py:21-85) ->
// encode_qual (src/python.rs:25-35) -> normalize_seq (src/python.rs:272-275) ->
// tokenize_and_align_labels_and_quals_ids (deepchopper/models/llm/tokenizer.py:145-178) ->
// DataCollatorForTokenClassificationWithQual (tokenizer.py:34-93, LEFT pad 4 / 0.0).
//
// One CTA per read (grid-stride).  The read's sequence and quality strings are staged in shared memory with coalesced,
// 16-byte aligned vector loads (the strings sit at arbitrary byte offsets of the FASTQ buffer: the copy starts at the
// aligned address below each string).  Pass 1: exact integer sum of squares of (q-33), four bytes per shared-memory
// word (so the L2 norm is the correctly rounded fp32 value of the exact sum).  Pass 2: each thread produces 4
// consecutive output columns: one 32-bit token store and one float4 quality store, both row-aligned and coalesced.
#include "common.cuh"

namespace dcb {

__device__ __forceinline__ uint32_t base_to_token(uint32_t c) {
  if (c >= 'a' && c <= 'z') c -= 32;  // pyfastx uppercase=True (only_fq.py:34)
  switch (c) {
    case 'A': return DCB200_TOK_A;
    case 'C': return DCB200_TOK_C;
    case 'G': return DCB200_TOK_G;
    case 'T':
    case 'U': return DCB200_TOK_T;    // needletail normalize: U -> T
    case 'N': return DCB200_TOK_N;
    case '-':
    case '.':
    case '~': return DCB200_TOK_UNK;  // normalize -> '-', which is not in the vocabulary
    default: return DCB200_TOK_N;     // any other byte -> N
  }
}

constexpr int kEncThreads = 256;

__global__ void __launch_bounds__(kEncThreads) encode_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ seq_off,
                                                             const int64_t* __restrict__ qual_off, const int32_t* __restrict__ len,
                                                             const int32_t* __restrict__ lpad_rows,
                                                             int32_t R, int32_t Lpad, int32_t Lrow, int32_t cap16,
                                                             uint8_t* __restrict__ tok, float* __restrict__ qual) {
  extern __shared__ uint4 enc_smem[];   // [2][cap16] 16-byte words: sequence, quality
  __shared__ unsigned long long red[kEncThreads / 32];
  __shared__ float nrm_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    // (contract: 0 <= len[r] <= Lpad - 1, include/dcb200.h; the table lives in device memory, so it is clamped here rather
    //  than trusted: the staging buffers are sized for Lpad - 1 bytes)
    // per-row collated length (several of the reference's batches in one launch, dcb200_encode_batch_rows), else the launch's
    const int lp = lpad_rows ? max(1, min(lpad_rows[r], Lpad)) : Lpad;
    const int n = max(0, min(len[r], lp - 1));
    const uintptr_t sa = reinterpret_cast<uintptr_t>(bytes + seq_off[r]);
    const uintptr_t qa = reinterpret_cast<uintptr_t>(bytes + qual_off[r]);
    const int smis = (int)(sa & 15), qmis = (int)(qa & 15);
    const uint4* sg = reinterpret_cast<const uint4*>(sa - smis);
    const uint4* qg = reinterpret_cast<const uint4*>(qa - qmis);
    const int ns16 = (n + smis + 15) >> 4, nq16 = (n + qmis + 15) >> 4;
    // interior 16-byte words with vector loads; the first and last word of a string byte by byte, so that nothing
    // outside [string, string + n) is ever read (the strings may touch the ends of the caller's buffer)
    auto stage = [&](const uint4* g, int mis, int n16, uint4* dst) {
      for (int i = tid; i < n16; i += kEncThreads) {
        if (i > 0 && i < n16 - 1) {
          dst[i] = __ldg(g + i);
        } else {
          const uint8_t* gb = reinterpret_cast<const uint8_t*>(g + i);
          uint8_t* db = reinterpret_cast<uint8_t*>(dst + i);
          for (int b = 0; b < 16; ++b) {
            const int pos = 16 * i + b - mis;  // index into the string
            db[b] = (pos >= 0 && pos < n) ? __ldg(gb + b) : (uint8_t)0;
          }
        }
      }
    };
    stage(sg, smis, ns16, enc_smem);
    stage(qg, qmis, nq16, enc_smem + cap16);
    __syncthreads();
    const uint8_t* s = reinterpret_cast<const uint8_t*>(enc_smem) + smis;
    const uint8_t* q = reinterpret_cast<const uint8_t*>(enc_smem + cap16) + qmis;
    // pass 1: ||q-33||^2, exact in 64-bit integers
    unsigned long long acc = 0;
    for (int i = tid; i < n; i += kEncThreads) {
      const int v = (int)q[i] - 33;
      acc += (unsigned long long)(v * v);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      unsigned long long t = 0;
      for (int w = 0; w < kEncThreads / 32; ++w) t += red[w];
      nrm_s = fmaxf(sqrtf((float)t), 1e-12f);    // F.normalize: x / max(||x||_2, eps), tokenizer.py:167
    }
    __syncthreads();
    const float nrm = nrm_s;
    const int pad = lp - (n + 1);
    uint32_t* trow = reinterpret_cast<uint32_t*>(tok + (int64_t)r * Lrow);
    float4* qrow = reinterpret_cast<float4*>(qual + (int64_t)r * Lrow);
    for (int c4 = tid; c4 < Lrow / 4; c4 += kEncThreads) {
      uint32_t tw = 0;
      float qv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 4 * c4 + j;
        const int i = col - pad;
        uint32_t t;
        float f = 0.0f;
        if (i < 0 || i > n) t = DCB200_TOK_PAD;  // left pad (semantic) / right filler up to the row stride (causal model: inert)
        else if (i == n) t = DCB200_TOK_SEP;
        else {
          t = base_to_token(s[i]);
          f = __fdiv_rn((float)((int)q[i] - 33), nrm);
        }
        tw |= t << (8 * j);
        qv[j] = f;
      }
      trow[c4] = tw;
      qrow[c4] = make_float4(qv[0], qv[1], qv[2], qv[3]);
    }
    __syncthreads();   // the staging buffers are reused by the next read
  }
}

int encode_device(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                  const int32_t* len, const int32_t* lpad_rows, int32_t R, int32_t Lpad, int32_t Lrow, uint8_t* tok,
                  float* qual) {
  if (R == 0) return DCB200_OK;
  // staging: two strings of at most Lpad - 1 bytes, each with up to 15 bytes of alignment slack on either side
  const int cap16 = (Lpad + 15 + 15) / 16 + 1;
  const size_t smem = (size_t)2 * cap16 * sizeof(uint4);
  DCB_CHECK(ctx->ensure_smem(reinterpret_cast<const void*>(&encode_kernel), smem > 48 * 1024 ? smem : 48 * 1024));
  int blocks = R;
  const int cap = ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ProfScope prof(ctx, K_ENCODE);
  encode_kernel<<<blocks, kEncThreads, smem, ctx->stream>>>(bytes, seq_off, qual_off, len, lpad_rows, R, Lpad, Lrow, cap16, tok,
                                                            qual);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
