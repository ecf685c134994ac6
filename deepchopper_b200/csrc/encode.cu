// FASTQ bytes -> left-padded token ids + L2-normalised quality.
//
// Replaces the per-read Python/PyO3 chain  parse_fastq_file (deepchopper/data/only_fq.py:21-85) ->
// encode_qual (src/python.rs:25-35) -> normalize_seq (src/python.rs:272-275) ->
// tokenize_and_align_labels_and_quals_ids (deepchopper/models/llm/tokenizer.py:145-178) ->
// DataCollatorForTokenClassificationWithQual (tokenizer.py:34-93, LEFT pad 4 / 0.0).
//
// One warp per read.  Pass 1: exact integer sum of squares of (q-33) (so the L2 norm is the
// correctly rounded fp32 value of the exact sum).  Pass 2: each lane produces 4 consecutive output
// columns: one 32-bit token store and one float4 quality store, both row-aligned and coalesced.
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

__global__ void __launch_bounds__(256) encode_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ seq_off,
                                                     const int64_t* __restrict__ qual_off, const int32_t* __restrict__ len,
                                                     int32_t R, int32_t Lpad, int32_t Lrow,
                                                     uint8_t* __restrict__ tok, float* __restrict__ qual) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  for (int r = warp; r < R; r += nwarps) {
    const int n = len[r];
    const uint8_t* s = bytes + seq_off[r];
    const uint8_t* q = bytes + qual_off[r];
    // pass 1: ||q-33||^2, exact in 64-bit integers
    unsigned long long acc = 0;
    for (int i = lane; i < n; i += 32) {
      int v = (int)q[i] - 33;
      acc += (unsigned long long)(v * v);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    float nrm = sqrtf((float)acc);               // F.normalize: x / max(||x||_2, eps), tokenizer.py:167
    nrm = fmaxf(nrm, 1e-12f);
    const int pad = Lpad - (n + 1);
    uint32_t* trow = reinterpret_cast<uint32_t*>(tok + (int64_t)r * Lrow);
    float4* qrow = reinterpret_cast<float4*>(qual + (int64_t)r * Lrow);
    for (int c4 = lane; c4 < Lrow / 4; c4 += 32) {
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
  }
}

int encode_device(dcb200_ctx* ctx, const uint8_t* bytes, const int64_t* seq_off, const int64_t* qual_off,
                  const int32_t* len, int32_t R, int32_t Lpad, int32_t Lrow, uint8_t* tok, float* qual) {
  if (R == 0) return DCB200_OK;
  const int threads = 256;
  int blocks = (R + 7) / 8;
  const int cap = ctx->sm_count * 8 * 4;
  if (blocks > cap) blocks = cap;
  ProfScope prof(ctx, K_ENCODE);
  encode_kernel<<<blocks, threads, 0, ctx->stream>>>(bytes, seq_off, qual_off, len, R, Lpad, Lrow, tok, qual);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
