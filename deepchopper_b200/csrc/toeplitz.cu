// Tensor-core form of the Hyena long convolution for L <= 4096:
//   y[b,c,t] = x0c[b,c,t] * sum_{s<=t} vv[b,c,s] * k'_c[t-s],   vv = sc(v) * sc(x1),  x0c = sc(x0),  k'[0] = k[0] + D
// The causal convolution of one channel is a lower-triangular Toeplitz matrix; cut into 128x128 blocks
// it is nb(nb+1)/2 dense bf16 GEMM tiles per 128 batch rows, and block (j,i) only depends on d = j-i.
// The blocks T_d[t',s'] = k'[128 d + t' - s'] are materialised once per weight set (bf16, K-major) and
// streamed by TMA as the B operand of gemm_kernel<G_TOEP>; the A operand is vv read channel-major.
// At L = 1-2k this is 10-20x less time than the fp32 shared-memory FFT (which stays for longer reads).
#include "common.cuh"
#include "toeplitz.h"

namespace dcb {

__global__ void __launch_bounds__(256) toeplitz_build_kernel(const float* __restrict__ k, int k_stride, int k_len,
                                                             const float* __restrict__ D, int nb_max,
                                                             __nv_bfloat16* __restrict__ T) {
  const int d = blockIdx.x, c = blockIdx.y;
  __nv_bfloat16* dst = T + ((size_t)c * nb_max + d) * 128 * 128;
  const float* kc = k + (size_t)c * k_stride;
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x) {
    const int tp = idx >> 7, sp = idx & 127;
    const int u = 128 * d + tp - sp;
    float v = 0.f;
    if (u >= 0 && u < k_len) v = kc[u];
    if (u == 0) v += D[c];
    dst[idx] = __float2bfloat16_rn(v);
  }
}

__device__ __forceinline__ void unpack8b(const uint4 v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// causal 3-tap depthwise conv of 8 consecutive tokens: out[t] = w0 z[t-2] + w1 z[t-1] + w2 z[t] + b
__device__ __forceinline__ void sconv8(const __nv_bfloat16* __restrict__ row, int t0, const float* __restrict__ w,
                                       float bias, float (&out)[8]) {
  float z[10], cur[8];
  unpack8b(__ldg(reinterpret_cast<const uint4*>(row + t0)), cur);
  z[0] = t0 >= 2 ? __bfloat162float(row[t0 - 2]) : 0.f;
  z[1] = t0 >= 1 ? __bfloat162float(row[t0 - 1]) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) z[2 + i] = cur[i];
  const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = fmaf(w0, z[i], fmaf(w1, z[i + 1], fmaf(w2, z[i + 2], bias)));
}

// z [B,768,L] -> vv = sc(v)*sc(x1) and gate = sc(x0), both bf16 [B,256,L]; 8 tokens per thread
__global__ void __launch_bounds__(256) shortconv_gate_kernel(const __nv_bfloat16* __restrict__ z,
                                                             const float* __restrict__ sw, const float* __restrict__ sb,
                                                             int B, int L, __nv_bfloat16* __restrict__ vv,
                                                             __nv_bfloat16* __restrict__ gate) {
  const int per_row = L / 8;
  const size_t total = (size_t)B * 256 * per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int g8 = (int)(idx % per_row);
    const size_t bc = idx / per_row;
    const int c = (int)(bc % 256);
    const size_t b = bc / 256;
    const int t0 = g8 * 8;
    const __nv_bfloat16* zb = z + b * 768 * L;
    float x0[8], x1[8], v[8];
    sconv8(zb + (size_t)c * L, t0, sw + c * 3, __ldg(sb + c), x0);
    sconv8(zb + (size_t)(256 + c) * L, t0, sw + (256 + c) * 3, __ldg(sb + 256 + c), x1);
    sconv8(zb + (size_t)(512 + c) * L, t0, sw + (512 + c) * 3, __ldg(sb + 512 + c), v);
    const size_t off = (b * 256 + c) * L + t0;
    *reinterpret_cast<uint4*>(vv + off) = make_uint4(pack2(v[0] * x1[0], v[1] * x1[1]), pack2(v[2] * x1[2], v[3] * x1[3]),
                                                     pack2(v[4] * x1[4], v[5] * x1[5]), pack2(v[6] * x1[6], v[7] * x1[7]));
    *reinterpret_cast<uint4*>(gate + off) = make_uint4(pack2(x0[0], x0[1]), pack2(x0[2], x0[3]), pack2(x0[4], x0[5]),
                                                       pack2(x0[6], x0[7]));
  }
}

int launch_toeplitz_build(dcb200_ctx* ctx, const float* k, int k_stride, int k_len, const float* D, int nb_max,
                          __nv_bfloat16* T) {
  dim3 grid(nb_max, 256);
  toeplitz_build_kernel<<<grid, 256, 0, ctx->stream>>>(k, k_stride, k_len, D, nb_max, T);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

int launch_shortconv_gate(dcb200_ctx* ctx, const __nv_bfloat16* z, const float* sw, const float* sb, int B, int L,
                          __nv_bfloat16* vv, __nv_bfloat16* gate) {
  const size_t total = (size_t)B * 256 * (L / 8);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 8 * 8;
  if (blocks > cap) blocks = cap;
  ProfScope prof(ctx, K_SCONV);
  shortconv_gate_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(z, sw, sb, B, L, vv, gate);
  DCB_LAUNCH_CHECK(ctx);
  return DCB200_OK;
}

}  // namespace dcb
