// Prepare-stage kernels (text encoders): token / position embedding gather and
// the T5 RMS layer norm. Both are HBM-bound row streams: one warp per row, 16-byte vectors.
// (CLIP's LayerNorm, every linear layer and the attention run on the denoising step's kernels:
// b200_layernorm_mod_bf16, b200_gemm_bf16, b200_attn_varlen_ex.)
#include <cuda_bf16.h>

#include "../../include/sduss_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

// out[i, :] = table[ids[i], :] (+ pos[i % seq_len, :])          CLIPTextEmbeddings / T5 embed_tokens
__global__ void embed_rows_kernel(const int* __restrict__ ids, const __nv_bfloat16* __restrict__ table,
                                  int vocab, const __nv_bfloat16* __restrict__ pos, int seq_len,
                                  int n, int D, __nv_bfloat16* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint4* src = reinterpret_cast<const uint4*>(table + size_t(id) * D);
  const uint4* ps = pos ? reinterpret_cast<const uint4*>(pos + size_t(row % seq_len) * D) : nullptr;
  uint4* dst = reinterpret_cast<uint4*>(out + size_t(row) * ldo);
  for (int c = lane; c < D / 8; c += 32) {
    uint4 v = src[c];
    if (ps) {
      const uint4 p = ps[c];
      float a[2], b[2];
      uint32_t* vv = &v.x;
      const uint32_t* pp = &p.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        unpack_bf16x2(vv[k], a[0], a[1]);
        unpack_bf16x2(pp[k], b[0], b[1]);
        vv[k] = pack_bf16x2(a[0] + b[0], a[1] + b[1]);
      }
    }
    dst[c] = v;
  }
}

// y = weight * x * rsqrt(mean(x^2) + eps)           T5LayerNorm (no mean subtraction, no bias)
// One warp per row; the row is read twice (second read from L1/L2: rows are <= 8 KB).
__global__ void rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int T, int D, float eps,
                               const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y,
                               int ldy) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + size_t(row) * ldx);
  float ss = 0.f;
  for (int c = lane; c < D / 8; c += 32) {
    const uint4 v = xr[c];
    const uint32_t* vv = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a, b;
      unpack_bf16x2(vv[k], a, b);
      ss += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / float(D) + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + size_t(row) * ldy);
  for (int c = lane; c < D / 8; c += 32) {
    uint4 v = xr[c];
    const uint4 g = wr[c];
    uint32_t* vv = &v.x;
    const uint32_t* gg = &g.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a, b, ga, gb;
      unpack_bf16x2(vv[k], a, b);
      unpack_bf16x2(gg[k], ga, gb);
      // transformers: hidden_states * rsqrt(var) in fp32, cast to the weight dtype, then * weight
      const float na = __bfloat162float(__float2bfloat16(a * r));
      const float nb = __bfloat162float(__float2bfloat16(b * r));
      vv[k] = pack_bf16x2(ga * na, gb * nb);
    }
    yr[c] = v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_embed_rows_bf16(const int32_t* ids, int n, const void* table, int vocab, int D,
                                    const void* pos, int seq_len, void* out, int ldo, void* stream) {
  if (!ids || !table || !out || n <= 0 || vocab <= 0 || D <= 0 || (D & 7) || (ldo & 7) ||
      (pos && seq_len <= 0))
    return B200_ERR_INVALID;
  return launch_pdl(embed_rows_kernel, dim3((n + 7) / 8), dim3(256), 0,
                    reinterpret_cast<cudaStream_t>(stream), ids, static_cast<const __nv_bfloat16*>(table),
                    vocab, static_cast<const __nv_bfloat16*>(pos), seq_len > 0 ? seq_len : 1, n, D,
                    static_cast<__nv_bfloat16*>(out), ldo);
}

extern "C" int b200_rmsnorm_bf16(const void* x, int ldx, int T, int D, float eps, const void* weight,
                                 void* y, int ldy, void* stream) {
  if (!x || !y || !weight || T <= 0 || D <= 0 || (D & 7) || (ldx & 7) || (ldy & 7))
    return B200_ERR_INVALID;
  return launch_pdl(rmsnorm_kernel, dim3((T + 7) / 8), dim3(256), 0,
                    reinterpret_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x), ldx, T,
                    D, eps, static_cast<const __nv_bfloat16*>(weight), static_cast<__nv_bfloat16*>(y), ldy);
}
