// Kernels the VAE decode stage adds to the denoising-step library (reference
// call sites pipeline_stable_diffusion_xl_esymred.py:406-462, pipeline_stable_diffusion_3_esymred.py:
// 391-415): the latent un-scaling + post_quant_conv as one per-pixel affine map, and the row softmax
// of the decoder's single-head mid-block attention (head_dim = 512 does not fit the packed
// head_dim-64 attention kernel; its two contractions run on the tcgen05 GEMM instead).
#include "../../include/sduss_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

// out_l[co, p] = b[co] + sum_ci W[co, ci] * in_l[ci, p]   (fp32 math, bf16 in / out, NCHW slices)
constexpr int LA_MAXC = 16;
__global__ void __launch_bounds__(256) latent_affine_kernel(const unsigned long long* in_ptr,
                                                            const unsigned long long* out_ptr,
                                                            const int4* desc, int Cin, int Cout,
                                                            const float* W, const float* b) {
  __shared__ float sW[LA_MAXC * LA_MAXC], sb[LA_MAXC];
  for (int i = threadIdx.x; i < Cin * Cout; i += blockDim.x) sW[i] = W[i];
  if (threadIdx.x < Cout) sb[threadIdx.x] = b[threadIdx.x];
  pdl_launch_dependents();
  pdl_wait();
  __syncthreads();
  const int l = blockIdx.y;
  const int4 d = desc[l];
  const int px = d.y * d.z;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= px) return;
  const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(in_ptr[l]);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(out_ptr[l]);
  float x[LA_MAXC];
#pragma unroll
  for (int c = 0; c < LA_MAXC; ++c)
    if (c < Cin) x[c] = __bfloat162float(in[size_t(c) * px + p]);
  for (int co = 0; co < Cout; ++co) {
    float acc = sb[co];
#pragma unroll
    for (int c = 0; c < LA_MAXC; ++c)
      if (c < Cin) acc = fmaf(sW[co * Cin + c], x[c], acc);
    out[size_t(co) * px + p] = __float2bfloat16(acc);
  }
}

// P[r, :] = softmax(scale * S[r, :]) ; S fp32 (GEMM output, fp32 keeps the logits exact), P bf16.
// One CTA per row, the row lives in registers (cols <= 16384), two block reductions.
constexpr int SM_THREADS = 256;
constexpr int SM_MAXV = 16;  // float4 per thread
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < SM_THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SM_THREADS) softmax_rows_kernel(const float* S, long lds, int cols,
                                                                  float scale_log2,
                                                                  __nv_bfloat16* P, long ldp) {
  __shared__ float red[SM_THREADS / 32];
  pdl_launch_dependents();
  pdl_wait();
  const float4* s = reinterpret_cast<const float4*>(S + size_t(blockIdx.x) * lds);
  const int nv = cols >> 2;
  float4 v[SM_MAXV];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    const int i = threadIdx.x + k * SM_THREADS;
    if (i < nv) {
      v[k] = s[i];
      mx = fmaxf(fmaxf(mx, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
  }
  mx = block_reduce(mx, true, red);
  const float off = -mx * scale_log2;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    if (threadIdx.x + k * SM_THREADS < nv) {
      v[k].x = fast_exp2(fmaf(v[k].x, scale_log2, off));
      v[k].y = fast_exp2(fmaf(v[k].y, scale_log2, off));
      v[k].z = fast_exp2(fmaf(v[k].z, scale_log2, off));
      v[k].w = fast_exp2(fmaf(v[k].w, scale_log2, off));
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  sum = block_reduce(sum, false, red);
  const float inv = 1.f / sum;
  uint2* p = reinterpret_cast<uint2*>(P + size_t(blockIdx.x) * ldp);
#pragma unroll
  for (int k = 0; k < SM_MAXV; ++k) {
    const int i = threadIdx.x + k * SM_THREADS;
    if (i < nv)
      p[i] = make_uint2(pack_bf16x2(v[k].x * inv, v[k].y * inv), pack_bf16x2(v[k].z * inv, v[k].w * inv));
  }
}

}  // namespace b200

using namespace b200;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int b200_latent_affine(const uint64_t* in_ptr, const uint64_t* out_ptr,
                                  const int32_t* desc, int n_latents, int max_pixels, int c_in,
                                  int c_out, const float* weight, const float* bias, void* stream) {
  if (!in_ptr || !out_ptr || !desc || !weight || !bias || n_latents <= 0 || max_pixels <= 0 ||
      c_in <= 0 || c_in > LA_MAXC || c_out <= 0 || c_out > LA_MAXC)
    return B200_ERR_INVALID;
  return launch_pdl(latent_affine_kernel, dim3((max_pixels + 255) / 256, n_latents), dim3(256), 0,
                    ST(stream), reinterpret_cast<const unsigned long long*>(in_ptr),
                    reinterpret_cast<const unsigned long long*>(out_ptr),
                    reinterpret_cast<const int4*>(desc), c_in, c_out, weight, bias);
}

extern "C" int b200_softmax_rows(const float* s, long long lds, int rows, int cols, float scale,
                                 void* p, long long ldp, void* stream) {
  if (!s || !p || rows <= 0 || cols <= 0 || (cols & 3) || cols > SM_MAXV * SM_THREADS * 4 ||
      (lds & 3) || (ldp & 3) || lds < cols || ldp < cols)
    return B200_ERR_INVALID;
  return launch_pdl(softmax_rows_kernel, dim3(rows), dim3(SM_THREADS), 0, ST(stream), s, long(lds),
                    cols, scale * 1.4426950408889634f, static_cast<__nv_bfloat16*>(p), long(ldp));
}
