// HBM-bound kernels of the denoising step: LayerNorm / AdaLN modulation, SiLU, sinusoidal
// timestep embedding, SD3 patchify / unpatchify (latent <-> packed token buffer), and the fused
// CFG-combine + scheduler update. All are coalesced 16-byte-vector streams; row reductions use
// warp shuffles only (one warp per token row).
#include "../../include/sduss_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------ LayerNorm (+ modulation)
// y  = LN(x) [* gamma + beta] [* (1 + scale[g]) + shift[g]]
// y2 = LN(x) * (1 + scale2[g]) + shift2[g]          (optional second output, SD35 AdaLN-Zero-X)
// Replaces diffusers LayerNorm / AdaLayerNormZero / SD35AdaLayerNormZeroX /
// AdaLayerNormContinuous as called from sduss modules/transformer.py:185-279,317-386 and
// SD3Transformer.py:238. One warp per row; the row lives in registers (D <= 2048).
struct LnArgs {
  const __nv_bfloat16* x; int ldx;
  int T, D; float eps;
  const __nv_bfloat16* gamma; const __nv_bfloat16* beta;
  const __nv_bfloat16* mod; int ldm; const int* row_group;
  int shift_col, scale_col;
  __nv_bfloat16* y; int ldy;
  int shift2_col, scale2_col;
  __nv_bfloat16* y2; int ldy2;
};

constexpr int LN_MAXIT = 8;   // 8 iterations * 32 lanes * 8 elements = 2048 columns
constexpr int LN_MAXD = LN_MAXIT * 256;
constexpr int LN_WARPS = 8;

// NIT = ceil(D / 256) 16-byte chunks per lane. A block owns a contiguous range of rows; its 8
// warps take them round robin and keep the next TWO rows' vectors in flight (raw bf16, 4 registers
// per chunk) while they reduce and write the current one (~100 KB of loads outstanding per SM); the first row is requested before the parameters are staged. The grid is sized
// to one resident wave (the first version ran 16 rows per block, 3 waves of a load -> reduce ->
// reduce -> store chain with 128 registers: 1.9 TB/s, profiles/r01_ncu_gemm_ln_sd3.txt).
// The per-column parameters (gamma/beta or the per-request shift/scale vectors, up to 4x the
// bytes of a row) are staged ONCE per block in shared memory for the request of the block's first
// row; rows of another request (only at request boundaries) read them from global memory.
// The arithmetic order per row is fixed (lane-local sums in column order, then a shuffle tree), so
// a row's result does not depend on the batch around it.
template <int NIT>
__global__ void __launch_bounds__(256, (NIT <= 4 ? 3 : 2)) ln_mod_kernel(LnArgs a, int rows_per_block) {
  __shared__ uint4 sp[4][LN_MAXD / 8];  // [scale|gamma, shift|beta, scale2, shift2][chunk]
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunk = a.D >> 3;
  const bool affine = a.gamma != nullptr;
  const bool modulated = a.mod != nullptr;
  const bool dual = a.y2 != nullptr;
  const int block_row0 = blockIdx.x * rows_per_block;
  const int block_end = min(block_row0 + rows_per_block, a.T);
  if (affine) {  // weights: not written by the preceding kernel
    for (int c = threadIdx.x; c < nchunk; c += 256) {
      sp[0][c] = *reinterpret_cast<const uint4*>(a.gamma + c * 8);
      sp[1][c] = *reinterpret_cast<const uint4*>(a.beta + c * 8);
    }
  }
  pdl_launch_dependents();
  pdl_wait();
  int row = block_row0 + wib;
  uint4 cur[NIT], nx1[NIT];  // rows row and row + 8; row + 16 is requested inside the loop
  auto load_row = [&](uint4* dst, int r) {
    if (r < block_end) {
      const __nv_bfloat16* xr = a.x + size_t(r) * a.ldx;
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        if (lane + 32 * i < nchunk) dst[i] = *reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8);
    }
  };
  load_row(cur, row);
  load_row(nx1, row + LN_WARPS);
  const int g0 = (modulated && a.row_group) ? a.row_group[block_row0] : 0;
  if (modulated) {
    const __nv_bfloat16* mrow = a.mod + size_t(g0) * a.ldm;
    for (int c = threadIdx.x; c < nchunk; c += 256) {
      sp[affine ? 2 : 0][c] = *reinterpret_cast<const uint4*>(mrow + a.scale_col + c * 8);
      sp[affine ? 3 : 1][c] = *reinterpret_cast<const uint4*>(mrow + a.shift_col + c * 8);
      if (dual) {
        sp[2][c] = *reinterpret_cast<const uint4*>(mrow + a.scale2_col + c * 8);
        sp[3][c] = *reinterpret_cast<const uint4*>(mrow + a.shift2_col + c * 8);
      }
    }
  }
  __syncthreads();
  const float inv_d = 1.f / float(a.D);
  const int ms = affine ? 2 : 0;  // smem slot of scale (modulation) -- affine+dual is rejected by the host
  for (; row < block_end; row += LN_WARPS) {
    uint4 nx2[NIT];
    load_row(nx2, row + 2 * LN_WARPS);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (lane + 32 * i < nchunk) {
        float v[8];
        unpack8(cur[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j];
      }
    }
    const int g = (modulated && a.row_group) ? a.row_group[row] : 0;
    const bool fast = g == g0;
    const __nv_bfloat16* mrow = modulated ? a.mod + size_t(g) * a.ldm : nullptr;
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      if (lane + 32 * i < nchunk) {
        float v[8];
        unpack8(cur[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[j] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + a.eps);
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int c = lane + 32 * i;
      if (c < nchunk) {
        float v[8], n[8], o[8];
        unpack8(cur[i], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) n[j] = (v[j] - mean) * rstd;
        if (affine) {
          float gg[8], bb[8];
          unpack8(sp[0][c], gg);
          unpack8(sp[1][c], bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) n[j] = n[j] * gg[j] + bb[j];
        }
        if (modulated) {
          float sc[8], sh[8];
          unpack8(fast ? sp[ms][c] : *reinterpret_cast<const uint4*>(mrow + a.scale_col + c * 8), sc);
          unpack8(fast ? sp[ms + 1][c] : *reinterpret_cast<const uint4*>(mrow + a.shift_col + c * 8), sh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = n[j] * (1.f + sc[j]) + sh[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = n[j];
        }
        *reinterpret_cast<uint4*>(a.y + size_t(row) * a.ldy + c * 8) = pack8(o);
        if (dual) {
          float sc[8], sh[8];
          unpack8(fast ? sp[2][c] : *reinterpret_cast<const uint4*>(mrow + a.scale2_col + c * 8), sc);
          unpack8(fast ? sp[3][c] : *reinterpret_cast<const uint4*>(mrow + a.shift2_col + c * 8), sh);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = n[j] * (1.f + sc[j]) + sh[j];
          *reinterpret_cast<uint4*>(a.y2 + size_t(row) * a.ldy2 + c * 8) = pack8(o);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      cur[i] = nx1[i];
      nx1[i] = nx2[i];
    }
  }
}

template <int NIT>
static int launch_ln(const LnArgs& a, cudaStream_t st) {
  // one resident wave (3 or 2 blocks per SM, see __launch_bounds__); at least one row per warp
  const int max_blocks = (NIT <= 4 ? 3 : 2) * device_sm_count();
  int rows_per_block = (a.T + max_blocks - 1) / max_blocks;
  rows_per_block = ((rows_per_block + LN_WARPS - 1) / LN_WARPS) * LN_WARPS;
  return launch_pdl(ln_mod_kernel<NIT>, dim3((a.T + rows_per_block - 1) / rows_per_block),
                    dim3(256), 0, st, a, rows_per_block);
}

// ------------------------------------------------------------------ small elementwise
__global__ void silu_kernel(const __nv_bfloat16* x, __nv_bfloat16* y, long n8) {
  long i = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (i >= n8) return;
  float f[8];
  unpack8(reinterpret_cast<const uint4*>(x)[i], f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = f[j] / (1.f + __expf(-f[j]));
  reinterpret_cast<uint4*>(y)[i] = pack8(f);
}

// out[i, :] = [cos(t_i f_k) | sin(t_i f_k)], f_k = exp(-ln(1e4) k / half)  (flip_sin_to_cos=True,
// downscale_freq_shift=0: diffusers Timesteps as used at sduss modules/unet.py:314 and
// SD3Transformer.py:81). fp32 math, bf16 out.
__global__ void timestep_embedding_kernel(const float* t, int n, int dim, __nv_bfloat16* out,
                                          int ldo) {
  const int half = dim >> 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  const float freq = expf(-9.210340371976184f * float(k) / float(half));
  const float arg = t[i] * freq;
  float sn, cs;
  sincosf(arg, &sn, &cs);
  out[size_t(i) * ldo + k] = __float2bfloat16(cs);
  out[size_t(i) * ldo + half + k] = __float2bfloat16(sn);
}

// ------------------------------------------------------------------ SD3 patchify / unpatchify
// desc[i] = {token offset of latent i, tokens per row (w/p), tokens per column (h/p), unused}
// lat_ptr[i] = device pointer of latent i: [C, h, w] bf16, contiguous (NCHW slice).
// patchify: tokens[tok, c*p*p + py*p + px] = lat[c, ty*p + py, tx*p + px]  (conv k=p, s=p im2col;
// the PatchEmbed conv of SD3Transformer.py:82-83 then becomes a GEMM with K = C*p*p).
__global__ void sd3_patchify_kernel(const unsigned long long* lat_ptr, const int4* desc, int C,
                                    int p, __nv_bfloat16* tokens, int ldt) {
  const int lat = blockIdx.y;
  const int4 d = desc[lat];
  const int wt = d.y, ht = d.z;
  const int K = C * p * p;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wt * ht * K) return;
  const int tok = idx / K, k = idx % K;
  const int c = k / (p * p), py = (k / p) % p, px = k % p;
  const int ty = tok / wt, tx = tok % wt;
  const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(lat_ptr[lat]);
  tokens[size_t(d.x + tok) * ldt + k] = src[(size_t(c) * ht * p + ty * p + py) * (wt * p) + tx * p + px];
}

// unpatchify: out[c, ty*p+py, tx*p+px] = tokens[tok, (py*p+px)*C + c]
// (einsum "nhwpqc->nchpwq", SD3Transformer.py:250-259)
__global__ void sd3_unpatchify_kernel(const __nv_bfloat16* tokens, int ldt, const int4* desc, int C,
                                      int p, const unsigned long long* out_ptr) {
  const int lat = blockIdx.y;
  const int4 d = desc[lat];
  const int wt = d.y, ht = d.z;
  const int W = wt * p, Hh = ht * p;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * Hh * W) return;
  const int c = idx / (Hh * W), y = (idx / W) % Hh, x = idx % W;
  const int tok = (y / p) * wt + x / p;
  const int k = ((y % p) * p + (x % p)) * C + c;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out_ptr[lat]);
  dst[idx] = tokens[size_t(d.x + tok) * ldt + k];
}

// ------------------------------------------------------------------ CFG + scheduler update
// Per request r (one CTA column): eps = u + g (c - u) when cfg, then
//   mode 0 (flow match, scheduling_flow_match_euler_discrete.py:159-203): x' = x + (s' - s) eps
//   mode 1 (Euler epsilon,  scheduling_euler_discrete.py:187-274): x0 = x - s eps; d = (x - x0)/s;
//          x' = x + d (s' - s)
//   mode 2 (Euler v_prediction): x0 = eps * (-s / sqrt(s^2+1)) + x / (s^2+1); rest as mode 1
// fp32 math on the fp32-upcast sample, result cast to the model dtype (bf16) exactly like the
// reference; the CFG combine happens in bf16 arithmetic order of the reference (u + g*(c-u)
// evaluated in fp32 then rounded to bf16).
// desc[r] = {element offset of request r in x / out, elements, offset of uncond in eps, offset
// of cond in eps}; sig[r] = {sigma, sigma_next}.
struct StepArgs {
  const __nv_bfloat16* eps; const __nv_bfloat16* x; __nv_bfloat16* out;
  const long long* desc;  // [R][4]
  const float* sig;       // [R][2]
  float guidance; int cfg; int mode;
};

__global__ void cfg_step_kernel(StepArgs a) {
  const int r = blockIdx.y;
  const long long x_off = a.desc[r * 4], n = a.desc[r * 4 + 1];
  const long long u_off = a.desc[r * 4 + 2], c_off = a.desc[r * 4 + 3];
  const float s = a.sig[r * 2], sn = a.sig[r * 2 + 1];
  for (long long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n;
       i += long(gridDim.x) * blockDim.x) {
    float e;
    if (a.cfg) {
      const float u = __bfloat162float(a.eps[u_off + i]);
      const float c = __bfloat162float(a.eps[c_off + i]);
      // reference: bf16 tensor ops -> each op rounds to bf16
      const float diff = __bfloat162float(__float2bfloat16(c - u));
      const float sc = __bfloat162float(__float2bfloat16(a.guidance * diff));
      e = __bfloat162float(__float2bfloat16(u + sc));
    } else {
      e = __bfloat162float(a.eps[c_off + i]);
    }
    const float x = __bfloat162float(a.x[x_off + i]);
    float xn;
    // explicit round-to-nearest mul/add (no FMA contraction): same op order as the reference
    if (a.mode == 0) {
      xn = __fadd_rn(x, __fmul_rn(__fsub_rn(sn, s), e));
    } else {
      float x0;
      if (a.mode == 1) {
        x0 = __fsub_rn(x, __fmul_rn(s, e));
      } else {
        const float s2p = __fadd_rn(__fmul_rn(s, s), 1.f);
        x0 = __fadd_rn(__fmul_rn(e, __fdiv_rn(-s, __fsqrt_rn(s2p))), __fdiv_rn(x, s2p));
      }
      const float d = __fdiv_rn(__fsub_rn(x, x0), s);
      xn = __fadd_rn(x, __fmul_rn(d, __fsub_rn(sn, s)));
    }
    a.out[x_off + i] = __float2bfloat16(xn);
  }
}

// x / sqrt(sigma^2 + 1), sigma per latent (scheduling_euler_discrete.py:161-184). The reference
// builds sigma in the sample dtype, so sigma is rounded to bf16 first.
// desc[l] = {element offset, elements}; sig[l] = sigma of latent l.
__global__ void scale_input_kernel(const __nv_bfloat16* x, __nv_bfloat16* y, const long long* desc,
                                   const float* sig) {
  const int l = blockIdx.y;
  const long long off = desc[l * 2], n = desc[l * 2 + 1];
  const float s = __bfloat162float(__float2bfloat16(sig[l]));
  // bf16 op chain of the reference: s**2 -> +1 -> **0.5 -> divide, each rounded to bf16
  const float s2 = __bfloat162float(__float2bfloat16(s * s));
  const float s2p = __bfloat162float(__float2bfloat16(s2 + 1.f));
  const float den = __bfloat162float(__float2bfloat16(sqrtf(s2p)));
  for (long long i = blockIdx.x * long(blockDim.x) + threadIdx.x; i < n;
       i += long(gridDim.x) * blockDim.x)
    y[off + i] = __float2bfloat16(__fdiv_rn(__bfloat162float(x[off + i]), den));
}

}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;

extern "C" int b200_layernorm_mod_bf16(const void* x, int ldx, int T, int D, float eps,
                                       const void* gamma, const void* beta, const void* mod,
                                       int ldm, const int32_t* row_group, int shift_col,
                                       int scale_col, void* y, int ldy, int shift2_col,
                                       int scale2_col, void* y2, int ldy2, void* stream) {
  if (!x || !y || T <= 0 || D <= 0 || (D & 7) || D > LN_MAXIT * 256 || (ldx & 7) || (ldy & 7))
    return B200_ERR_INVALID;
  if ((gamma == nullptr) != (beta == nullptr)) return B200_ERR_INVALID;
  if (y2 && (!mod || gamma || (ldy2 & 7))) return B200_ERR_INVALID;
  if (mod && ((ldm & 7) || (shift_col & 7) || (scale_col & 7))) return B200_ERR_INVALID;
  LnArgs a{static_cast<const bf16*>(x), ldx, T, D, eps, static_cast<const bf16*>(gamma),
           static_cast<const bf16*>(beta), static_cast<const bf16*>(mod), ldm, row_group,
           shift_col, scale_col, static_cast<bf16*>(y), ldy, shift2_col, scale2_col,
           static_cast<bf16*>(y2), ldy2};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch ((D + 255) / 256) {
    case 1: return launch_ln<1>(a, st);
    case 2: return launch_ln<2>(a, st);
    case 3: return launch_ln<3>(a, st);
    case 4: return launch_ln<4>(a, st);
    case 5: return launch_ln<5>(a, st);
    case 6: return launch_ln<6>(a, st);
    case 7: return launch_ln<7>(a, st);
    default: return launch_ln<8>(a, st);
  }
}

extern "C" int b200_silu_bf16(const void* x, void* y, long long n, void* stream) {
  if (!x || !y || n <= 0 || (n & 7)) return B200_ERR_INVALID;
  const long n8 = n >> 3;
  silu_kernel<<<unsigned((n8 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), n8);
  return launch_status();
}

extern "C" int b200_timestep_embedding(const float* t, int n, int dim, void* out, int ldo,
                                       void* stream) {
  if (!t || !out || n <= 0 || dim <= 0 || (dim & 1)) return B200_ERR_INVALID;
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      t, n, dim, static_cast<bf16*>(out), ldo);
  return launch_status();
}

extern "C" int b200_sd3_patchify(const uint64_t* lat_ptr, const int32_t* desc, int n_latents,
                                 int max_tokens, int C, int p, void* tokens, int ldt, void* stream) {
  if (!lat_ptr || !desc || !tokens || n_latents <= 0 || max_tokens <= 0) return B200_ERR_INVALID;
  const int K = C * p * p;
  dim3 grid((unsigned(max_tokens) * K + 255) / 256, n_latents);
  sd3_patchify_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(lat_ptr), reinterpret_cast<const int4*>(desc), C,
      p, static_cast<bf16*>(tokens), ldt);
  return launch_status();
}

extern "C" int b200_sd3_unpatchify(const void* tokens, int ldt, const int32_t* desc, int n_latents,
                                   int max_tokens, int C, int p, const uint64_t* out_ptr,
                                   void* stream) {
  if (!out_ptr || !desc || !tokens || n_latents <= 0 || max_tokens <= 0) return B200_ERR_INVALID;
  dim3 grid((unsigned(max_tokens) * C * p * p + 255) / 256, n_latents);
  sd3_unpatchify_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(tokens), ldt, reinterpret_cast<const int4*>(desc), C, p,
      reinterpret_cast<const unsigned long long*>(out_ptr));
  return launch_status();
}

extern "C" int b200_cfg_scheduler_step(const void* eps, const void* x, void* out,
                                       const int64_t* desc, const float* sigmas, int n_requests,
                                       long long max_elems, float guidance, int cfg, int mode,
                                       void* stream) {
  if (!eps || !x || !out || !desc || !sigmas || n_requests <= 0 || max_elems <= 0 || mode < 0 ||
      mode > 2)
    return B200_ERR_INVALID;
  StepArgs a{static_cast<const bf16*>(eps), static_cast<const bf16*>(x), static_cast<bf16*>(out),
             reinterpret_cast<const long long*>(desc), sigmas, guidance, cfg, mode};
  unsigned gx = unsigned((max_elems + 1023) / 1024);
  if (gx > 256) gx = 256;
  cfg_step_kernel<<<dim3(gx, n_requests), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  return launch_status();
}

extern "C" int b200_euler_scale_input(const void* x, void* y, const int64_t* desc,
                                      const float* sigmas, int n_latents, long long max_elems,
                                      void* stream) {
  if (!x || !y || !desc || !sigmas || n_latents <= 0 || max_elems <= 0) return B200_ERR_INVALID;
  unsigned gx = unsigned((max_elems + 1023) / 1024);
  if (gx > 256) gx = 256;
  scale_input_kernel<<<dim3(gx, n_latents), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(y), reinterpret_cast<const long long*>(desc),
      sigmas);
  return launch_status();
}
