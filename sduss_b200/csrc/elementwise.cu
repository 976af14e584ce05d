// HBM-bound kernels of the denoising step: LayerNorm / AdaLN modulation, SiLU, sinusoidal
// timestep embedding, SD3 patchify / unpatchify (latent <-> packed token buffer), and the fused
// CFG-combine + scheduler update. All are coalesced 16-byte-vector streams; row reductions use
// warp shuffles only (one warp per token row).
#include "../../include/sduss_b200.h"
#include <cstdlib>

#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]);
  unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]);
  unpack_bf16x2(u.w, f[6], f[7]);
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
  const int* row_mask; int row_mask_shift;  // patch cache: rows of chunk (row >> shift) with mask 0 are skipped
};

constexpr int LN_MAXIT = 8;   // 8 iterations * 32 lanes * 8 elements = 2048 columns
constexpr int LN_MAXD = LN_MAXIT * 256;
constexpr int LN_WARPS = 8;

// NIT = ceil(D / 256) 16-byte chunks per lane (FULL: D == NIT * 256, no column predicates). A block
// owns a contiguous range of rows (one resident wave); its 8 warps take them round robin and keep
// the next PF rows' vectors in flight (raw bf16) while the current row is reduced and written.
// Every output is y = n * A + B with n = (x - mean) * rstd and per-column A, B folded from
// gamma / beta / (1 + scale) / shift ONCE per block into shared memory as fp32 (for the requests
// of the block's first and last row; a row of a third request - only when requests are shorter
// than a block's row range - folds them on the fly with the same arithmetic). The first version
// of this kernel re-unpacked the row three times and converted the bf16 parameters per element:
// 1250 warp instructions per row, issue-bound at 63 % (profiles/r01_ncu_norm_v2.txt); this one
// unpacks once and spends ~8 instructions per element.
// The arithmetic order per row is fixed (lane-local sums over 4 interleaved accumulators in column
// order, then a shuffle tree), so a row's result does not depend on the batch around it.
constexpr int LN_THREADS = 256;

// A, B of output `which` (0: y, 1: y2) for the 8 columns of chunk c and request g.
__device__ __forceinline__ void ln_fold_params(const LnArgs& a, int g, int c, int which, float* A,
                                               float* B) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    A[j] = 1.f;
    B[j] = 0.f;
  }
  if (a.gamma != nullptr) {
    unpack8(*reinterpret_cast<const uint4*>(a.gamma + c * 8), A);
    unpack8(*reinterpret_cast<const uint4*>(a.beta + c * 8), B);
  }
  if (a.mod != nullptr) {
    const __nv_bfloat16* mrow = a.mod + size_t(g) * a.ldm;
    float sc[8], sh[8];
    unpack8(*reinterpret_cast<const uint4*>(mrow + (which ? a.scale2_col : a.scale_col) + c * 8), sc);
    unpack8(*reinterpret_cast<const uint4*>(mrow + (which ? a.shift2_col : a.shift_col) + c * 8), sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float m = 1.f + sc[j];
      B[j] = fmaf(B[j], m, sh[j]);
      A[j] = A[j] * m;
    }
  }
}

// Row statistics on the unpacked row: v becomes x - mean; returns rstd. Shared by the fast and the
// slow path so both round identically.
template <int NIT, bool FULL>
__device__ __forceinline__ float ln_center(float* v, int lane, int nchunk, float inv_d, float eps) {
  float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < NIT * 8; ++k) s4[k & 3] += v[k];
  const float mean = warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3])) * inv_d;
  float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    if (FULL || lane + 32 * i < nchunk) {  // padded columns must not add mean^2
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[8 * i + j] - mean;
        v[8 * i + j] = d;
        q4[j & 3] = fmaf(d, d, q4[j & 3]);
      }
    }
  }
  return rsqrtf(warp_sum((q4[0] + q4[1]) + (q4[2] + q4[3])) * inv_d + eps);
}

template <int NIT, bool FULL>
__device__ __forceinline__ void ln_unpack_row(const uint4* raw, float* v, int lane, int nchunk) {
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    if (FULL || lane + 32 * i < nchunk) {
      unpack8(raw[i], v + 8 * i);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[8 * i + j] = 0.f;
    }
  }
}

// A row whose request is neither of the two staged ones (requests shorter than a block's row
// range): the whole row again from global memory, parameters folded on the fly. Out of line so
// that its registers do not count against the main loop.
template <int NIT, bool FULL>
__device__ __noinline__ void ln_row_slow(const LnArgs& a, int row, int g, int lane) {
  const int nchunk = a.D >> 3;
  const __nv_bfloat16* xr = a.x + size_t(row) * a.ldx;
  float v[NIT * 8];
  {
    uint4 raw[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i)
      if (FULL || lane + 32 * i < nchunk) raw[i] = *reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8);
    ln_unpack_row<NIT, FULL>(raw, v, lane, nchunk);
  }
  const float rstd = ln_center<NIT, FULL>(v, lane, nchunk, 1.f / float(a.D), a.eps);
  const int nout = a.y2 != nullptr ? 2 : 1;
#pragma unroll 1
  for (int i = 0; i < NIT; ++i) {
    const int c = lane + 32 * i;
    if (FULL || c < nchunk) {
      for (int w = 0; w < nout; ++w) {
        float A[8], B[8], o[8];
        ln_fold_params(a, g, c, w, A, B);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(v[8 * i + j] * rstd, A[j], B[j]);
        __nv_bfloat16* yr = w ? a.y2 + size_t(row) * a.ldy2 : a.y + size_t(row) * a.ldy;
        *reinterpret_cast<uint4*>(yr + c * 8) = pack8(o);
      }
    }
  }
}

template <int NIT, bool FULL, int PF, int MB>
__global__ void __launch_bounds__(LN_THREADS, MB)
ln_mod_kernel(const __grid_constant__ LnArgs a, int rows_per_block) {
  // [slot 0..1][output 0..nout-1][A lo, A hi, B lo, B hi][nchunk] float4 (lo / hi = columns 0-3 /
  // 4-7 of a chunk: consecutive lanes read consecutive 16-byte words, conflict-free)
  extern __shared__ float4 sp[];
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunk = a.D >> 3;
  const bool has_params = a.gamma != nullptr || a.mod != nullptr;
  const int nout = a.y2 != nullptr ? 2 : 1;
  const int block_row0 = blockIdx.x * rows_per_block;
  const int block_end = min(block_row0 + rows_per_block, a.T);
  pdl_launch_dependents();
  pdl_wait();
  int row = block_row0 + wib;
  uint4 cur[NIT], nx[PF][NIT];  // rows row, row + 8 (, row + 16): raw bf16
  auto load_row = [&](uint4* dst, int r) {
    if (r < block_end) {
      const __nv_bfloat16* xr = a.x + size_t(r) * a.ldx;
#pragma unroll
      for (int i = 0; i < NIT; ++i)
        if (FULL || lane + 32 * i < nchunk) dst[i] = *reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8);
    }
  };
  load_row(cur, row);
#pragma unroll
  for (int p = 0; p < PF; ++p) load_row(nx[p], row + (p + 1) * LN_WARPS);
  const bool grouped = a.mod != nullptr && a.row_group != nullptr;
  const int g0 = grouped ? a.row_group[block_row0] : 0;
  const int g1 = grouped ? a.row_group[block_end - 1] : 0;
  if (has_params) {
    for (int slot = 0; slot < (g1 != g0 ? 2 : 1); ++slot) {
      for (int c = threadIdx.x; c < nchunk; c += LN_THREADS) {
        for (int w = 0; w < nout; ++w) {
          float A[8], B[8];
          ln_fold_params(a, slot ? g1 : g0, c, w, A, B);
          float4* d = sp + size_t((slot * nout + w) * 4) * nchunk + c;
          d[0] = make_float4(A[0], A[1], A[2], A[3]);
          d[nchunk] = make_float4(A[4], A[5], A[6], A[7]);
          d[2 * nchunk] = make_float4(B[0], B[1], B[2], B[3]);
          d[3 * nchunk] = make_float4(B[4], B[5], B[6], B[7]);
        }
      }
    }
  }
  __syncthreads();
  const float inv_d = 1.f / float(a.D);
  for (; row < block_end; row += LN_WARPS) {
    float v[NIT * 8];
    ln_unpack_row<NIT, FULL>(cur, v, lane, nchunk);
    // the row after the prefetched ones takes the registers the unpacked row has just left
#pragma unroll
    for (int i = 0; i < NIT; ++i) cur[i] = nx[0][i];
#pragma unroll
    for (int p = 0; p + 1 < PF; ++p)
#pragma unroll
      for (int i = 0; i < NIT; ++i) nx[p][i] = nx[p + 1][i];
    load_row(nx[PF - 1], row + (PF + 1) * LN_WARPS);
    if (a.row_mask != nullptr && a.row_mask[row >> a.row_mask_shift] == 0) continue;  // clean patch: keep y
    const int g = grouped ? a.row_group[row] : 0;
    const int slot = g == g0 ? 0 : (g == g1 ? 1 : -1);
    if (has_params && slot < 0) {  // warp-uniform
      ln_row_slow<NIT, FULL>(a, row, g, lane);
      continue;
    }
    const float rstd = ln_center<NIT, FULL>(v, lane, nchunk, inv_d, a.eps);
    const float4* sp_row = sp + size_t(slot * nout * 4) * nchunk;
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int c = lane + 32 * i;
      if (FULL || c < nchunk) {
        float n[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) n[j] = v[8 * i + j] * rstd;
        for (int w = 0; w < nout; ++w) {
          float o[8];
          if (!has_params) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = n[j];
          } else {
            const float4* d = sp_row + size_t(w * 4) * nchunk + c;
            const float4 a0 = d[0], a1 = d[nchunk], b0 = d[2 * nchunk], b1 = d[3 * nchunk];
            o[0] = fmaf(n[0], a0.x, b0.x); o[1] = fmaf(n[1], a0.y, b0.y);
            o[2] = fmaf(n[2], a0.z, b0.z); o[3] = fmaf(n[3], a0.w, b0.w);
            o[4] = fmaf(n[4], a1.x, b1.x); o[5] = fmaf(n[5], a1.y, b1.y);
            o[6] = fmaf(n[6], a1.z, b1.z); o[7] = fmaf(n[7], a1.w, b1.w);
          }
          __nv_bfloat16* yr = w ? a.y2 + size_t(row) * a.ldy2 : a.y + size_t(row) * a.ldy;
          *reinterpret_cast<uint4*>(yr + c * 8) = pack8(o);
        }
      }
    }
  }
}

template <int NIT, bool FULL, int PF, int MB>
static int launch_ln_v(const LnArgs& a, cudaStream_t st) {
  // one resident wave (MB blocks per SM, see __launch_bounds__); at least one row per warp
  const int max_blocks = MB * device_sm_count();
  int rows_per_block = (a.T + max_blocks - 1) / max_blocks;
  rows_per_block = ((rows_per_block + LN_WARPS - 1) / LN_WARPS) * LN_WARPS;
  const int nout = a.y2 ? 2 : 1;
  const bool has_params = a.gamma || a.mod;
  const size_t smem = has_params ? size_t(2) * nout * 4 * (a.D >> 3) * sizeof(float4) : 0;
  auto kern = ln_mod_kernel<NIT, FULL, PF, MB>;
  static unsigned long long configured = 0;
  if (int rc = ensure_dynamic_smem(kern, 2 * 2 * 4 * LN_MAXIT * 32 * 16, &configured)) return rc;
  return launch_pdl(kern, dim3((a.T + rows_per_block - 1) / rows_per_block), dim3(LN_THREADS), smem, st,
                    a, rows_per_block);
}

// Registers decide the shape: a row of NIT * 8 floats plus (PF + 1) raw rows must stay below the
// limit of MB blocks per SM without spilling (spills cost 25-80 % here, profiles/r01_norm_microbench.txt).
template <int NIT>
static int launch_ln(const LnArgs& a, cudaStream_t st) {
  const bool full = a.D == NIT * 256;
  constexpr int PF = NIT <= 3 ? 2 : 1;
  constexpr int MB = 2;
  return full ? launch_ln_v<NIT, true, PF, MB>(a, st) : launch_ln_v<NIT, false, PF, MB>(a, st);
}

// ------------------------------------------------------------------ row statistics (LN folded into a GEMM)
// stats[row] = (mean, rstd) of x[row, :D]: what is left of a LayerNorm whose affine part has been
// folded into the weights of the GEMM that consumes x (B200EpilogueDesc.ln_stats): one read of x
// and 8 bytes per row written, instead of a full read + write. One warp per row, the row in
// registers, two-pass variance like ln_center.
__global__ void __launch_bounds__(256) row_stats_kernel(const __nv_bfloat16* __restrict__ x, int ldx, int T,
                                                        int D, float eps, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= T) return;
  const int nchunk = D >> 3;
  const __nv_bfloat16* xr = x + size_t(row) * ldx;
  float v[LN_MAXIT * 8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXIT; ++i) {
    if (lane + 32 * i < nchunk) {
      unpack8(*reinterpret_cast<const uint4*>(xr + (lane + 32 * i) * 8), v + 8 * i);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[8 * i + j];
    }
  }
  const float mean = warp_sum(s) / float(D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXIT; ++i) {
    if (lane + 32 * i < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[8 * i + j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  q = warp_sum(q);
  if (lane == 0) stats[row] = make_float2(mean, rsqrtf(q / float(D) + eps));
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

}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;

extern "C" int b200_layernorm_mod_bf16(const void* x, int ldx, int T, int D, float eps,
                                       const void* gamma, const void* beta, const void* mod,
                                       int ldm, const int32_t* row_group, int shift_col,
                                       int scale_col, void* y, int ldy, int shift2_col,
                                       int scale2_col, void* y2, int ldy2, const int32_t* row_mask,
                                       int row_mask_shift, void* stream) {
  if (!x || !y || T <= 0 || D <= 0 || (D & 7) || D > LN_MAXIT * 256 || (ldx & 7) || (ldy & 7))
    return B200_ERR_INVALID;
  if ((gamma == nullptr) != (beta == nullptr)) return B200_ERR_INVALID;
  if (y2 && (!mod || gamma || (ldy2 & 7))) return B200_ERR_INVALID;
  if (mod && ((ldm & 7) || (shift_col & 7) || (scale_col & 7))) return B200_ERR_INVALID;
  LnArgs a{static_cast<const bf16*>(x), ldx, T, D, eps, static_cast<const bf16*>(gamma),
           static_cast<const bf16*>(beta), static_cast<const bf16*>(mod), ldm, row_group,
           shift_col, scale_col, static_cast<bf16*>(y), ldy, shift2_col, scale2_col,
           static_cast<bf16*>(y2), ldy2, row_mask, row_mask_shift};
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

extern "C" int b200_row_stats_bf16(const void* x, int ldx, int T, int D, float eps, float* stats,
                                   void* stream) {
  if (!x || !stats || T <= 0 || D <= 0 || (D & 7) || D > LN_MAXD || (ldx & 7)) return B200_ERR_INVALID;
  return launch_pdl(row_stats_kernel, dim3((T + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
                    static_cast<const bf16*>(x), ldx, T, D, eps, reinterpret_cast<float2*>(stats));
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
