// HBM-bound kernels of the SDXL UNet path on the packed NHWC layout [sum_i H_i*W_i, C]:
// per-request GroupNorm(+SiLU), latent pack (NCHW -> im2col rows for conv_in), noise scatter
// (NHWC -> NCHW), nearest 2x upsample, column-block copy (skip concat), and the reference's own
// patch pack/scatter format (split_sample / concat_sample) for index-exact interoperability.
#include "../../include/sduss_b200.h"
#include <cstdlib>

#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

__device__ __forceinline__ void unpack8s(const uint4& u, float* f) {
  unpack_bf16x2(u.x, f[0], f[1]);
  unpack_bf16x2(u.y, f[2], f[3]);
  unpack_bf16x2(u.z, f[4], f[5]);
  unpack_bf16x2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8s(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------------ GroupNorm
// Exact per-request statistics over the WHOLE latent (the reference approximates the variance by
// the mean of per-patch variances and races while doing so: kernels/norm_silu_concat.cu:361-386,
// deviations D1 and the in-place race). Three launches:
//   stats   : CTA (chunk, slice) -> partial (sum, sumsq) per group of the slice    [G][T/64][2]
//   finalize: one warp per (latent, group): fixed-order fp64 reduction              [L][G][2]
//   apply   : y = x * a + b per channel (a, b folded from mean, rstd, gamma, beta once per
//             thread), optional SiLU, bf16
// The two streaming kernels share one geometry: a CTA owns a 64-row chunk of ONE latent and a
// slice of whole groups (S slices, S a function of C only), every thread owns ONE 16-byte channel
// vector and the rows rl, rl + RL, ... of the chunk, so per-channel state lives in registers and a
// thread's loads are independent (4 in flight, 4-5 CTAs per SM). The summation order is a function
// of (C, row index inside the latent) only - never of the launch geometry - so a latent's output
// does not depend on what is packed around it (batch invariance); no atomics (determinism).
constexpr int GN_CHUNK = 64;   // rows per CTA == host-side unit of the latent tables (layout.py)
constexpr int GN_MAXG = 32;
constexpr int GN_MAXC = 2560;
constexpr int GN_THREADS = 256;

struct GnMap {
  int S;        // channel slices (grid.y); a slice holds G / S whole groups
  int W;        // 16-byte vector columns per slice
  int RL;       // row lanes: threads sharing one column
  int threads;  // RL * W
};
inline GnMap gn_map(int C, int G) {
  const int nvec = C >> 3, cpg = C / G;
  GnMap m;
  m.S = 1;
  for (int s = 8; s > 1; s >>= 1)
    if (G % s == 0 && ((G / s) * cpg) % 8 == 0 && nvec / s >= 32) {
      m.S = s;
      break;
    }
  m.W = nvec / m.S;
  m.RL = m.W <= GN_THREADS ? GN_THREADS / m.W : 0;
  if (m.RL > GN_CHUNK) m.RL = GN_CHUNK;
  m.threads = m.RL * m.W;
  return m;
}

// ---- the three phases as device functions: the three-launch path and the single-launch kernel below
// run the SAME code per (chunk, slice) / per (latent, group), so a latent's result does not depend on
// which of the two ran (the choice is a function of the batch size).
__device__ __forceinline__ void gn_stats_item(const __nv_bfloat16* x, int ldx, int W, int RL, int cpg,
                                              int gps, int n_chunks, float* partial, int chunk, int slice,
                                              float* ch_sum, float* ch_sq) {
  const int rl = threadIdx.x / W, col = threadIdx.x - rl * W;
  const int CW = W * 8;  // channels of the slice
  const __nv_bfloat16* base = x + size_t(chunk) * GN_CHUNK * ldx + size_t(slice) * CW + col * 8;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = rl + k * RL;
    if (r < GN_CHUNK) v[k] = *reinterpret_cast<const uint4*>(base + size_t(r) * ldx);
  }
  for (int r0 = rl;;) {  // fixed order: rows rl, rl + RL, ...
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r0 + k * RL < GN_CHUNK) {
        float f[8];
        unpack8s(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sum[j] += f[j];
          sq[j] = fmaf(f[j], f[j], sq[j]);
        }
      }
    }
    r0 += 4 * RL;
    if (r0 >= GN_CHUNK) break;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k * RL;
      if (r < GN_CHUNK) v[k] = *reinterpret_cast<const uint4*>(base + size_t(r) * ldx);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ch_sum[rl * CW + col * 8 + j] = sum[j];
    ch_sq[rl * CW + col * 8 + j] = sq[j];
  }
  __syncthreads();
  if (RL > 1) {
    for (int c = threadIdx.x; c < CW; c += blockDim.x) {
      float s = ch_sum[c], q = ch_sq[c];
      for (int l = 1; l < RL; ++l) {
        s += ch_sum[l * CW + c];
        q += ch_sq[l * CW + c];
      }
      ch_sum[c] = s;
      ch_sq[c] = q;
    }
    __syncthreads();
  }
  if (threadIdx.x < gps) {
    float s = 0.f, q = 0.f;
    for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
      s += ch_sum[c];
      q += ch_sq[c];
    }
    const int g = slice * gps + threadIdx.x;
    *reinterpret_cast<float2*>(partial + (size_t(g) * n_chunks + chunk) * 2) = make_float2(s, q);
  }
}

__global__ void __launch_bounds__(GN_THREADS, 5) gn_stats_kernel(const __nv_bfloat16* x, int ldx,
                                                                 int W, int RL, int cpg,
                                                                 int gps, int n_chunks,
                                                                 float* partial) {
  __shared__ float ch_sum[GN_THREADS * 8], ch_sq[GN_THREADS * 8];  // [RL][W * 8]
  pdl_launch_dependents();
  pdl_wait();
  gn_stats_item(x, ldx, W, RL, cpg, gps, n_chunks, partial, blockIdx.x, blockIdx.y, ch_sum, ch_sq);
}

// lat: [L][4] = {first 64-row chunk, number of chunks, 0, 0}. One warp per (latent, group): lanes
// stride the latent's chunks (coalesced float2 reads), then a fixed-order fp64 shuffle tree.
__device__ __forceinline__ void gn_finalize_one(const float* partial, const int4* lat, int G, int cpg,
                                                int n_chunks, float eps, float* stats, int idx, int lane) {
  const int l = idx / G, g = idx % G;
  const int4 d = lat[l];
  const float2* p = reinterpret_cast<const float2*>(partial) + size_t(g) * n_chunks + d.x;
  double s = 0.0, q = 0.0;
#pragma unroll 4
  for (int c = lane; c < d.y; c += 32) {
    const float2 v = __ldcg(p + c);  // written by other SMs (of this launch, in the single-launch kernel)
    s += double(v.x);
    q += double(v.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (lane == 0) {
    const double n = double(d.y) * GN_CHUNK * cpg;
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[idx * 2] = float(mean);
    stats[idx * 2 + 1] = float(1.0 / sqrt(var + double(eps)));
  }
}

__global__ void gn_finalize_kernel(const float* partial, const int4* lat, int L, int G, int cpg,
                                   int n_chunks, float eps, float* stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (idx >= L * G) return;
  gn_finalize_one(partial, lat, G, cpg, n_chunks, eps, stats, idx, threadIdx.x & 31);
}

// Finalize from the statistics the producing convolution left behind (conv_sm100.cu, stats_out):
// part: [n_mtiles * 2][C] float2 per (16x8-pixel tile, row half, channel). One CTA per (latent,
// group): thread t adds the entries t, t + 256, ... of the latent's tiles x halves x the group's
// channels in fp64, then a fixed shared-memory tree. lat_tiles: [L][4] = {first tile, tiles, pixels, 0}.
__global__ void __launch_bounds__(256) gn_finalize_tiles_kernel(const float2* part, const int4* lat_tiles,
                                                                int G, int cpg, int C, float eps,
                                                                float* stats) {
  pdl_launch_dependents();
  pdl_wait();
  const int l = blockIdx.x / G, g = blockIdx.x % G;
  const int4 d = lat_tiles[l];
  const long n_entries = long(d.y) * 2 * cpg;
  const float2* base = part + size_t(d.x) * 2 * C + g * cpg;
  double s = 0.0, q = 0.0;
  for (long i = threadIdx.x; i < n_entries; i += 256) {
    const long slot = i / cpg;
    const int ch = int(i - slot * cpg);
    const float2 v = base[slot * C + ch];
    s += double(v.x);
    q += double(v.y);
  }
  __shared__ double rs[256], rq[256];
  rs[threadIdx.x] = s;
  rq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      rs[threadIdx.x] += rs[threadIdx.x + o];
      rq[threadIdx.x] += rq[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = double(d.z) * cpg;
    const double mean = rs[0] / n;
    double var = rq[0] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(size_t(l) * G + g) * 2] = float(mean);
    stats[(size_t(l) * G + g) * 2 + 1] = float(1.0 / sqrt(var + double(eps)));
  }
}

// SiLU as h + h * tanh(h), h = v / 2: one MUFU op per element instead of ex2 + rcp (+ the Newton
// steps of an IEEE division); tanh.approx is accurate to 2^-11, below the bf16 rounding of y.
// Chunks are walked from the END of the tensor (CTA 0 takes the last one): the statistics pass
// has just streamed x front to back, so its tail is what the 126 MB L2 still holds when x is larger
// than L2 (and the whole of x when it is not).
// st: (mean, rstd) pairs of the latent's groups, indexed by group - g_base: global memory (read through
// L1: every thread of a CTA asks for the same few pairs) in the three-launch path, a shared-memory copy
// of the slice's groups in the single-launch kernel (the statistics were written by other SMs during
// the same launch: fetched once per item through L2).
template <bool SILU, bool ST_SMEM>
__device__ __forceinline__ void gn_apply_item(const __nv_bfloat16* x, int ldx, int W, int RL, int cpg, int G,
                                              const float2* st_in, int g_base, const int* row_group,
                                              const float* a0, const float* b0, __nv_bfloat16* y, int ldy,
                                              int chunk, int ch0) {
  const int rl = threadIdx.x / W;
  const size_t row0 = size_t(chunk) * GN_CHUNK;
  const __nv_bfloat16* xb = x + row0 * ldx + ch0;
  __nv_bfloat16* yb = y + row0 * ldy + ch0;
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // first rows in flight while a, b are folded
    const int r = rl + k * RL;
    if (r < GN_CHUNK) v[k] = *reinterpret_cast<const uint4*>(xb + size_t(r) * ldx);
  }
  // (the x loads above are already in flight when the latent index is fetched)
  const float2* st = ST_SMEM ? st_in : st_in + size_t(row_group[row0]) * G;
  float a[8], b[8];  // (a0, b0 die here in the three-launch kernel: 64 registers without spills)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 mr = st[(ch0 + j) / cpg - g_base];
    a[j] = a0[j] * mr.y;
    b[j] = fmaf(-mr.x, a[j], b0[j]);
    if (SILU) {
      a[j] *= 0.5f;
      b[j] *= 0.5f;
    }
  }
  for (int r0 = rl;;) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k * RL;
      if (r < GN_CHUNK) {
        float f[8];
        unpack8s(v[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float h = fmaf(f[j], a[j], b[j]);
          f[j] = SILU ? fmaf(h, tanh_approx(h), h) : h;
        }
        *reinterpret_cast<uint4*>(yb + size_t(r) * ldy) = pack8s(f);
      }
    }
    r0 += 4 * RL;
    if (r0 >= GN_CHUNK) break;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + k * RL;
      if (r < GN_CHUNK) v[k] = *reinterpret_cast<const uint4*>(xb + size_t(r) * ldx);
    }
  }
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, 4) gn_apply_kernel(
    const __nv_bfloat16* x, int ldx, int W, int RL, int cpg, int G, int n_chunks,
    const float* stats, const int* row_group, const __nv_bfloat16* gamma,
    const __nv_bfloat16* beta, __nv_bfloat16* y, int ldy) {
  pdl_launch_dependents();
  const int col = threadIdx.x % W;
  const int ch0 = (blockIdx.y * W + col) * 8;  // first channel of this thread's vector
  float a[8], b[8];
  unpack8s(*reinterpret_cast<const uint4*>(gamma + ch0), a);  // parameters: not written by the
  unpack8s(*reinterpret_cast<const uint4*>(beta + ch0), b);   // preceding kernels
  pdl_wait();
  gn_apply_item<SILU, false>(x, ldx, W, RL, cpg, G, reinterpret_cast<const float2*>(stats), 0, row_group, a, b, y,
                             ldy, n_chunks - 1 - int(blockIdx.x), ch0);
}

// ---- single launch: statistics -> grid barrier -> finalize -> grid barrier -> apply, by a grid that
// is co-resident by construction (host: grid <= occupancy x SMs; every CTA runs
// griddepcontrol.launch_dependents first, so the next kernel cannot take a slot before all CTAs of
// this one are running). The barrier costs no atomics: CTA b stores the epoch into flag[b], CTA 0
// polls the flags (coalesced) and publishes the epoch in `gen`, everyone else polls `gen`
// (~1 us per barrier; 444 atomics on one word would serialise to ~6 us). bar[0] = gen, bar[1 + b] =
// flags; zero-initialised once by the caller, the epoch keeps counting across launches.
// The second read of x comes from L2 whenever the tensor fits (<= ~100 MB), walked in reverse order.
__device__ __forceinline__ void gn_grid_barrier(volatile unsigned int* bar, unsigned int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    bar[1 + blockIdx.x] = epoch;
  }
  if (blockIdx.x == 0) {
    for (int b = threadIdx.x; b < int(gridDim.x); b += blockDim.x)
      while (bar[1 + b] != epoch) {
      }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      bar[0] = epoch;
    }
  } else if (threadIdx.x == 0) {
    while (bar[0] != epoch) {
    }
  }
  __syncthreads();
  __threadfence();
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, 3) gn_fused_kernel(
    const __nv_bfloat16* x, int ldx, int W, int RL, int cpg, int G, int S, int n_chunks, int L, float eps,
    const int4* lat, const int* row_group, const __nv_bfloat16* gamma, const __nv_bfloat16* beta,
    __nv_bfloat16* y, int ldy, float* partial, float* stats, unsigned int* bar) {
  __shared__ float ch_sum[GN_THREADS * 8], ch_sq[GN_THREADS * 8];
  pdl_launch_dependents();
  const int n_items = n_chunks * S;
  pdl_wait();
  // last epoch of the previous launch on this workspace (complete and flushed: read after the wait);
  // CTA 0 publishes epoch0 + 1 only after every CTA has stored its flag, i.e. has read epoch0
  const unsigned int epoch0 = *reinterpret_cast<volatile unsigned int*>(bar);
  for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
    gn_stats_item(x, ldx, W, RL, cpg, G / S, n_chunks, partial, it / S, it % S, ch_sum, ch_sq);
    __syncthreads();
  }
  gn_grid_barrier(bar, epoch0 + 1);
  {
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;  // full warps only
    if (w < nw)
      for (int idx = blockIdx.x * nw + w; idx < L * G; idx += gridDim.x * nw)
        gn_finalize_one(partial, lat, G, cpg, n_chunks, eps, stats, idx, threadIdx.x & 31);
  }
  gn_grid_barrier(bar, epoch0 + 2);
  const int col = threadIdx.x % W;
  const int gps = G / S;
  __shared__ float2 st_s[GN_MAXG];
  int last_slice = -1;
  float a[8], b[8];
  for (int it = n_items - 1 - int(blockIdx.x); it >= 0; it -= gridDim.x) {
    const int chunk = it / S, slice = it % S;
    const int ch0 = (slice * W + col) * 8;
    if (slice != last_slice) {
      unpack8s(*reinterpret_cast<const uint4*>(gamma + ch0), a);
      unpack8s(*reinterpret_cast<const uint4*>(beta + ch0), b);
      last_slice = slice;
    }
    __syncthreads();  // the previous item's readers of st_s are done
    if (int(threadIdx.x) < gps)
      st_s[threadIdx.x] = __ldcg(reinterpret_cast<const float2*>(stats) +
                                 size_t(row_group[size_t(chunk) * GN_CHUNK]) * G + slice * gps + threadIdx.x);
    __syncthreads();
    gn_apply_item<SILU, true>(x, ldx, W, RL, cpg, G, st_s, slice * gps, row_group, a, b, y, ldy, chunk, ch0);
  }
}

// ------------------------------------------------------------------ latent pack for conv_in
// NCHW latent [C, H, W] bf16 -> im2col rows of the 3x3/pad-1 conv: out[row0 + y*W + x,
// c*9 + ky*3 + kx] = lat[c, y+ky-1, x+kx-1] (0 outside), columns >= 9C are zero (K padded to
// ldo). conv_in then is one GEMM with the [Cout, C*9] weight matrix of diffusers as is.
// Replaces split_sample's halo windows + F.conv2d(padding=0) (modules/unet.py:104-184,344).
// desc: [n][4] = {row offset, H, W, 0}
__global__ void pack_im2col3x3_kernel(const unsigned long long* lat_ptr, const int4* desc, int C,
                                      __nv_bfloat16* out, int ldo) {
  const int l = blockIdx.y;
  const int4 d = desc[l];
  const int H = d.y, W = d.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * W * ldo) return;
  const int pix = idx / ldo, k = idx % ldo;
  float v = 0.f;
  if (k < C * 9) {
    const int c = k / 9, ky = (k % 9) / 3, kx = k % 3;
    const int y = pix / W + ky - 1, x = pix % W + kx - 1;
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(lat_ptr[l])[(size_t(c) * H + y) * W + x]);
  }
  out[size_t(d.x + pix) * ldo + k] = __float2bfloat16(v);
}

// NHWC rows [T, ldx] (first C columns) -> NCHW latents (scatter of the noise prediction;
// replaces concat_sample, modules/unet.py:187-202).
__global__ void scatter_nchw_kernel(const __nv_bfloat16* x, int ldx, const int4* desc, int C,
                                    const unsigned long long* out_ptr) {
  const int l = blockIdx.y;
  const int4 d = desc[l];
  const int HW = d.y * d.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * HW) return;
  const int c = idx / HW, pix = idx % HW;
  reinterpret_cast<__nv_bfloat16*>(out_ptr[l])[idx] = x[size_t(d.x + pix) * ldx + c];
}

// Nearest-neighbour 2x upsample on the packed NHWC layout (F.interpolate(scale_factor=2,
// mode="nearest"), modules/resnet.py:316). in_desc/out_desc: [n][4] = {row offset, H, W, 0}.
__global__ void upsample2x_kernel(const __nv_bfloat16* x, int ldx, const int4* in_desc,
                                  const int4* out_desc, int C, __nv_bfloat16* y, int ldy) {
  const int l = blockIdx.y;
  const int4 di = in_desc[l], dout = out_desc[l];
  const int nvec = C >> 3;
  const long idx = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (idx >= long(dout.y) * dout.z * nvec) return;
  const int pix = int(idx / nvec), vc = int(idx % nvec);
  const int oy = pix / dout.z, ox = pix % dout.z;
  const uint4 v = *reinterpret_cast<const uint4*>(
      x + size_t(di.x + (oy >> 1) * di.z + (ox >> 1)) * ldx + vc * 8);
  *reinterpret_cast<uint4*>(y + size_t(dout.x + pix) * ldy + vc * 8) = v;
}

// dst[:, dst_col : dst_col + cols] = src[:, src_col : src_col + cols]  (skip-connection concat)
__global__ void copy_cols_kernel(const __nv_bfloat16* src, int lds, __nv_bfloat16* dst, int ldd,
                                 long T, int cols) {
  pdl_launch_dependents();
  pdl_wait();
  const int nvec = cols >> 3;
  const long idx = blockIdx.x * long(blockDim.x) + threadIdx.x;
  if (idx >= T * nvec) return;
  const long row = idx / nvec;
  const int vc = int(idx - row * nvec);
  *reinterpret_cast<uint4*>(dst + row * ldd + vc * 8) =
      *reinterpret_cast<const uint4*>(src + row * lds + vc * 8);
}

// ------------------------------------------------------------------ reference patch format
// split_sample (modules/unet.py:104-184): NCHW latents -> [P, C, ps+2, ps+2] haloed windows
// (zero outside the image), patches of a latent in row-major (h, w) order.
// pdesc: [P][4] = {latent, patch row h, patch col w, 0}; ldesc: [n][4] = {0, H, W, 0}
__global__ void split_patches_kernel(const unsigned long long* lat_ptr, const int4* ldesc,
                                     const int4* pdesc, int C, int ps, __nv_bfloat16* out) {
  const int p = blockIdx.y;
  const int4 pd = pdesc[p];
  const int4 ld = ldesc[pd.x];
  const int H = ld.y, W = ld.z, pw = ps + 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * pw * pw) return;
  const int c = idx / (pw * pw), yy = (idx / pw) % pw, xx = idx % pw;
  const int y = pd.y * ps + yy - 1, x = pd.z * ps + xx - 1;
  __nv_bfloat16 v = __float2bfloat16(0.f);
  if (y >= 0 && y < H && x >= 0 && x < W)
    v = reinterpret_cast<const __nv_bfloat16*>(lat_ptr[pd.x])[(size_t(c) * H + y) * W + x];
  out[size_t(p) * C * pw * pw + idx] = v;
}

// concat_sample (modules/unet.py:187-202): [P, C, ps, ps] -> NCHW latents.
__global__ void concat_patches_kernel(const __nv_bfloat16* patches, const int4* ldesc,
                                      const int4* pdesc, int C, int ps,
                                      const unsigned long long* out_ptr) {
  const int p = blockIdx.y;
  const int4 pd = pdesc[p];
  const int4 ld = ldesc[pd.x];
  const int H = ld.y, W = ld.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * ps * ps) return;
  const int c = idx / (ps * ps), yy = (idx / ps) % ps, xx = idx % ps;
  reinterpret_cast<__nv_bfloat16*>(out_ptr[pd.x])[(size_t(c) * H + pd.y * ps + yy) * W + pd.z * ps + xx] =
      patches[size_t(p) * C * ps * ps + idx];
}

}  // namespace b200

using namespace b200;
typedef __nv_bfloat16 bf16;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

constexpr int GN_MAX_GRID = 1023;  // CTAs of the single-launch kernel (flags of its grid barrier)
extern "C" long long b200_groupnorm_workspace_bytes(long long total_rows, int n_latents) {
  return (total_rows / GN_CHUNK) * GN_MAXG * 2 * 4 + (long long)n_latents * GN_MAXG * 2 * 4 +
         (1 + GN_MAX_GRID) * 4;
}

// co-resident CTAs of gn_fused_kernel<SILU> with `threads` threads (queried once per variant)
template <bool SILU>
static int gn_fused_capacity(int threads) {
  static int per_sm[GN_THREADS / 32 + 1] = {0};
  int& v = per_sm[(threads + 31) / 32];
  if (v == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gn_fused_kernel<SILU>, threads, 0) != cudaSuccess || n <= 0)
      n = -1;
    v = n;
  }
  if (v <= 0) return 0;
  // two CTAs per SM: enough loads in flight for the HBM-bound passes (296 x 256 x 64 B), and half of what
  // fits, so that two such kernels on different streams are always co-resident together (a third
  // concurrent one could starve the barriers: one GroupNorm stream per workspace, two per device)
  const int per = v < 2 ? v : 2;
  const int cap = per * device_sm_count();
  return cap < GN_MAX_GRID ? cap : GN_MAX_GRID;
}

extern "C" int b200_groupnorm_nhwc_bf16(const void* x, int ldx, long long T, int C, int groups,
                                        float eps, const void* gamma, const void* beta,
                                        const int32_t* row_group, const int32_t* lat_chunks,
                                        int n_latents, int silu, void* y, int ldy, void* workspace,
                                        void* stream) {
  if (!x || !y || !gamma || !beta || !row_group || !lat_chunks || !workspace || T <= 0 ||
      (T % GN_CHUNK) || groups <= 0 || groups > GN_MAXG || (C % groups) || (C & 7) || C > GN_MAXC || (ldx & 7) ||
      (ldy & 7))
    return B200_ERR_INVALID;
  const int cpg = C / groups;
  const int chunks = int(T / GN_CHUNK);
  const GnMap m = gn_map(C, groups);
  if (m.RL < 1) return B200_ERR_INVALID;
  float* partial = static_cast<float*>(workspace);
  float* stats = partial + size_t(chunks) * GN_MAXG * 2;
  // SDUSS_B200_GN_FUSED=1: one launch (statistics, finalize and apply around two grid barriers). Same
  // arithmetic per chunk / per (latent, group), bit-identical results -- and measured SLOWER than the
  // three launches (profiles/r02_gn_single_launch.txt: a barrier costs what a kernel boundary costs,
  // and 2 CTAs per SM keep fewer loads in flight than the 4-5 of the separate kernels), so it is off.
  {
    const char* vf = getenv("SDUSS_B200_GN_FUSED");
    const int cap = !(vf && vf[0] == '1') ? 0 : silu ? gn_fused_capacity<true>(m.threads) : gn_fused_capacity<false>(m.threads);
    if (cap > 0) {
      const long items = long(chunks) * m.S;
      const int grid = int(items < cap ? items : cap);
      unsigned int* bar = reinterpret_cast<unsigned int*>(stats + size_t(n_latents) * GN_MAXG * 2);
      auto fk = silu ? gn_fused_kernel<true> : gn_fused_kernel<false>;
      return launch_pdl(fk, dim3(grid), dim3(m.threads), 0, ST(stream), static_cast<const bf16*>(x), ldx, m.W,
                        m.RL, cpg, groups, m.S, chunks, n_latents, eps, reinterpret_cast<const int4*>(lat_chunks),
                        row_group, static_cast<const bf16*>(gamma), static_cast<const bf16*>(beta),
                        static_cast<bf16*>(y), ldy, partial, stats, bar);
    }
  }
  int rc = launch_pdl(gn_stats_kernel, dim3(chunks, m.S), dim3(m.threads), 0, ST(stream),
                      static_cast<const bf16*>(x), ldx, m.W, m.RL, cpg, groups / m.S, chunks,
                      partial);
  if (rc) return rc;
  rc = launch_pdl(gn_finalize_kernel, dim3((n_latents * groups + 3) / 4), dim3(128), 0, ST(stream),
                  static_cast<const float*>(partial), reinterpret_cast<const int4*>(lat_chunks),
                  n_latents, groups, cpg, chunks, eps, stats);
  if (rc) return rc;
  auto kern = silu ? gn_apply_kernel<true> : gn_apply_kernel<false>;
  return launch_pdl(kern, dim3(chunks, m.S), dim3(m.threads), 0, ST(stream),
                    static_cast<const bf16*>(x), ldx, m.W, m.RL, cpg, groups, chunks,
                    static_cast<const float*>(stats), row_group, static_cast<const bf16*>(gamma),
                    static_cast<const bf16*>(beta), static_cast<bf16*>(y), ldy);
}

// GroupNorm whose statistics the producing convolution already computed (b200_conv3x3_bf16 with
// ep->stats_out): finalize over the latent's tile partials + the apply pass -- x is read ONCE.
extern "C" int b200_groupnorm_from_conv_stats(const void* x, int ldx, long long T, int C, int groups,
                                              float eps, const void* gamma, const void* beta,
                                              const int32_t* row_group, const float* conv_stats,
                                              const int32_t* lat_tiles, int n_latents, int silu,
                                              void* y, int ldy, void* workspace, void* stream) {
  if (!x || !y || !gamma || !beta || !row_group || !conv_stats || !lat_tiles || !workspace || T <= 0 ||
      (T % GN_CHUNK) || groups <= 0 || groups > GN_MAXG || (C % groups) || (C & 7) || C > GN_MAXC ||
      (ldx & 7) || (ldy & 7) || n_latents <= 0)
    return B200_ERR_INVALID;
  const int cpg = C / groups;
  const int chunks = int(T / GN_CHUNK);
  const GnMap m = gn_map(C, groups);
  if (m.RL < 1) return B200_ERR_INVALID;
  float* stats = static_cast<float*>(workspace) + size_t(chunks) * GN_MAXG * 2;  // same slot as the 3-launch path
  int rc = launch_pdl(gn_finalize_tiles_kernel, dim3(n_latents * groups), dim3(256), 0, ST(stream),
                      reinterpret_cast<const float2*>(conv_stats), reinterpret_cast<const int4*>(lat_tiles),
                      groups, cpg, C, eps, stats);
  if (rc) return rc;
  auto kern = silu ? gn_apply_kernel<true> : gn_apply_kernel<false>;
  return launch_pdl(kern, dim3(chunks, m.S), dim3(m.threads), 0, ST(stream),
                    static_cast<const bf16*>(x), ldx, m.W, m.RL, cpg, groups, chunks,
                    static_cast<const float*>(stats), row_group, static_cast<const bf16*>(gamma),
                    static_cast<const bf16*>(beta), static_cast<bf16*>(y), ldy);
}

extern "C" int b200_pack_im2col3x3(const uint64_t* lat_ptr, const int32_t* desc, int n_latents,
                                   int max_pixels, int C, void* out, int ldo, void* stream) {
  if (!lat_ptr || !desc || !out || n_latents <= 0 || max_pixels <= 0 || ldo < 9 * C || (ldo & 7))
    return B200_ERR_INVALID;
  dim3 grid((unsigned(max_pixels) * ldo + 255) / 256, n_latents);
  pack_im2col3x3_kernel<<<grid, 256, 0, ST(stream)>>>(
      reinterpret_cast<const unsigned long long*>(lat_ptr), reinterpret_cast<const int4*>(desc), C,
      static_cast<bf16*>(out), ldo);
  return launch_status();
}

extern "C" int b200_scatter_nchw(const void* x, int ldx, const int32_t* desc, int n_latents,
                                 int max_pixels, int C, const uint64_t* out_ptr, void* stream) {
  if (!x || !desc || !out_ptr || n_latents <= 0 || max_pixels <= 0) return B200_ERR_INVALID;
  dim3 grid((unsigned(max_pixels) * C + 255) / 256, n_latents);
  scatter_nchw_kernel<<<grid, 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(x), ldx, reinterpret_cast<const int4*>(desc), C,
      reinterpret_cast<const unsigned long long*>(out_ptr));
  return launch_status();
}

extern "C" int b200_upsample2x_nhwc(const void* x, int ldx, const int32_t* in_desc,
                                    const int32_t* out_desc, int n_latents, int max_out_pixels,
                                    int C, void* y, int ldy, void* stream) {
  if (!x || !y || !in_desc || !out_desc || n_latents <= 0 || (C & 7) || (ldx & 7) || (ldy & 7))
    return B200_ERR_INVALID;
  dim3 grid(unsigned((long(max_out_pixels) * (C >> 3) + 255) / 256), n_latents);
  upsample2x_kernel<<<grid, 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(x), ldx, reinterpret_cast<const int4*>(in_desc),
      reinterpret_cast<const int4*>(out_desc), C, static_cast<bf16*>(y), ldy);
  return launch_status();
}

extern "C" int b200_copy_cols_bf16(const void* src, int lds, void* dst, int ldd, long long T,
                                   int cols, void* stream) {
  if (!src || !dst || T <= 0 || cols <= 0 || (cols & 7) || (lds & 7) || (ldd & 7))
    return B200_ERR_INVALID;
  const long total = T * (cols >> 3);
  return launch_pdl(copy_cols_kernel, dim3(unsigned((total + 255) / 256)), dim3(256), 0, ST(stream),
                    static_cast<const bf16*>(src), lds, static_cast<bf16*>(dst), ldd, long(T), cols);
}

extern "C" int b200_split_patches(const uint64_t* lat_ptr, const int32_t* ldesc,
                                  const int32_t* pdesc, int n_patches, int C, int ps, void* out,
                                  void* stream) {
  if (!lat_ptr || !ldesc || !pdesc || !out || n_patches <= 0) return B200_ERR_INVALID;
  dim3 grid((unsigned(C) * (ps + 2) * (ps + 2) + 255) / 256, n_patches);
  split_patches_kernel<<<grid, 256, 0, ST(stream)>>>(
      reinterpret_cast<const unsigned long long*>(lat_ptr), reinterpret_cast<const int4*>(ldesc),
      reinterpret_cast<const int4*>(pdesc), C, ps, static_cast<bf16*>(out));
  return launch_status();
}

extern "C" int b200_concat_patches(const void* patches, const int32_t* ldesc, const int32_t* pdesc,
                                   int n_patches, int C, int ps, const uint64_t* out_ptr,
                                   void* stream) {
  if (!patches || !ldesc || !pdesc || !out_ptr || n_patches <= 0) return B200_ERR_INVALID;
  dim3 grid((unsigned(C) * ps * ps + 255) / 256, n_patches);
  concat_patches_kernel<<<grid, 256, 0, ST(stream)>>>(
      static_cast<const bf16*>(patches), reinterpret_cast<const int4*>(ldesc),
      reinterpret_cast<const int4*>(pdesc), C, ps,
      reinterpret_cast<const unsigned long long*>(out_ptr));
  return launch_status();
}
