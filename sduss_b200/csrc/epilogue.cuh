// Fused epilogues shared by the tcgen05 GEMM and implicit-GEMM convolution kernels.
// One epilogue thread owns one accumulator row (TMEM lane) and walks it in chunks of 64
// fp32 columns, so per-row reductions (q/k RMSNorm over a 64-wide head) are thread-local.
#pragma once
#include "ptx.cuh"

namespace b200 {

enum EpiMode : int {
  EPI_BIAS = 0,        // C = acc + bias
  EPI_GELU_TANH = 1,   // C = gelu_tanh(acc + bias)                       (SD3 FeedForward)
  EPI_GATE_RESID = 2,  // C = resid + gate[group[row]] * (acc + bias)     (gate null -> 1)
  EPI_QK_RMSNORM = 3,  // per 64-col head: q,k column ranges RMS-normalised, v passthrough
  EPI_GEGLU = 4,       // weights interleaved [32 h | 32 g]: C[:, n/2] = h * gelu_erf(g)
  EPI_ROWVEC = 5,      // C = acc + bias + rowvec[group[row]]            (resnet conv1 + temb)
};

struct EpiArgs {
  void* C;              // bf16 (or fp32 when out_fp32) [M, ldc]
  int ldc;
  int out_fp32;
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* resid;  // [M, ldr] or null (may alias C)
  int ldr;
  const __nv_bfloat16* gate;   // [G, ldg] or null
  int ldg;
  const int* row_group;        // [M] group (request/latent) id of each row, or null
  const __nv_bfloat16* rowvec; // [G, ldv]
  int ldv;
  const __nv_bfloat16* rms_wq; // [64]
  const __nv_bfloat16* rms_wk; // [64]
  int rms_q_cols;              // columns [0, q_cols) are q heads
  int rms_k_cols;              // columns [q_cols, q_cols + k_cols) are k heads
  float rms_eps;
  float q_scale;               // multiplied into normalised q (softmax scale * log2 e)
  int act;                     // activation of EPI_GELU_TANH / gate of EPI_GEGLU: 0 = the mode's
                               // default (tanh-GELU / erf-GELU), 1 tanh-GELU, 2 erf-GELU, 3 quick-GELU
  const int* row_mask;         // patch cache: M tiles of chunk (m0 >> row_mask_shift) with mask 0 are
  int row_mask_shift;          // skipped entirely (no loads, no MMA, C untouched); null = all tiles
  float2* stats_out;           // conv only: per (M tile, row half, channel) (sum, sum of squares) of the
                               // STORED bf16 outputs, for the GroupNorm that follows; null = off
  // LayerNorm folded into this GEMM (A is the UN-normalised activation, W already carries gamma):
  //   LN(x) W^T = rstd_row * (x (gamma o W)^T - mean_row * colsum), + beta W^T inside `bias`
  const float2* ln_stats;      // [M] (mean, rstd) per row, or null
  const float* ln_colsum;      // [N] sum_k gamma_k W[n, k] (of the bf16 weights actually multiplied)
  // ... with the row statistics taken from the per-chunk partial sums the PRODUCER of A left behind:
  const float2* ln_rowpart;    // [ln_nparts][M] (sum, sum of squares) over each 64-column chunk of A's row
  int ln_nparts;               // = K / 64 of this GEMM   (chunk-major: a thread per row reads / writes
  int part_ld;                 // = M                      coalesced across the 128 rows of a tile)
  float ln_eps;
  float2* rowpart_out;         // producer side: [ceil(N / 64)][M] partial sums of THIS GEMM's output rows
  int w_static;                // W is never written on the stream (weights): may be loaded before pdl_wait
};

// (mean, rstd) of row `row` of A for the folded LayerNorm, from whichever source the caller gave
__device__ __forceinline__ float2 ln_row_stats(const EpiArgs& e, int row, bool row_ok, int K) {
  if (e.ln_colsum == nullptr || !row_ok) return make_float2(0.f, 0.f);
  if (e.ln_rowpart == nullptr) return e.ln_stats[row];
  const float2* p = e.ln_rowpart + row;
  float s = 0.f, q = 0.f;
  // all loads of a batch are in flight together (one L2 round trip per 20 chunks = K 1280), then
  // summed in a fixed order
  for (int i0 = 0; i0 < e.ln_nparts; i0 += 20) {
    float2 v[20];
#pragma unroll
    for (int j = 0; j < 20; ++j)
      v[j] = i0 + j < e.ln_nparts ? p[size_t(i0 + j) * e.part_ld] : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      s += v[j].x;
      q += v[j].y;
    }
  }
  const float mean = s / float(K);
  const float var = fmaxf(q / float(K) - mean * mean, 0.f);
  return make_float2(mean, rsqrtf(var + e.ln_eps));
}

// acc <- rstd * (acc - mean * colsum): turns x (gamma o W)^T into LN_noaffine(x) (gamma o W)^T
// cs: the 64 column sums of this chunk (global, or staged in shared memory by the GEMM epilogue)
__device__ __forceinline__ void ln_fold64(const float* cs_ptr, float* acc, float2 mr, int ncols) {
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    if (j < ncols) {
      const float4 cs = *reinterpret_cast<const float4*>(cs_ptr + j);
      acc[j] = mr.y * fmaf(-mr.x, cs.x, acc[j]);
      acc[j + 1] = mr.y * fmaf(-mr.x, cs.y, acc[j + 1]);
      acc[j + 2] = mr.y * fmaf(-mr.x, cs.z, acc[j + 2]);
      acc[j + 3] = mr.y * fmaf(-mr.x, cs.w, acc[j + 3]);
    }
  }
}

// true when the M tile starting at row m0 belongs to a clean patch (kept as it is)
__device__ __forceinline__ bool tile_skipped(const EpiArgs& e, int m0, int M) {
  return e.row_mask != nullptr && m0 < M && e.row_mask[m0 >> e.row_mask_shift] == 0;
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * (x + k1 * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.f + t);
}
// x * Phi(x), Phi through erf(|x| / sqrt 2) = 1 - (a1 t + ... + a5 t^5) exp(-x^2 / 2), t = 1 / (1 + p |x| / sqrt 2)
// (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7 -- five orders below the bf16 rounding of the result).
// Branch-free: 2 MUFU (rcp, ex2) + 11 FMA-pipe ops. The libdevice erff() takes ~40 instructions with a
// divergent branch at |x| ~ 0.9, which made the GEGLU epilogue (128 x 128 activations per tile on 4
// warps) 1.5x longer than the tile's main loop (SDXL ff1: 70 us per call against 47 us plain).
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float z = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(z, 0.3275911f * 0.7071067811865476f, 1.f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * (-0.5f * 1.4426950408889634f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  p *= t;
  const float erf_abs = fmaf(-p, e, 1.f);                   // erf(|x| / sqrt 2) in [0, 1]
  const float phi = fmaf(copysignf(0.5f, x), erf_abs, 0.5f);
  return x * phi;
}

__device__ __forceinline__ float quick_gelu_f(float x) {  // x * sigmoid(1.702 x)   (CLIP-L)
  return x / (1.f + __expf(-1.702f * x));
}
// Runtime-selected activation (uniform per launch): the text encoders of the prepare stage use
// erf-GELU (CLIP-G), quick-GELU (CLIP-L) and a tanh-GELU gate (T5 v1.1) on the same templates.
// The selector is tested ONCE per chunk, outside the element loops: the hot epilogues (SD3 FF1: tanh,
// SDXL GEGLU: erf) keep a single straight-line activation loop.
template <int N>
__device__ __forceinline__ void act_inplace(float* v, int act, int dflt) {
  const int a = act == 0 ? dflt : act;
  if (a == 1) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = gelu_tanh_f(v[j]);
  } else if (a == 2) {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = gelu_erf_f(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) v[j] = quick_gelu_f(v[j]);
  }
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float* out) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(v.x, out[0], out[1]);
  unpack_bf16x2(v.y, out[2], out[3]);
  unpack_bf16x2(v.z, out[4], out[5]);
  unpack_bf16x2(v.w, out[6], out[7]);
}

// acc: 64 fp32 accumulator columns of one row; n0 = first global column of this chunk.
// Writes the finished chunk to global memory. N is the logical GEMM N (pre-GEGLU halving).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk64(const EpiArgs& e, float* acc, int row, int n0,
                                                 int M, int N) {
  if (row >= M || n0 >= N) return;
  const int ncols = min(64, N - n0);  // multiple of 8 by contract
  if (e.ln_colsum != nullptr) ln_fold64(e.ln_colsum + n0, acc, ln_row_stats(e, row, true, e.ln_nparts * 64), ncols);
  if (e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      if (j < ncols) {
        float b[8];
        load8_bf16(e.bias + n0 + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] += b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GELU_TANH) {
    act_inplace<64>(acc, e.act, 1);
  }
  if constexpr (EPI == EPI_ROWVEC) {
    const int g = e.row_group[row];
    const __nv_bfloat16* v = e.rowvec + size_t(g) * e.ldv + n0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      if (j < ncols) {
        float b[8];
        load8_bf16(v + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] += b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GATE_RESID) {
    if (e.gate != nullptr) {
      const int g = e.row_group[row];
      const __nv_bfloat16* gp = e.gate + size_t(g) * e.ldg + n0;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        if (j < ncols) {
          float b[8];
          load8_bf16(gp + j, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j + i] *= b[i];
        }
      }
    }
    if (e.resid != nullptr) {
      const __nv_bfloat16* rp = e.resid + size_t(row) * e.ldr + n0;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        if (j < ncols) {
          float b[8];
          load8_bf16(rp + j, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j + i] += b[i];
        }
      }
    }
  }
  if constexpr (EPI == EPI_QK_RMSNORM) {
    // n0 is 64-aligned, so a chunk is exactly one head.
    const bool is_q = n0 < e.rms_q_cols;
    const bool is_k = !is_q && n0 < e.rms_q_cols + e.rms_k_cols;
    if (is_q || is_k) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) ss += acc[j] * acc[j];
      float r = rsqrtf(ss * (1.f / 64.f) + e.rms_eps);
      if (is_q) r *= e.q_scale;
      const __nv_bfloat16* w = is_q ? e.rms_wq : e.rms_wk;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        float b[8];
        load8_bf16(w + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] = acc[j + i] * r * b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GEGLU) {
    // chunk = [32 hidden | 32 gate] -> 32 output columns at n0/2
    act_inplace<32>(acc + 32, e.act, 2);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = acc[j] * acc[32 + j];
    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(e.C) + size_t(row) * e.ldc + (n0 >> 1);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 v;
      v.x = pack_bf16x2(acc[j], acc[j + 1]);
      v.y = pack_bf16x2(acc[j + 2], acc[j + 3]);
      v.z = pack_bf16x2(acc[j + 4], acc[j + 5]);
      v.w = pack_bf16x2(acc[j + 6], acc[j + 7]);
      *reinterpret_cast<uint4*>(cp + j) = v;
    }
    return;
  }
  if (e.out_fp32) {
    float* cp = reinterpret_cast<float*>(e.C) + size_t(row) * e.ldc + n0;
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      if (j < ncols)
        *reinterpret_cast<float4*>(cp + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
  } else {
    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(e.C) + size_t(row) * e.ldc + n0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      if (j < ncols) {
        uint4 v;
        v.x = pack_bf16x2(acc[j], acc[j + 1]);
        v.y = pack_bf16x2(acc[j + 2], acc[j + 3]);
        v.z = pack_bf16x2(acc[j + 4], acc[j + 5]);
        v.w = pack_bf16x2(acc[j + 6], acc[j + 7]);
        *reinterpret_cast<uint4*>(cp + j) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Shared-memory staged epilogue: the finished 128 x 64 bf16 chunk is written to a 128B-swizzled
// staging tile and leaves through ONE TMA bulk store (residual tiles arrive the same way through
// TMA loads). Direct register->global stores make every warp instruction touch 32 different
// 128-byte lines (one row per lane); for K <= 1536 GEMMs with a residual that LSU traffic, not
// the tensor pipe, set the pace (752 vs 1190 TFLOP/s measured on [14848,1536]x[1536,1536]).
// ------------------------------------------------------------------------------------------
constexpr int EPI_STAGE_BYTES = 128 * 128;  // 128 rows x 64 bf16

__device__ __forceinline__ uint32_t sw128_off(int r, int k) {  // row r, 16-byte chunk k of the row
  return uint32_t(r) * 128u + (uint32_t(k ^ (r & 7)) << 4);
}

// Applies the fused epilogue to one row chunk (acc[64], fp32, in place). resid_stage: smem tile
// holding the residual chunk (swizzled) or nullptr. Returns the number of valid output columns
// (32 for GEGLU, else 64).
template <int EPI>
__device__ __forceinline__ void epilogue_math64(const EpiArgs& e, float* acc, int row, bool row_ok,
                                                int n0, int N, const uint8_t* resid_stage, int r,
                                                float2 ln_mr = make_float2(0.f, 0.f),
                                                const float* ln_cs = nullptr) {
  const int ncols = min(64, N - n0);
  if (ln_cs != nullptr && e.ln_colsum != nullptr) {
    // folded LayerNorm + bias in two FMAs per element: acc * rstd + (bias - mean * rstd * colsum);
    // ln_cs[0..63] = column sums, ln_cs[256..319] = the bias in fp32 (both staged per tile)
    const float nm = -ln_mr.x * ln_mr.y;
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      const float4 cs = *reinterpret_cast<const float4*>(ln_cs + j);
      const float4 bs = *reinterpret_cast<const float4*>(ln_cs + 256 + j);
      acc[j] = fmaf(acc[j], ln_mr.y, fmaf(nm, cs.x, bs.x));
      acc[j + 1] = fmaf(acc[j + 1], ln_mr.y, fmaf(nm, cs.y, bs.y));
      acc[j + 2] = fmaf(acc[j + 2], ln_mr.y, fmaf(nm, cs.z, bs.z));
      acc[j + 3] = fmaf(acc[j + 3], ln_mr.y, fmaf(nm, cs.w, bs.w));
    }
  } else if (ln_cs != nullptr) {
    // bias staged in shared memory as fp32 (one add per element; the bf16 global path below costs
    // a shift / mask per element on top, in epilogues that are issue-bound)
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      const float4 bs = *reinterpret_cast<const float4*>(ln_cs + 256 + j);
      acc[j] += bs.x;
      acc[j + 1] += bs.y;
      acc[j + 2] += bs.z;
      acc[j + 3] += bs.w;
    }
  } else if (e.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      if (j < ncols) {
        float b[8];
        load8_bf16(e.bias + n0 + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] += b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GELU_TANH) {
    act_inplace<64>(acc, e.act, 1);
  }
  if constexpr (EPI == EPI_ROWVEC) {
    const int g = row_ok ? e.row_group[row] : 0;
    const __nv_bfloat16* v = e.rowvec + size_t(g) * e.ldv + n0;
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
      if (j < ncols) {
        float b[8];
        load8_bf16(v + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] += b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GATE_RESID) {
    if (e.gate != nullptr) {
      const int g = row_ok ? e.row_group[row] : 0;
      const __nv_bfloat16* gp = e.gate + size_t(g) * e.ldg + n0;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        if (j < ncols) {
          float b[8];
          load8_bf16(gp + j, b);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j + i] *= b[i];
        }
      }
    }
    if (resid_stage != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float b[8];
        load8_bf16(reinterpret_cast<const __nv_bfloat16*>(resid_stage + sw128_off(r, k)), b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k * 8 + i] += b[i];
      }
    }
  }
  if constexpr (EPI == EPI_QK_RMSNORM) {
    const bool is_q = n0 < e.rms_q_cols;
    const bool is_k = !is_q && n0 < e.rms_q_cols + e.rms_k_cols;
    if (is_q || is_k) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) ss += acc[j] * acc[j];
      float rr = rsqrtf(ss * (1.f / 64.f) + e.rms_eps);
      if (is_q) rr *= e.q_scale;
      const __nv_bfloat16* w = is_q ? e.rms_wq : e.rms_wk;
#pragma unroll
      for (int j = 0; j < 64; j += 8) {
        float b[8];
        load8_bf16(w + j, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j + i] = acc[j + i] * rr * b[i];
      }
    }
  }
  if constexpr (EPI == EPI_GEGLU) {
    act_inplace<32>(acc + 32, e.act, 2);
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = acc[j] * acc[32 + j];
  } else {
    if (e.rowpart_out != nullptr && row_ok) {
      // partial row statistics of the output for a LayerNorm folded into the NEXT GEMM
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        if (j < ncols) {
          s0 += acc[j];
          s1 += acc[j + 1];
          q0 = fmaf(acc[j], acc[j], q0);
          q1 = fmaf(acc[j + 1], acc[j + 1], q1);
        }
      }
      e.rowpart_out[size_t(n0 >> 6) * e.part_ld + row] = make_float2(s0 + s1, q0 + q1);
    }
  }
}

// Writes the row chunk to the staging tile as bf16 (128B-swizzled rows of 64 values; GEGLU:
// plain rows of 32 values).
template <int EPI>
__device__ __forceinline__ void epilogue_stage64(uint8_t* stage, const float* acc, int r) {
  if constexpr (EPI == EPI_GEGLU) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 v;
      v.x = pack_bf16x2(acc[8 * k], acc[8 * k + 1]);
      v.y = pack_bf16x2(acc[8 * k + 2], acc[8 * k + 3]);
      v.z = pack_bf16x2(acc[8 * k + 4], acc[8 * k + 5]);
      v.w = pack_bf16x2(acc[8 * k + 6], acc[8 * k + 7]);
      *reinterpret_cast<uint4*>(stage + r * 64 + k * 16) = v;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      uint4 v;
      v.x = pack_bf16x2(acc[8 * k], acc[8 * k + 1]);
      v.y = pack_bf16x2(acc[8 * k + 2], acc[8 * k + 3]);
      v.z = pack_bf16x2(acc[8 * k + 4], acc[8 * k + 5]);
      v.w = pack_bf16x2(acc[8 * k + 6], acc[8 * k + 7]);
      *reinterpret_cast<uint4*>(stage + sw128_off(r, k)) = v;
    }
  }
}

}  // namespace b200
