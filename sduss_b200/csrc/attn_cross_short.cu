// SDXL cross attention: queries = the image tokens of a latent (thousands of rows), keys / values = its
// 77 text tokens. Replaces the per-patch cross attention of the reference (sduss/model_executor/modules/
// attention.py:59-110: to_q on the patch, xformers.memory_efficient_attention against the text K / V).
//
// Why its own kernel. One launch is ~1 GFLOP (SDXL level 2, config-1: 2560 rows x 20 heads x 77 keys):
// nothing a tensor pipe or HBM could bound, only latency. In the persistent tcgen05 kernel
// (attn_sm100.cu) such a launch costs 17.5-20.5 us (profiles/r02_launches_sdxl_step.txt: 70 launches,
// 1.3 ms of the 21 ms step) -- TMEM allocation, the barrier network, two dependent table loads, the
// TMA round trips and, with 160 work units on 148 SMs, a second round for 12 CTAs, all for two 64-key
// steps. Here a CTA is 128 query rows of one (latent, head): eight warps of 16 rows, the whole K / V of
// the head (80 x 64 bf16 each) in shared memory, S = Q K^T and O = P V as register-level mma.sync
// m16n8k16 tiles (80 + 80 per warp), the softmax over all keys at once in registers (no online
// rescaling, no running state). 41 KB of shared memory, 256 threads and 80 registers: three CTAs per SM, the
// whole launch (400 CTAs at level 2) is resident at once and finishes in one latency chain:
// load -> 160 MMAs -> store. The legacy warp-level MMA is the right tool at this size; everything with
// a tensor-pipe-sized problem stays on tcgen05.
#include "../../include/sduss_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

constexpr int XS_ROWS = 128;    // query rows per CTA
constexpr int XS_THREADS = 256; // 8 warps x 16 rows
constexpr int XS_KV = 80;       // keys of a sequence, at most (5 k-steps of 16)
constexpr int XS_PITCH = 72;    // bf16 per shared-memory row: 64 + 8 keeps ldmatrix conflict-free

struct XsArgs {
  const __nv_bfloat16* q; int ldq, q_col;
  const __nv_bfloat16* k; int ldk, k_col;
  const __nv_bfloat16* v; int ldv, v_col;
  __nv_bfloat16* out; int ldo, o_col;
  const int* seq_table;  // [n_seq][8] as for attn_fwd_kernel: Q = segment A, K / V = segment B
  float scale_log2;
  const int* q_mask; int q_mask_shift;  // patch cache: 128-row tiles of clean patches are skipped
};

__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(p)));
}
// 16 bytes global -> shared without passing through registers; !valid: nothing is read, zeros are written
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(XS_THREADS) attn_cross_short_kernel(const XsArgs a) {
  __shared__ __align__(16) __nv_bfloat16 sQ[XS_ROWS * XS_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sK[XS_KV * XS_PITCH];
  __shared__ __align__(16) __nv_bfloat16 sV[XS_KV * XS_PITCH];
  pdl_launch_dependents();
  const int seq = blockIdx.z, head = blockIdx.y, tile = blockIdx.x;
  // (the tables are written once per plan, not by the preceding kernel)
  const int4 qd = __ldg(reinterpret_cast<const int4*>(a.seq_table) + seq * 2);
  const int4 kd = __ldg(reinterpret_cast<const int4*>(a.seq_table) + seq * 2 + 1);
  const int rows = min(XS_ROWS, qd.y - tile * XS_ROWS);
  const int q_row0 = qd.x + tile * XS_ROWS;
  const int kv_row0 = kd.z, kv_len = kd.w;
  pdl_wait();  // every CTA waits: the launch must not complete before its predecessor has
  if (rows <= 0) return;
  if (a.q_mask != nullptr && a.q_mask[q_row0 >> a.q_mask_shift] == 0) return;

  // Every 16-byte chunk of Q, K and V is requested at once (cp.async: ~9 per thread in flight, rows beyond
  // the tile / the sequence zero-filled: their P is zero and 0 x garbage must not be NaN) and waited for
  // once: one memory round trip instead of one per loop iteration (12.9 -> 9.9 us per launch under ncu).
  const int tid = threadIdx.x;
  for (int i = tid; i < XS_ROWS * 8; i += XS_THREADS) {  // row i / 8, columns 8 (i % 8) ..
    const int r = i >> 3, c = i & 7;
    const bool ok = r < rows;
    cp_async16(sQ + r * XS_PITCH + c * 8,
               a.q + size_t(q_row0 + (ok ? r : 0)) * a.ldq + a.q_col + head * 64 + c * 8, ok);
  }
  for (int i = tid; i < XS_KV * 8; i += XS_THREADS) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < kv_len;
    const size_t row = size_t(kv_row0 + (ok ? r : 0));
    cp_async16(sK + r * XS_PITCH + c * 8, a.k + row * a.ldk + a.k_col + head * 64 + c * 8, ok);
    cp_async16(sV + r * XS_PITCH + c * 8, a.v + row * a.ldv + a.v_col + head * 64 + c * 8, ok);
  }
  asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  if (warp * 16 >= rows) return;  // (no block-wide barrier below)
  __nv_bfloat16* myQ = sQ + warp * 16 * XS_PITCH;

  // ---- S = Q K^T: 16 rows x 80 keys per warp, fp32 accumulators in registers
  float s[XS_KV / 8][4];
#pragma unroll
  for (int n = 0; n < XS_KV / 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {  // 16 of the 64 head dimensions per step
    uint32_t a0, a1, a2, a3;
    ldsm_x4(a0, a1, a2, a3, myQ + (lane & 15) * XS_PITCH + kk * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int j = 0; j < XS_KV / 16; ++j) {  // two 8-key tiles per ldmatrix.x4
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b0, b1, b2, b3,
              sK + (j * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * XS_PITCH + kk * 16 + ((lane >> 3) & 1) * 8);
      mma_bf16(s[2 * j], a0, a1, a2, a3, b0, b1);
      mma_bf16(s[2 * j + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- softmax over the keys: this thread holds rows r = lane / 4 (c0, c1) and r + 8 (c2, c3),
  // columns 8 n + 2 (lane % 4) + {0, 1} of every tile n; a row is spread over the 4 lanes of a quad
  const int col0 = (lane & 3) * 2;
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < XS_KV / 8; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if (n * 8 + col0 + e >= kv_len) s[n][e] = s[n][2 + e] = -INFINITY;
      m0 = fmaxf(m0, s[n][e]);
      m1 = fmaxf(m1, s[n][2 + e]);
    }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float sc = a.scale_log2, ms0 = m0 * sc, ms1 = m1 * sc;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int n = 0; n < XS_KV / 8; ++n) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[n][e] = fast_exp2(fmaf(s[n][e], sc, -ms0));
      s[n][2 + e] = fast_exp2(fmaf(s[n][2 + e], sc, -ms1));
      l0 += s[n][e];
      l1 += s[n][2 + e];
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  // ---- O = P V: the S accumulators of key tiles 2t, 2t + 1 ARE the A fragment of k-step t
  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int t = 0; t < XS_KV / 16; ++t) {
    const uint32_t p0 = pack_bf16x2(s[2 * t][0], s[2 * t][1]), p1 = pack_bf16x2(s[2 * t][2], s[2 * t][3]);
    const uint32_t p2 = pack_bf16x2(s[2 * t + 1][0], s[2 * t + 1][1]), p3 = pack_bf16x2(s[2 * t + 1][2], s[2 * t + 1][3]);
#pragma unroll
    for (int jd = 0; jd < 4; ++jd) {  // two 8-wide tiles of the head dimension per ldmatrix.x4.trans
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b0, b1, b2, b3,
                sV + (t * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XS_PITCH + jd * 16 + ((lane >> 4) & 1) * 8);
      mma_bf16(o[2 * jd], p0, p1, p2, p3, b0, b1);
      mma_bf16(o[2 * jd + 1], p0, p1, p2, p3, b2, b3);
    }
  }
  // ---- output: normalise, stage the warp's 16 x 64 tile over its own query rows, store 16-byte vectors
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = lane >> 2;
  __syncwarp();  // every lane has read its Q fragments
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(myQ + r0 * XS_PITCH + n * 8 + col0) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
    *reinterpret_cast<uint32_t*>(myQ + (r0 + 8) * XS_PITCH + n * 8 + col0) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
  __syncwarp();
#pragma unroll
  for (int i = lane; i < 16 * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    if (warp * 16 + r < rows)
      *reinterpret_cast<uint4*>(a.out + size_t(q_row0 + warp * 16 + r) * a.ldo + a.o_col + head * 64 + c * 8) =
          *reinterpret_cast<const uint4*>(myQ + r * XS_PITCH + c * 8);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_cross_short_max_keys(void) { return XS_KV; }

extern "C" int b200_attn_cross_short_bf16(const B200AttnSource* q_src, const B200AttnSource* kv_src,
                                          const int32_t* seq_table, int n_seq, int n_heads, int max_q_len,
                                          int max_kv_len, float softmax_scale, const int32_t* q_mask,
                                          int q_mask_shift, void* stream) {
  if (!q_src || !kv_src || !seq_table || n_seq <= 0 || n_heads <= 0 || max_q_len <= 0 || max_kv_len <= 0)
    return B200_ERR_INVALID;
  if (max_kv_len > XS_KV) return B200_ERR_UNSUPPORTED;
  if (!q_src->q || !q_src->out || !kv_src->k || !kv_src->v) return B200_ERR_INVALID;
  if ((q_src->ldq & 7) || (q_src->q_col & 7) || (q_src->ldo & 7) || (q_src->o_col & 7) || (kv_src->ldk & 7) ||
      (kv_src->k_col & 7) || (kv_src->ldv & 7) || (kv_src->v_col & 7))
    return B200_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(q_src->q) | reinterpret_cast<uintptr_t>(q_src->out) |
       reinterpret_cast<uintptr_t>(kv_src->k) | reinterpret_cast<uintptr_t>(kv_src->v)) & 15)
    return B200_ERR_INVALID;
  if (q_mask && q_mask_shift < 7) return B200_ERR_INVALID;  // a CTA is one 128-row tile
  if (n_seq > 65535 || n_heads > 65535) return B200_ERR_INVALID;
  XsArgs a;
  a.q = static_cast<const __nv_bfloat16*>(q_src->q); a.ldq = q_src->ldq; a.q_col = q_src->q_col;
  a.k = static_cast<const __nv_bfloat16*>(kv_src->k); a.ldk = kv_src->ldk; a.k_col = kv_src->k_col;
  a.v = static_cast<const __nv_bfloat16*>(kv_src->v); a.ldv = kv_src->ldv; a.v_col = kv_src->v_col;
  a.out = static_cast<__nv_bfloat16*>(q_src->out); a.ldo = q_src->ldo; a.o_col = q_src->o_col;
  a.seq_table = seq_table;
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  a.q_mask = q_mask; a.q_mask_shift = q_mask_shift;
  const dim3 grid((max_q_len + XS_ROWS - 1) / XS_ROWS, n_heads, n_seq);
  return launch_pdl(attn_cross_short_kernel, grid, dim3(XS_THREADS), 0, reinterpret_cast<cudaStream_t>(stream), a);
}
