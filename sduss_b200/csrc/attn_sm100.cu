// Packed variable-length attention (head_dim 64, non-causal) on tcgen05 / TMEM / TMA.
//
// One launch covers every request of a mixed-resolution batch. A "sequence" is one latent;
// its queries and keys/values are each the concatenation of up to two row segments that live
// in (possibly different) packed buffers:
//   SD3 joint attention : Q = K = V = [image tokens of latent i ; 333 context tokens of i]
//   SD3 attn2 / SDXL self: segment A only (image tokens)
//   SDXL cross attention : Q = segment A (image tokens), K/V = segment B (77 text tokens)
// This replaces the per-resolution Python loops around xformers / SDPA in the reference
// (sduss/model_executor/modules/attention.py:155-203 self, :59-110 cross, :297-368 joint).
//
// CTA = TWO 128-row query tiles (A, B) of one (sequence, head), so the tensor pipe works on one
// tile while the other is in softmax. Warp 0: TMA producer (Q once, K and V rings). Warp 1:
// tcgen05.mma issuer: S_t = Q_t K^T into TMEM, then O_t += P_t V with P read back from TMEM
// (A-from-TMEM MMA) and O accumulated in TMEM. Warps 2-5 / 6-9: online softmax of tile A / B,
// one thread per query row (TMEM lane), whole 128-column S row in registers, P written to TMEM
// as packed bf16 over S. The running max is only moved when it grows by more than 2^8 (lazy
// rescale), in which case the softmax warps rescale O in TMEM before releasing P.
#include "../../include/sduss_b200.h"
#include "host_util.h"
#include <cstdio>

#include "ptx.cuh"

namespace b200 {

constexpr int ATT_BM = 128;   // query rows per tile; a CTA owns TWO tiles (256 rows)
constexpr int ATT_BN = 128;   // kv rows per tile
constexpr int ATT_D = 64;     // head dim
constexpr int ATT_KS = 3;     // K ring depth
constexpr int ATT_VS = 3;     // V ring depth
constexpr int ATT_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-5 softmax A, warps 6-9 softmax B
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KB
constexpr int ATT_SMEM = (2 + ATT_KS + ATT_VS) * ATT_TILE_BYTES + 1024 + 256;
constexpr float ATT_RESCALE_THRESHOLD = 8.f;  // lazy rescale: keep a stale max while p <= 2^8

struct AttnArgs {
  const int* seq_table;   // [n_seq][8]: qa_row, qa_len, qb_row, qb_len, ka_row, ka_len, kb_row, kb_len
  const int* work_items;  // [n_items][4]: seq, q_seg (0 = A, 1 = B), row offset in segment, unused
  int q_col[2], k_col[2], v_col[2];  // column of head 0 inside each source buffer
  __nv_bfloat16* out[2];             // output buffers for Q segment A / B
  int ldo[2];
  int o_col[2];
  float scale_log2;                  // softmax scale * log2(e)
  long long* dbg;                    // optional [gridDim.x * gridDim.y][8] phase cycle counters (ATT_TIMING builds)
};

#ifdef ATT_TIMING
#define ATT_T(var) const long long var = clock64()
#define ATT_ACC(slot, a, b) tacc[slot] += (b) - (a)
#else
#define ATT_T(var)
#define ATT_ACC(slot, a, b)
#endif

// Named barriers (ids 1, 2) ping-pong the MUFU-heavy exp phase between the two softmax
// warpgroups: while one group exponentiates (MUFU at full rate), the other loads its next S
// row, reduces the max, stores P and talks to the MMA warp. Without it both groups drift into
// lock-step, share the MUFU pipe for ~2200 cycles and leave it idle for ~1200 (measured).
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// volatile variant: keeps its source position relative to the tcgen05.st of the previous chunk,
// which pins the software pipelining of the exp loop (ptxas otherwise sinks all packs to the end)
__device__ __forceinline__ float fast_exp2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max rel. error
// 7.5e-5, far below the bf16 rounding of P). A quarter of the exponentials go through it so the
// MUFU pipe (16 ex2/clk/SM, the co-bottleneck of head_dim-64 attention) is relieved.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float magic = 12582912.f;          // 1.5 * 2^23: x + magic rounds x to an integer
  const float xf = x + magic;
  const float f = x - (xf - magic);        // fractional part in [-0.5, 0.5]
  const float p = fmaf(fmaf(fmaf(0.0551708528f, f, 0.242609396f), f, 0.693260959f), f, 0.999928182f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xf) << 23));  // p * 2^round(x)
}

// fp32 -> bf16 pairs on the ALU pipe (IADD + PRMT) instead of F2FP: the conversion instruction
// shares the 16/clk/SM XU pipe with MUFU.EX2, where it would add 50% to the exp phase.
// Round-half-up on the magnitude (p >= 0); differs from RN only on exact ties.
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
  return __byte_perm(a, b, 0x7632);
}

constexpr int ATT_PINGPONG = 0;  // measured: 574 -> 519 TFLOP/s (a lone warp's exp phase is not MUFU-bound)
constexpr int ATT_ALU_PACK = 0;  // measured: F2FP 574 vs IADD+PRMT 562 TFLOP/s
constexpr int ATT_POLY_MASK = 0;  // 1: every 4th exponential by polynomial. Measured on B200: 574 -> 520
                                  // TFLOP/s (issue slots, not MUFU alone, bound the softmax warps), so off.

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQA, const __grid_constant__ CUtensorMap tmQB,
                const __grid_constant__ CUtensorMap tmKA, const __grid_constant__ CUtensorMap tmKB,
                const __grid_constant__ CUtensorMap tmVA, const __grid_constant__ CUtensorMap tmVB,
                AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                                 // 2 tiles
  uint8_t* sK = smem + 2 * ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_KS * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_VS * ATT_TILE_BYTES);
  uint64_t* q_full = bars;             // 1
  uint64_t* k_full = bars + 1;         // KS
  uint64_t* k_empty = k_full + ATT_KS; // KS
  uint64_t* v_full = k_empty + ATT_KS; // VS
  uint64_t* v_empty = v_full + ATT_VS; // VS
  uint64_t* s_full = v_empty + ATT_VS; // 2 (per query tile)
  uint64_t* p_full = s_full + 2;       // 2
  uint64_t* o_full = p_full + 2;       // 2
  uint64_t* s_free = o_full + 2;       // 2: softmax has copied S_t to registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;

  // work tables are written once per plan (not by the preceding kernel): safe before pdl_wait()
  const int4 item = reinterpret_cast<const int4*>(a.work_items)[blockIdx.x];
  const int* st = a.seq_table + item.x * 8;
  const int q_seg = item.y;
  const int q_row0 = st[q_seg * 2] + item.z;
  const int q_rows = min(2 * ATT_BM, st[q_seg * 2 + 1] - item.z);  // valid query rows in this CTA
  const int nq = q_rows > ATT_BM ? 2 : 1;                          // active query tiles
  const int ka_row = st[4], ka_len = st[5], kb_row = st[6], kb_len = st[7];
  const int nA = (ka_len + ATT_BN - 1) / ATT_BN;
  const int nB = (kb_len + ATT_BN - 1) / ATT_BN;
  const int n_tiles = nA + nB;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < ATT_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: S_t at 128 t, P_t (packed bf16) at 256 + 64 t, O_t at 384 + 64 t. S and P are
  // separate so that S_t(j+1) = Q_t K(j+1)^T can be issued while softmax still works on tile j.
  const uint32_t tS = tmem_base;
  const uint32_t tP = tmem_base + 256;
  const uint32_t tO = tmem_base + 384;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const CUtensorMap* qm = q_seg == 0 ? &tmQA : &tmQB;
      mbar_expect_tx(q_full, nq * ATT_TILE_BYTES);
      tma_load_2d(sQ, qm, q_full, a.q_col[q_seg] + head * ATT_D, q_row0);
      if (nq == 2)
        tma_load_2d(sQ + ATT_TILE_BYTES, qm, q_full, a.q_col[q_seg] + head * ATT_D, q_row0 + ATT_BM);
    }
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    for (int j = 0; j < n_tiles; ++j) {
      const bool inA = j < nA;
      const int row = inA ? ka_row + j * ATT_BN : kb_row + (j - nA) * ATT_BN;
      mbar_wait(&k_empty[ks], kph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&k_full[ks], ATT_TILE_BYTES);
        tma_load_2d(sK + ks * ATT_TILE_BYTES, inA ? &tmKA : &tmKB, &k_full[ks],
                    a.k_col[inA ? 0 : 1] + head * ATT_D, row);
      }
      mbar_wait(&v_empty[vs], vph ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&v_full[vs], ATT_TILE_BYTES);
        tma_load_2d(sV + vs * ATT_TILE_BYTES, inA ? &tmVA : &tmVB, &v_full[vs],
                    a.v_col[inA ? 0 : 1] + head * ATT_D, row);
      }
      __syncwarp();
      if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
      if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // V is MN-major
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    mbar_wait(q_full, 0);
    tc_fence_after();
    auto issue_qk = [&](int t, uint32_t k_addr) {  // S_t = Q_t K^T
      if (lane == 0) {
        const uint64_t dq = make_sdesc_sw128(smem_u32(sQ + t * ATT_TILE_BYTES));
        const uint64_t dk = make_sdesc_sw128(k_addr);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_ss(tS + t * ATT_BN, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_qk,
                  k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, uint32_t v_addr, int j) {  // O_t (+)= P_t V
      if (lane == 0) {
        const uint64_t dv = make_sdesc_sw128(v_addr);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k)
          // A: 16 bf16 of K = 8 TMEM columns; B: 16 kv rows = 2048 bytes (>>4 = 128)
          umma_ts(tO + t * ATT_D, tP + t * ATT_D + 8 * k, dv + uint64_t(128 * k), idesc_pv,
                  (j | k) != 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
      }
      __syncwarp();
    };
    // prologue: S_A(0), S_B(0)
    mbar_wait(&k_full[ks], kph);
    tc_fence_after();
    for (int t = 0; t < nq; ++t) issue_qk(t, smem_u32(sK + ks * ATT_TILE_BYTES));
    if (lane == 0) umma_commit(&k_empty[ks]);
    __syncwarp();
    if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const bool more = j + 1 < n_tiles;
      if (more) {
        // S_t(j+1) as soon as softmax has pulled S_t(j) into registers
        mbar_wait(&k_full[ks], kph);
        const uint32_t k_addr = smem_u32(sK + ks * ATT_TILE_BYTES);
        for (int t = 0; t < nq; ++t) {
          mbar_wait(&s_free[t], ph);
          tc_fence_after();
          issue_qk(t, k_addr);
        }
        if (lane == 0) umma_commit(&k_empty[ks]);
        __syncwarp();
        if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
      }
      mbar_wait(&v_full[vs], vph);
      const uint32_t v_addr = smem_u32(sV + vs * ATT_TILE_BYTES);
      for (int t = 0; t < nq; ++t) {
        mbar_wait(&p_full[t], ph);
        tc_fence_after();
        issue_pv(t, v_addr, j);
      }
      if (lane == 0) umma_commit(&v_empty[vs]);
      __syncwarp();
      if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ softmax + output
    const int t = (warp - 2) >> 2;  // query tile of this warpgroup
    if (t < nq) {
      const int qd = warp & 3;
      const int r = qd * 32 + lane;  // query row inside the tile == TMEM lane
      const uint32_t lane_off = uint32_t(qd * 32) << 16;
      const uint32_t t_s = tS + lane_off + t * ATT_BN;
      const uint32_t t_o = tO + lane_off + t * ATT_D;
      const uint32_t t_p = tP + lane_off + t * ATT_D;
      float m_run = -INFINITY, l_run = 0.f;
      const float sc = a.scale_log2;
#ifdef ATT_TIMING
      long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const long long t_begin = clock64();
#endif
      // Ping-pong of the XU-heavy exp phase between the two warpgroups (named barriers 1, 2).
      const bool pingpong = ATT_PINGPONG && nq == 2;
      if (pingpong && t == 1) named_bar_arrive(1, 256);  // group A exponentiates first

      for (int j = 0; j < n_tiles; ++j) {
        const bool inA = j < nA;
        const int n_valid =
            inA ? min(ATT_BN, ka_len - j * ATT_BN) : min(ATT_BN, kb_len - (j - nA) * ATT_BN);
        ATT_T(c0);
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        ATT_T(c1);
        float s[ATT_BN];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(t_s + c * 32, reinterpret_cast<uint32_t*>(s) + c * 32);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);  // S_t may be overwritten by the next Q K^T
        ATT_T(c2);
        if (n_valid < ATT_BN) {
#pragma unroll
          for (int i = 0; i < ATT_BN; ++i)
            if (i >= n_valid) s[i] = -INFINITY;
        }
        // 8 independent chains (a single 127-deep fmax chain would expose ~500 cycles of latency)
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = s[i];
#pragma unroll
        for (int i = 8; i < ATT_BN; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], s[i]);
        const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                               fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        const float m_new = fmaxf(m_run, mx);
        // lazy rescale: only move the reference max when it grew by more than 2^8
        const bool resc = (m_new - m_run) * sc > ATT_RESCALE_THRESHOLD;
        float alpha = 1.f;
        if (resc) {
          alpha = fast_exp2((m_run - m_new) * sc);
          m_run = m_new;
        }
        const float m_sc = m_run * sc;
        if (pingpong) named_bar_sync(1 + t, 256);  // my turn on the XU pipe
        ATT_T(c3);
        // exp in chunks of 16 columns, in place, software-pipelined by one chunk: the MUFU.EX2 of
        // chunk c are issued (pinned order) before chunk c-1 is summed and packed to bf16, so the
        // FADD / F2FP work sits in the shadow of the 8-cycle MUFU issue interval.
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[ATT_BN / 2];
#pragma unroll
        for (int c = 0; c <= ATT_BN / 16; ++c) {
          if (c < ATT_BN / 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[c * 16 + i] = fast_exp2_pinned(fmaf(s[c * 16 + i], sc, -m_sc));
          }
          if (c > 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const int k = (c - 1) * 16 + i;
              sum4[(i >> 1) & 3] += s[k] + s[k + 1];
              pk[k >> 1] = pack_bf16x2(s[k], s[k + 1]);
            }
          }
        }
        const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        l_run = l_run * alpha + sum;
        if (pingpong && !(t == 1 && j == n_tiles - 1)) named_bar_arrive(2 - t, 256);  // hand over
        ATT_T(c4);
        if (j > 0) {
          // PV_t(j-1) must be complete before P_t is overwritten and before O_t is rescaled
          mbar_wait(&o_full[t], (j - 1) & 1);
          tc_fence_after();
          if (__any_sync(0xffffffffu, resc)) {
            uint32_t o[ATT_D];
            tmem_ld32(t_o, o);
            tmem_ld32(t_o + 32, o + 32);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < ATT_D; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(t_o, o);
            tmem_st32(t_o + 32, o + 32);
          }
        }
        ATT_T(c5);
        tmem_st32(t_p, pk);
        tmem_st32(t_p + 32, pk + 32);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        ATT_T(c6);
        ATT_ACC(0, c0, c1); ATT_ACC(1, c1, c2); ATT_ACC(2, c2, c3); ATT_ACC(3, c3, c4);
        ATT_ACC(4, c4, c5); ATT_ACC(5, c5, c6);
      }
#ifdef ATT_TIMING
      if (a.dbg != nullptr && lane == 0 && (warp == 2 || warp == 6)) {
        long long* d = a.dbg + (size_t(blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (warp == 6 ? 8 : 0);
        for (int i = 0; i < 6; ++i) d[i] = tacc[i];
        d[6] = clock64() - t_begin;
        d[7] = n_tiles;
      }
#endif
      mbar_wait(&o_full[t], (n_tiles - 1) & 1);
      tc_fence_after();
      uint32_t o[ATT_D];
      tmem_ld32(t_o, o);
      tmem_ld32(t_o + 32, o + 32);
      tmem_wait_ld();
      if (t * ATT_BM + r < q_rows) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* op = a.out[q_seg] + size_t(q_row0 + t * ATT_BM + r) * a.ldo[q_seg] +
                            a.o_col[q_seg] + head * ATT_D;
#pragma unroll
        for (int i = 0; i < ATT_D; i += 8) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + i) = v;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_rows_map(CUtensorMap* m, const void* base, int rows, int cols, int ld) {
  if (base == nullptr) return B200_ERR_INVALID;
  uint64_t d[2] = {uint64_t(cols), uint64_t(rows)}, s[1] = {uint64_t(ld) * 2};
  uint32_t b[2] = {ATT_D, 128};
  return get_tmap_bf16_sw128(m, base, 2, d, s, b);
}

}  // namespace b200

using namespace b200;

#ifdef ATT_TIMING
long long* g_att_dbg = nullptr;
extern "C" long long* b200_attn_debug_buffer(void) { return g_att_dbg; }
#endif

extern "C" int b200_attn_varlen_bf16(const B200AttnSource* src_a, const B200AttnSource* src_b,
                                     const int32_t* seq_table, const int32_t* work_items,
                                     int n_items, int n_heads, float softmax_scale, void* stream_) {
  if (!src_a || !seq_table || !work_items || n_items <= 0 || n_heads <= 0) return B200_ERR_INVALID;
  const B200AttnSource* srcs[2] = {src_a, src_b ? src_b : src_a};
  CUtensorMap tm[2][3];
  bool have[2][3] = {{false, false, false}, {false, false, false}};
  const CUtensorMap* any = nullptr;
  AttnArgs a;
  for (int s = 0; s < 2; ++s) {
    const B200AttnSource* p = srcs[s];
    const void* bases[3] = {p->q, p->k, p->v};
    const int rows[3] = {p->q_rows, p->kv_rows, p->kv_rows};
    const int lds[3] = {p->ldq, p->ldk, p->ldv};
    for (int t = 0; t < 3; ++t) {
      if (bases[t] == nullptr) continue;  // this side has no such segment
      if ((lds[t] & 7) || rows[t] <= 0) return B200_ERR_INVALID;
      int rc = make_rows_map(&tm[s][t], bases[t], rows[t], lds[t], lds[t]);
      if (rc) return rc;
      have[s][t] = true;
      if (!any) any = &tm[s][t];
    }
    a.q_col[s] = p->q_col;
    a.k_col[s] = p->k_col;
    a.v_col[s] = p->v_col;
    a.out[s] = static_cast<__nv_bfloat16*>(p->out);
    a.ldo[s] = p->ldo;
    a.o_col[s] = p->o_col;
  }
  if (!any) return B200_ERR_INVALID;
  // Absent segments get a valid (never dereferenced) map so the kernel signature stays fixed.
  for (int s = 0; s < 2; ++s)
    for (int t = 0; t < 3; ++t)
      if (!have[s][t]) tm[s][t] = *any;
  a.seq_table = seq_table;
  a.work_items = work_items;
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  a.dbg = nullptr;
#ifdef ATT_TIMING
  {
    static long long* dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, size_t(1) << 26);
    a.dbg = dbg_buf;
    cudaMemsetAsync(dbg_buf, 0, size_t(n_items) * n_heads * 16 * 8, reinterpret_cast<cudaStream_t>(stream_));
    extern long long* g_att_dbg;
    g_att_dbg = dbg_buf;
  }
#endif
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM);
    if (err != cudaSuccess) return static_cast<int>(err);
    configured = true;
  }
#ifdef ATT_TIMING
  {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, attn_fwd_kernel);
    static bool once = false;
    if (!once) {
      once = true;
      printf("attn attrs: maxThreadsPerBlock=%d numRegs=%d sharedStatic=%zu maxDyn=%d local=%zu\n",
             fa.maxThreadsPerBlock, fa.numRegs, fa.sharedSizeBytes, fa.maxDynamicSharedSizeBytes,
             fa.localSizeBytes);
    }
  }
#endif
  dim3 grid(n_items, n_heads);
  return launch_pdl(attn_fwd_kernel, grid, dim3(ATT_THREADS), ATT_SMEM,
                    reinterpret_cast<cudaStream_t>(stream_), tm[0][0], tm[1][0], tm[0][1], tm[1][1],
                    tm[0][2], tm[1][2], a);
}
