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
// The kernel is persistent: one CTA per SM walks a host-built list of work units (schedule below).
// A unit = ATT_NQ 128-row query tiles of one (sequence, head), walking the keys in steps of ATT_BN
// rows. Warps [0, 4 NQ): online softmax, one warpgroup per query tile, one thread per query row
// (TMEM lane), the whole ATT_BN-column S row in registers, P written back to TMEM as packed
// bf16. Warp 4 NQ: TMA producer (Q once, K and V rings). Warps 4 NQ + 1 + t: one tcgen05.mma
// issuer per query tile: S_t = Q_t K^T into TMEM, then O_t += P_t V with P read from TMEM
// (A-from-TMEM MMA) and O accumulated in TMEM.
//
// Why this shape (measured on B200, profiles/r01_attn_*): at head_dim 64 the kernel is bound by
// the softmax, not by the tensor pipe: 16 MUFU.EX2 / clk / SM against 4096 MAC / clk / SM means
// 1024 cycles of exponentials per 512 cycles of MMA. A softmax warp is a long dependent chain
// (wait S, TMEM load, max, exp, pack, TMEM store, hand over) and the only thing that keeps the
// MUFU pipe busy while one warp is in its non-exp part is another warp of the same SM
// sub-partition. NQ = 3 tiles with 64-key steps gives three softmax warps per sub-partition with
// chains half as long as the former 2 x 128 layout (and fits TMEM: 3 x (64 S + 32 P + 64 O) = 480
// columns). One issuer warp per tile because a single in-order issuer makes the tiles wait for
// each other (PV_0 queued behind the wait for tile 1's S buffer) and pushes them into lock-step.
// The running max is only moved when it grows by more than 2^8 (lazy rescale), in which case the
// softmax warps rescale O in TMEM before releasing P.
#include "../../include/sduss_b200.h"
#include "host_util.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <queue>
#include <vector>

#include "ptx.cuh"

namespace b200 {

#ifndef ATT_NQ
#define ATT_NQ 3              // 128-row query tiles per CTA
#endif
#ifndef ATT_BN_
#define ATT_BN_ 64            // key rows per step
#endif
#ifndef ATT_KS_
#define ATT_KS_ (ATT_BN_ == 64 ? 4 : 3)
#endif
#ifndef ATT_VS_
#define ATT_VS_ (ATT_BN_ == 64 ? 4 : 3)
#endif
constexpr int ATT_BM = 128;      // query rows per tile
constexpr int ATT_BN = ATT_BN_;  // kv rows per step
constexpr int ATT_D = 64;        // head dim
constexpr int ATT_KS = ATT_KS_;  // K ring depth
constexpr int ATT_VS = ATT_VS_;  // V ring depth
constexpr int ATT_SM_WARPS = 4 * ATT_NQ;
constexpr int ATT_TMA_WARP = ATT_SM_WARPS;
constexpr int ATT_MMA_WARP0 = ATT_SM_WARPS + 1;
constexpr int ATT_THREADS = 32 * (ATT_SM_WARPS + 1 + ATT_NQ);
constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;   // 16 KB
constexpr int ATT_KV_BYTES = ATT_BN * ATT_D * 2;  // 8 or 16 KB
constexpr int ATT_BOX_ROWS = 64;                  // rows per TMA box (all maps)
constexpr int ATT_BOX_BYTES = ATT_BOX_ROWS * ATT_D * 2;
constexpr int ATT_QBUF = 2;  // Q tile sets: the next unit's queries load under the current unit
constexpr int ATT_SD = 4;    // depth of the in-CTA unit-id ring (scheduler warp -> other roles)
constexpr int ATT_SMEM =
    ATT_QBUF * ATT_NQ * ATT_Q_BYTES + (ATT_KS + ATT_VS) * ATT_KV_BYTES + 1024 + 512;
constexpr int ATT_TMEM_USED = ATT_NQ * (ATT_BN + ATT_BN / 2 + ATT_D);
constexpr int ATT_TMEM_COLS = ATT_TMEM_USED <= 256 ? 256 : 512;
static_assert(ATT_TMEM_USED <= 512, "TMEM budget");
static_assert(ATT_BN == 64 || ATT_BN == 128, "kv step");
constexpr float ATT_RESCALE_THRESHOLD = 8.f;  // lazy rescale: keep a stale max while p <= 2^8
#ifndef ATT_F32X2
#define ATT_F32X2 1  // packed fp32 pairs (FFMA2 / FADD2) in the softmax
#endif
#ifndef ATT_POLY_MOD
#define ATT_POLY_MOD 0  // > 0: every ATT_POLY_MOD-th element pair takes exp2 on the FMA pipe (below)
#endif
#ifndef ATT_POLY_DEG
#define ATT_POLY_DEG 3
#endif
#ifndef ATT_SKEW
#define ATT_SKEW 0  // one-time start offset (cycles) between the softmax groups of a CTA (A/B knob)
#endif

struct AttnArgs {
  const int* seq_table;   // [n_seq][8]: qa_row, qa_len, qb_row, qb_len, ka_row, ka_len, kb_row, kb_len
  const int* work_units;  // [n_units][4]: seq, q_seg (0 = A, 1 = B), row offset in segment, head
  int* sched;             // [2]: next unit to hand out (minus gridDim.x), CTAs finished; self-resetting
  int n_units;
  int q_col[2], k_col[2], v_col[2];  // column of head 0 inside each source buffer
  __nv_bfloat16* out[2];             // output buffers for Q segment A / B
  int ldo[2];
  int o_col[2];
  float scale_log2;                  // softmax scale * log2(e)
  // prepare-stage variants (text encoders; uniform per launch, off on the denoising path):
  int causal;                        // CLIP: key position <= query position
  const float* rel_bias;             // T5: logits += rel_bias[head][k_pos - q_pos + rel_len - 1]
  int rel_len, rel_ld;               //     (already divided by the softmax scale)
  // patch cache: query tiles of segment A whose chunk (row >> shift) has mask 0 are
  // skipped (their output rows stay as they are); keys / values are read whatever their mask
  const int* q_mask; int q_mask_shift;
  long long* dbg;                    // optional phase cycle counters (ATT_TIMING builds)
};
// BOUNDED (second template flag of the kernel; denoising instantiations, with or without the query-tile mask): the caller guarantees
// |logit * softmax_scale * log2(e)| <= 64 for every (query, key) pair -- SD3.5's joint attention, whose q
// and k are RMS-normalised per head (|q|, |k| <= 8 max|w|: the model checks the learned weights at load,
// sd3_transformer.py). exp2 of such a value needs no reference maximum: bf16 / fp32 hold 2^-64 .. 2^64
// with full relative precision, the row sum and the P V accumulation stay far inside fp32. The softmax
// warps then skip the row-maximum pass (255 of the ~2300 cycles of a step's dependent chain, the bound of
// this kernel: DESIGN.md 4.1), the running maximum and the O rescale; the result is the same softmax,
// rounded at different points.

#ifdef ATT_TIMING
#define ATT_T(var) const long long var = clock64()
#define ATT_ACC(slot, a, b) tacc[slot] += (b) - (a)
#else
#define ATT_T(var)
#define ATT_ACC(slot, a, b)
#endif

// One work unit, decoded from the tables (written once per plan, not by the preceding kernel:
// safe to read before pdl_wait()).
struct AttnUnit {
  int q_seg, q_row0, q_rows, nq, head, q_pos0;
  int active;  // bit t: query tile t of the unit is computed
  int ka_row, ka_len, kb_row, kb_len, nA, n_tiles;
};

// EXTRAS: 0 = the denoising path; 1 = + query-tile skip (patch cache); 2 = + causal mask / relative
// position bias (text encoders). Separate instantiations because at 126-128 registers per thread
// the extra live values of the in-loop variants cost 6 % (685 vs 732 TFLOP/s on the config-2 joint
// shapes, profiles/r02_attn_variants.txt), branch or no branch.
template <int EXTRAS>
__device__ __forceinline__ AttnUnit load_unit(const AttnArgs& a, int u) {
  const int4 item = __ldg(reinterpret_cast<const int4*>(a.work_units) + u);
  const int4* st = reinterpret_cast<const int4*>(a.seq_table + item.x * 8);
  const int4 q = __ldg(st), k = __ldg(st + 1);
  AttnUnit w;
  w.q_seg = item.y;
  w.head = item.w;
  const int seg_row = w.q_seg == 0 ? q.x : q.z, seg_len = w.q_seg == 0 ? q.y : q.w;
  w.q_row0 = seg_row + item.z;
  w.q_pos0 = (w.q_seg == 0 ? 0 : q.y) + item.z;     // position of the unit's first query in its sequence
  w.q_rows = min(ATT_NQ * ATT_BM, seg_len - item.z);  // valid query rows of this unit
  w.nq = (w.q_rows + ATT_BM - 1) / ATT_BM;            // query tiles with rows
  w.active = 0;
  for (int t = 0; t < w.nq; ++t)
    if (EXTRAS == 0 || w.q_seg != 0 || a.q_mask == nullptr ||
        a.q_mask[(w.q_row0 + t * ATT_BM) >> a.q_mask_shift] != 0)
      w.active |= 1 << t;
  w.ka_row = k.x; w.ka_len = k.y; w.kb_row = k.z; w.kb_len = k.w;
  w.nA = (w.ka_len + ATT_BN - 1) / ATT_BN;
  w.n_tiles = w.nA + (w.kb_len + ATT_BN - 1) / ATT_BN;
  return w;
}

#ifndef ATT_POLY_BOUNDED
#define ATT_POLY_BOUNDED 5  // the same for the BOUNDED instantiations, where it pays (below); 0 = off
#endif
// exp2 of a pair on the FMA + ALU pipes instead of MUFU (the kernel's bound, see the header):
// Cody-Waite split x = n + r with the 1.5 * 2^23 magic add (n lands in the low mantissa bits of t),
// packed degree-2/3 minimax of 2^r on [-0.5, 0.5] (relative error 1.7e-3 / 7.5e-5: below the bf16
// rounding of P), exponent inserted with a shift + integer add. 3-4 FFMA2/FADD2 issue slots + 6
// ALU ops per pair against two 8-cycle MUFU slots.
__device__ __forceinline__ float2 poly_exp2_pair(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = __fadd2_rn(x, magic);
  const float2 fl = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 r = __fadd2_rn(x, make_float2(-fl.x, -fl.y));
#if ATT_POLY_DEG == 2
  float2 p = __ffma2_rn(make_float2(0.238428926f, 0.238428926f), r, make_float2(0.703448005f, 0.703448005f));
  p = __ffma2_rn(p, r, make_float2(1.00044314f, 1.00044314f));
#else
  float2 p = __ffma2_rn(make_float2(0.0551716677f, 0.0551716677f), r, make_float2(0.242611122f, 0.242611122f));
  p = __ffma2_rn(p, r, make_float2(0.693260986f, 0.693260986f));
  p = __ffma2_rn(p, r, make_float2(0.999928074f, 0.999928074f));
  float2 out;
  out.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  out.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return out;
}
#endif

template <int EXTRAS, bool BOUNDED = false>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQA, const __grid_constant__ CUtensorMap tmQB,
                const __grid_constant__ CUtensorMap tmKA, const __grid_constant__ CUtensorMap tmKB,
                const __grid_constant__ CUtensorMap tmVA, const __grid_constant__ CUtensorMap tmVB,
                AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* sQ = smem;                                 // QBUF sets of NQ tiles
  uint8_t* sK = smem + ATT_QBUF * ATT_NQ * ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KS * ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_VS * ATT_KV_BYTES);
  uint64_t* q_full = bars;                  // QBUF
  uint64_t* q_empty = q_full + ATT_QBUF;    // QBUF: every issuer has issued its last Q K^T of the unit
  uint64_t* k_full = q_empty + ATT_QBUF;    // KS
  uint64_t* k_empty = k_full + ATT_KS;      // KS
  uint64_t* v_full = k_empty + ATT_KS;      // VS
  uint64_t* v_empty = v_full + ATT_VS;      // VS
  uint64_t* s_full = v_empty + ATT_VS;      // NQ (per query tile)
  uint64_t* p_full = s_full + ATT_NQ;       // NQ
  uint64_t* o_full = p_full + ATT_NQ;       // NQ
  uint64_t* s_free = o_full + ATT_NQ;       // NQ: softmax has copied S_t to registers
  uint64_t* u_full = s_free + ATT_NQ;       // SD: unit id published
  uint64_t* u_empty = u_full + ATT_SD;      // SD: read by every other warp
  volatile int* u_ids = reinterpret_cast<volatile int*>(u_empty + ATT_SD);  // SD
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(const_cast<int*>(u_ids) + ATT_SD);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Persistent CTA, one per SM. Units are sorted longest first (b200_attn_build_schedule); CTA b
  // starts with unit b and then takes the next unit from a global counter, which balances the SMs
  // like the hardware CTA dispatcher would, without paying launch + TMEM allocation + load latency
  // per unit (~2 us, 7-20 % of a short sequence): the roles below walk the unit sequence
  // independently and meet only through mbarriers, so the Q / K / V loads and the first Q K^T of
  // the next unit run under the last softmax steps and the output write of the current one, and
  // the query tiles of a CTA drift apart freely (Q is double-buffered, the K / V rings bound the
  // drift). The producer warp draws the unit ids and hands them to the other warps through a
  // small smem ring (u_ids).

  if (warp == ATT_TMA_WARP && lane == 0) {
    for (int i = 0; i < ATT_QBUF; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], ATT_NQ);
    }
    // Every ring stage is released by all ATT_NQ issuer warps (idle tiles release it unread), so
    // the arrival counts do not depend on the unit.
    for (int i = 0; i < ATT_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], ATT_NQ);
    }
    for (int i = 0; i < ATT_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], ATT_NQ);
    }
    for (int i = 0; i < ATT_NQ; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], 4);
    }
    for (int i = 0; i < ATT_SD; ++i) {
      mbar_init(&u_full[i], 1);
      mbar_init(&u_empty[i], ATT_SM_WARPS + ATT_NQ);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();
  // TMEM columns: S_t at BN t, P_t (packed bf16) after the S block, O_t after the P block. S and
  // P are separate so that S_t(j+1) = Q_t K(j+1)^T can be issued while softmax works on step j.
  const uint32_t tS = tmem_base;
  const uint32_t tP = tmem_base + ATT_NQ * ATT_BN;
  const uint32_t tO = tP + ATT_NQ * (ATT_BN / 2);
  // next unit id of this warp (consumer side of the u_ids ring); -1 = no more work
  auto take_unit = [&](uint32_t it) -> int {
    const uint32_t slot = it % ATT_SD;
    mbar_wait(&u_full[slot], (it / ATT_SD) & 1);
    const int u = u_ids[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&u_empty[slot]);
    return u;
  };

  if (warp == ATT_TMA_WARP) {
    // ------------------------------------------------------------ TMA producer
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    int u = blockIdx.x;  // grid <= n_units
    uint32_t qi = 0;     // units with at least one computed tile (they own a Q buffer set)
    for (uint32_t it = 0;; ++it) {
      // draw the unit after this one now; the result is only needed at the end of the iteration
      int drawn = 0;
      if (u >= 0 && lane == 0) drawn = atomicAdd(a.sched, 1);
      const uint32_t slot = it % ATT_SD;
      mbar_wait(&u_empty[slot], ((it / ATT_SD) & 1) ^ 1);
      if (lane == 0) {
        u_ids[slot] = u;
        mbar_arrive(&u_full[slot]);
      }
      if (u < 0) break;
      const AttnUnit w = load_unit<EXTRAS>(a, u);
      const int n_steps_u = w.active != 0 ? w.n_tiles : 0;  // a unit of clean patches loads nothing
      const uint32_t qb = qi % ATT_QBUF;
      if (w.active != 0) mbar_wait(&q_empty[qb], ((qi / ATT_QBUF) & 1) ^ 1);  // unit qi - QBUF no longer reads this set
      if (lane == 0 && w.active != 0) {
        const CUtensorMap* qm = w.q_seg == 0 ? &tmQA : &tmQB;
        const int boxes = (w.q_rows + ATT_BOX_ROWS - 1) / ATT_BOX_ROWS;
        mbar_expect_tx(&q_full[qb], boxes * ATT_BOX_BYTES);
        for (int b = 0; b < boxes; ++b)
          tma_load_2d(sQ + qb * (ATT_NQ * ATT_Q_BYTES) + b * ATT_BOX_BYTES, qm, &q_full[qb],
                      a.q_col[w.q_seg] + w.head * ATT_D, w.q_row0 + b * ATT_BOX_ROWS);
      }
      if (w.active != 0) ++qi;
      for (int j = 0; j < n_steps_u; ++j) {
        const bool inA = j < w.nA;
        const int row = inA ? w.ka_row + j * ATT_BN : w.kb_row + (j - w.nA) * ATT_BN;
        mbar_wait(&k_empty[ks], kph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&k_full[ks], ATT_KV_BYTES);
#pragma unroll
          for (int b = 0; b < ATT_BN / ATT_BOX_ROWS; ++b)
            tma_load_2d(sK + ks * ATT_KV_BYTES + b * ATT_BOX_BYTES, inA ? &tmKA : &tmKB,
                        &k_full[ks], a.k_col[inA ? 0 : 1] + w.head * ATT_D, row + b * ATT_BOX_ROWS);
        }
        mbar_wait(&v_empty[vs], vph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&v_full[vs], ATT_KV_BYTES);
#pragma unroll
          for (int b = 0; b < ATT_BN / ATT_BOX_ROWS; ++b)
            tma_load_2d(sV + vs * ATT_KV_BYTES + b * ATT_BOX_BYTES, inA ? &tmVA : &tmVB,
                        &v_full[vs], a.v_col[inA ? 0 : 1] + w.head * ATT_D, row + b * ATT_BOX_ROWS);
        }
        __syncwarp();
        if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
        if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
      }
      drawn = __shfl_sync(0xffffffffu, drawn, 0) + int(gridDim.x);
      u = drawn < a.n_units ? drawn : -1;
    }
  } else if (warp > ATT_TMA_WARP) {
    // ------------------------------------------------------------ MMA issuer of query tile t
    const int t = warp - ATT_MMA_WARP0;
    constexpr uint32_t idesc_qk = make_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BM, ATT_D, 0, 1);  // V is MN-major
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t g = 0;  // key steps tile t has run in earlier units (phase of its s/p/o barriers)
    uint32_t qi = 0; // units with at least one computed tile (Q buffer sets, as in the producer)
    uint64_t dq = 0;
    auto issue_qk = [&](uint32_t k_addr) {  // S_t = Q_t K^T
      if (lane == 0) {
        const uint64_t dk = make_sdesc_sw128(k_addr);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          umma_ss(tS + t * ATT_BN, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_qk,
                  k != 0 ? 1u : 0u);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](uint32_t v_addr, int j) {  // O_t (+)= P_t V
      if (lane == 0) {
        const uint64_t dv = make_sdesc_sw128(v_addr);
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k)
          // A: 16 bf16 of K = 8 TMEM columns; B: 16 kv rows = 2048 bytes (>>4 = 128)
          umma_ts(tO + t * ATT_D, tP + t * (ATT_BN / 2) + 8 * k, dv + uint64_t(128 * k), idesc_pv,
                  (j | k) != 0 ? 1u : 0u);
        umma_commit(&o_full[t]);
      }
      __syncwarp();
    };
    for (uint32_t it = 0;; ++it) {
      const int u = take_unit(it);
      if (u < 0) break;
      const AttnUnit w = load_unit<EXTRAS>(a, u);
      if (w.active == 0) continue;  // every query tile of the unit belongs to a clean patch
      const uint32_t qb = qi % ATT_QBUF;
      uint64_t* q_empty_u = &q_empty[qb];
      dq = make_sdesc_sw128(smem_u32(sQ + (qb * ATT_NQ + t) * ATT_Q_BYTES));
      mbar_wait(&q_full[qb], (qi / ATT_QBUF) & 1);
      ++qi;
      tc_fence_after();
      if (((w.active >> t) & 1) == 0) {
        // idle tile of a short unit: release Q and every ring stage unread
        if (lane == 0) mbar_arrive(q_empty_u);
        for (int j = 0; j < w.n_tiles; ++j) {
          mbar_wait(&k_full[ks], kph);
          if (lane == 0) mbar_arrive(&k_empty[ks]);
          if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
          mbar_wait(&v_full[vs], vph);
          if (lane == 0) mbar_arrive(&v_empty[vs]);
          if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
        }
        __syncwarp();
        continue;
      }
      // prologue: S_t(0); the softmax warps must have pulled the last S_t of the previous unit
      mbar_wait(&k_full[ks], kph);
      if (g > 0) mbar_wait(&s_free[t], (g - 1) & 1);
      tc_fence_after();
      issue_qk(smem_u32(sK + ks * ATT_KV_BYTES));
      if (lane == 0) {
        umma_commit(&k_empty[ks]);
        if (w.n_tiles == 1) umma_commit(q_empty_u);
      }
      __syncwarp();
      if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
      for (int j = 0; j < w.n_tiles; ++j) {
        const uint32_t ph = (g + j) & 1;
        if (j + 1 < w.n_tiles) {
          // S_t(j+1) as soon as softmax has pulled S_t(j) into registers
          mbar_wait(&k_full[ks], kph);
          mbar_wait(&s_free[t], ph);
          tc_fence_after();
          issue_qk(smem_u32(sK + ks * ATT_KV_BYTES));
          if (lane == 0) {
            umma_commit(&k_empty[ks]);
            if (j + 2 == w.n_tiles) umma_commit(q_empty_u);  // last read of Q_t in this unit
          }
          __syncwarp();
          if (++ks == ATT_KS) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs], vph);
        mbar_wait(&p_full[t], ph);
        tc_fence_after();
        issue_pv(smem_u32(sV + vs * ATT_KV_BYTES), j);
        if (lane == 0) umma_commit(&v_empty[vs]);
        __syncwarp();
        if (++vs == ATT_VS) { vs = 0; vph ^= 1; }
      }
      g += uint32_t(w.n_tiles);
    }
  } else {
    // ------------------------------------------------------------ softmax + output
    const int t = warp >> 2;  // query tile of this warpgroup
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = uint32_t(qd * 32) << 16;
    const uint32_t t_s = tS + lane_off + t * ATT_BN;
    const uint32_t t_o = tO + lane_off + t * ATT_D;
    const uint32_t t_p = tP + lane_off + t * (ATT_BN / 2);
    const float sc = a.scale_log2;
    uint32_t g = 0;  // key steps this tile has run in earlier units
#ifdef ATT_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
    long long n_steps = 0, n_units = 0, wait_s_first = 0, wait_o_max = 0;
#endif
    for (uint32_t it = 0;; ++it) {
      const int u = take_unit(it);
      if (u < 0) break;
      const AttnUnit w = load_unit<EXTRAS>(a, u);
      if (((w.active >> t) & 1) == 0) continue;
      float m_run = -INFINITY, l_run = 0.f;

      for (int j = 0; j < w.n_tiles; ++j) {
        const bool inA = j < w.nA;
        const int n_valid = inA ? min(ATT_BN, w.ka_len - j * ATT_BN)
                                : min(ATT_BN, w.kb_len - (j - w.nA) * ATT_BN);
        const uint32_t ph = (g + j) & 1;
        ATT_T(c0);
        mbar_wait(&s_full[t], ph);
        // Optional start offset between the groups (groups that exponentiate at the same time also
        // leave the MUFU pipe idle at the same time). It gave +2 % with one unit per CTA; in the
        // persistent kernel the tiles drift on their own and it measures as noise (0 / 400 / 800
        // cycles: 731 / 733 / 731 TFLOP/s), so it is off.
        if (ATT_SKEW > 0 && g == 0 && j == 0 && t > 0) {
          const long long until = clock64() + t * ATT_SKEW;
          while (clock64() < until) {
          }
        }
        tc_fence_after();
        ATT_T(c1);
        float s[ATT_BN];
#pragma unroll
        for (int c = 0; c < ATT_BN / 32; ++c)
          tmem_ld32(t_s + c * 32, reinterpret_cast<uint32_t*>(s) + c * 32);
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
        if (EXTRAS == 2 && (a.causal != 0 || a.rel_bias != nullptr)) {
          // text-encoder variants: T5 relative position bias and / or CLIP's causal mask
          const int qpos = w.q_pos0 + t * ATT_BM + r;
          const int kpos0 = inA ? j * ATT_BN : w.ka_len + (j - w.nA) * ATT_BN;
          if (a.rel_bias != nullptr && t * ATT_BM + r < w.q_rows) {
            const float* bh = a.rel_bias + size_t(w.head) * a.rel_ld + (kpos0 - qpos + a.rel_len - 1);
#pragma unroll
            for (int i = 0; i < ATT_BN; ++i)
              if (i < n_valid) s[i] += __ldg(bh + i);
          }
          if (a.causal != 0) {
#pragma unroll
            for (int i = 0; i < ATT_BN; ++i)
              if (kpos0 + i > qpos) s[i] = -INFINITY;
          }
        }
        float alpha = 1.f;
        bool resc = false;
        if constexpr (!BOUNDED) {
          // 8 independent chains (one long fmax chain would expose its full latency)
          float mx8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) mx8[i] = s[i];
#pragma unroll
          for (int i = 8; i < ATT_BN; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], s[i]);
          const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                                 fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          const float m_new = fmaxf(m_run, mx);
          // lazy rescale: only move the reference max when it grew by more than 2^8
          resc = (m_new - m_run) * sc > ATT_RESCALE_THRESHOLD;
          if (resc) {
            alpha = fast_exp2((m_run - m_new) * sc);
            m_run = m_new;
          }
        }
        const float m_sc = BOUNDED ? 0.f : m_run * sc;  // BOUNDED: exp2(s * sc) as it is
        ATT_T(c3);
        uint32_t pk[ATT_BN / 2];
#if ATT_F32X2
        // Packed fp32 pairs (FFMA2 / FADD2, sm_100): the scale-and-shift and the row sum take one
        // issue slot per two elements; the warps are issue-bound next to the MUFU pipe.
        float2* s2 = reinterpret_cast<float2*>(s);
        const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_sc, -m_sc);
#pragma unroll
        for (int i = 0; i < ATT_BN / 2; ++i) {
          // (folding the scale into the queries would save this FFMA2: measured +0.4 %, not done)
          const float2 x = __ffma2_rn(s2[i], sc2, nm2);
          // With the running maximum the softmax warps are bound by their dependent chain and freeing MUFU
          // slots buys nothing (+0.5 %); without it (BOUNDED) the chain is 11 % shorter, the
          // MUFU pipe is the tighter bound again and every 5th pair on the FMA pipe gives 766 -> 820 TFLOP/s
          // (every 6th 809, 4th 788-791, 3rd 786, 2nd 732: profiles/r02_attn_variants.txt).
          constexpr int kPoly = BOUNDED ? ATT_POLY_BOUNDED : ATT_POLY_MOD;
          if (kPoly > 0 && i % (kPoly > 0 ? kPoly : 1) == 0) { s2[i] = poly_exp2_pair(x); continue; }
          s2[i] = make_float2(fast_exp2(x.x), fast_exp2(x.y));
        }
        float2 acc2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc2[i] = s2[i];
#pragma unroll
        for (int i = 4; i < ATT_BN / 2; ++i) acc2[i & 3] = __fadd2_rn(acc2[i & 3], s2[i]);
#pragma unroll
        for (int i = 0; i < ATT_BN / 2; ++i) pk[i] = pack_bf16x2(s2[i].x, s2[i].y);
        const float2 acc = __fadd2_rn(__fadd2_rn(acc2[0], acc2[1]), __fadd2_rn(acc2[2], acc2[3]));
        const float sum = acc.x + acc.y;
#else
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < ATT_BN; ++i) {
          s[i] = fast_exp2(fmaf(s[i], sc, -m_sc));
        }
#pragma unroll
        for (int i = 0; i < ATT_BN; i += 2) {
          sum4[(i >> 1) & 3] += s[i] + s[i + 1];
          pk[i >> 1] = pack_bf16x2(s[i], s[i + 1]);
        }
        const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
#endif
        l_run = l_run * alpha + sum;
        ATT_T(c4);
        if (j > 0) {
          // PV_t(j-1) must be complete before P_t is overwritten and before O_t is rescaled
          mbar_wait(&o_full[t], ph ^ 1);
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
#pragma unroll
        for (int c = 0; c < ATT_BN / 64; ++c) tmem_st32(t_p + c * 32, pk + c * 32);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
        ATT_T(c6);
        ATT_ACC(0, c0, c1); ATT_ACC(1, c1, c2); ATT_ACC(2, c2, c3); ATT_ACC(3, c3, c4);
        ATT_ACC(4, c4, c5); ATT_ACC(5, c5, c6);
#ifdef ATT_TIMING
        if (j == 0) wait_s_first += c1 - c0;   // the unit-boundary share of wait_s
        if (c5 - c4 > wait_o_max) wait_o_max = c5 - c4;
#endif
      }
      g += uint32_t(w.n_tiles);
#ifdef ATT_TIMING
      n_steps += w.n_tiles;
      ++n_units;
#endif
      // Output of the unit. The first P V of the next unit overwrites O_t, but it needs p_full of
      // that unit, which these warps only give after this read has completed.
      ATT_T(e0);
      mbar_wait(&o_full[t], (g - 1) & 1);
      tc_fence_after();
      uint32_t o[ATT_D];
      tmem_ld32(t_o, o);
      tmem_ld32(t_o + 32, o + 32);
      tmem_wait_ld();
      tc_fence_before();
      if (t * ATT_BM + r < w.q_rows) {
        const float inv = 1.f / l_run;
        __nv_bfloat16* op = a.out[w.q_seg] +
                            size_t(w.q_row0 + t * ATT_BM + r) * a.ldo[w.q_seg] +
                            a.o_col[w.q_seg] + w.head * ATT_D;
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
      ATT_T(e1);
      ATT_ACC(6, e0, e1);
    }
#ifdef ATT_TIMING
    if (a.dbg != nullptr && lane == 0 && qd == 2 && t < 2) {
      long long* d = a.dbg + size_t(blockIdx.x) * 32 + t * 16;
      for (int i = 0; i < 7; ++i) d[i] = tacc[i];
      d[7] = n_steps;
      d[8] = clock64() - t_begin;
      d[9] = wait_s_first;
      d[10] = n_units;
      d[11] = wait_o_max;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
  if (threadIdx.x == 0) {
    // the last CTA to finish re-arms the unit counter for the next launch (graph replays included)
    __threadfence();
    if (atomicInc(reinterpret_cast<unsigned*>(a.sched) + 1, gridDim.x - 1) == gridDim.x - 1) a.sched[0] = 0;
  }
}

static int make_rows_map(CUtensorMap* m, const void* base, int rows, int cols, int ld) {
  if (base == nullptr) return B200_ERR_INVALID;
  uint64_t d[2] = {uint64_t(cols), uint64_t(rows)}, s[1] = {uint64_t(ld) * 2};
  uint32_t b[2] = {ATT_D, ATT_BOX_ROWS};
  return get_tmap_bf16_sw128(m, base, 2, d, s, b);
}

}  // namespace b200

using namespace b200;

#ifdef ATT_TIMING
long long* g_att_dbg = nullptr;
extern "C" long long* b200_attn_debug_buffer(void) { return g_att_dbg; }
#endif

extern "C" int b200_attn_rows_per_item(void) { return ATT_NQ * ATT_BM; }
extern "C" long long b200_attn_workspace_bytes(void) { return 2 * sizeof(int32_t); }  // sched_state

// Host-side work list of the persistent kernel. Units = (query block of ATT_NQ tiles, head),
// sorted longest first by the softmax time the kernel is bound by: key steps x max(active tiles x
// MUFU time per tile-step, length of one warp's dependent chain). The kernel hands them out in
// this order (longest-processing-time-first list scheduling).
extern "C" int b200_attn_build_schedule(const int32_t* seq_table, int n_seq, int n_heads,
                                        int32_t* work_units, int* n_units_out) {
  if (!seq_table || n_seq <= 0 || n_heads <= 0 || !n_units_out) return B200_ERR_INVALID;
  struct Block { long cost; int seq, seg, off; };
  std::vector<Block> blocks;
  for (int i = 0; i < n_seq; ++i) {
    const int32_t* s = seq_table + i * 8;
    const long steps = (s[5] + ATT_BN - 1) / ATT_BN + (s[7] + ATT_BN - 1) / ATT_BN;
    for (int seg = 0; seg < 2; ++seg)
      for (int off = 0; off < s[2 * seg + 1]; off += ATT_NQ * ATT_BM) {
        const int rows = std::min(ATT_NQ * ATT_BM, s[2 * seg + 1] - off);
        const long nq = (rows + ATT_BM - 1) / ATT_BM;
        blocks.push_back({steps * std::max(nq * 560L, 900L), i, seg, off});
      }
  }
  std::stable_sort(blocks.begin(), blocks.end(),
                   [](const Block& x, const Block& y) { return x.cost > y.cost; });
  const long n_units = long(blocks.size()) * n_heads;
  if (n_units <= 0 || n_units > (1L << 28)) return B200_ERR_INVALID;
  *n_units_out = int(n_units);
  if (!work_units) return B200_OK;  // size query
  int32_t* w = work_units;
  for (const Block& bl : blocks)
    for (int h = 0; h < n_heads; ++h, w += 4) {
      w[0] = bl.seq; w[1] = bl.seg; w[2] = bl.off; w[3] = h;
    }
  return B200_OK;
}

extern "C" int b200_attn_varlen_bf16(const B200AttnSource* src_a, const B200AttnSource* src_b,
                                     const int32_t* seq_table, const int32_t* work_units,
                                     int n_units, int32_t* sched_state, int max_ctas,
                                     float softmax_scale, void* stream_) {
  return b200_attn_varlen_ex(src_a, src_b, seq_table, work_units, n_units, sched_state, max_ctas,
                             softmax_scale, nullptr, stream_);
}

extern "C" int b200_attn_varlen_ex(const B200AttnSource* src_a, const B200AttnSource* src_b,
                                   const int32_t* seq_table, const int32_t* work_units,
                                   int n_units, int32_t* sched_state, int max_ctas,
                                   float softmax_scale, const B200AttnExtra* extra, void* stream_) {
  if (!src_a || !seq_table || !work_units || !sched_state || n_units <= 0 || max_ctas < 0)
    return B200_ERR_INVALID;
  if (extra && extra->rel_bias && (extra->rel_len <= 0 || extra->rel_ld < 2 * extra->rel_len - 1))
    return B200_ERR_INVALID;
  const int n_ctas = std::min(n_units, max_ctas > 0 ? max_ctas : device_sm_count());
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
  a.work_units = work_units;
  a.sched = sched_state;
  a.n_units = n_units;
  a.scale_log2 = softmax_scale * 1.4426950408889634f;
  a.causal = extra ? extra->causal : 0;
  a.rel_bias = extra ? extra->rel_bias : nullptr;
  a.rel_len = extra ? extra->rel_len : 0;
  a.rel_ld = extra ? extra->rel_ld : 0;
  a.q_mask = extra ? extra->q_mask : nullptr;
  a.q_mask_shift = extra ? extra->q_mask_shift : 0;
  if (a.q_mask && a.q_mask_shift < 7) return B200_ERR_INVALID;  // a query tile is 128 rows
  a.dbg = nullptr;
#ifdef ATT_TIMING
  {
    static long long* dbg_buf = nullptr;
    if (!dbg_buf) cudaMalloc(&dbg_buf, size_t(1) << 26);
    a.dbg = dbg_buf;
    cudaMemsetAsync(dbg_buf, 0, (size_t(n_ctas) * 32 + 1024) * 8, reinterpret_cast<cudaStream_t>(stream_));
    extern long long* g_att_dbg;
    g_att_dbg = dbg_buf;
  }
#endif
  const int extras = (a.causal != 0 || a.rel_bias != nullptr) ? 2 : (a.q_mask != nullptr ? 1 : 0);
  // bounded logits (see AttnArgs): the denoising instantiation without the row-maximum pass
  static const bool no_bounded = []() { const char* v = getenv("SDUSS_B200_NO_BOUNDED"); return v && v[0] == '1'; }();
  // (also with the patch cache's query-tile mask: a cached step with every patch flagged stays
  //  bit-identical to the uncached one)
  const bool bounded = extras <= 1 && extra != nullptr && extra->bounded_logits != 0 && !no_bounded;
  auto kern = extras == 2 ? attn_fwd_kernel<2>
            : extras == 1 ? (bounded ? attn_fwd_kernel<1, true> : attn_fwd_kernel<1>)
                          : (bounded ? attn_fwd_kernel<0, true> : attn_fwd_kernel<0>);
  static unsigned long long configured[5] = {0, 0, 0, 0, 0};
  if (int rc = ensure_dynamic_smem(kern, ATT_SMEM, &configured[bounded ? 3 + extras : extras])) return rc;
  dim3 grid(n_ctas);
  return launch_pdl(kern, grid, dim3(ATT_THREADS), ATT_SMEM,
                    reinterpret_cast<cudaStream_t>(stream_), tm[0][0], tm[1][0], tm[0][1], tm[1][1],
                    tm[0][2], tm[1][2], a);
}
