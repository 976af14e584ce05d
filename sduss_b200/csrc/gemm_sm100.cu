// C[M,N] = epilogue(A[M,K] * W[N,K]^T) in bf16 with fp32 accumulation on the sm_100a tensor
// cores. Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer,
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> fused epilogue -> HBM).
// Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the main loop
// of tile i+1. Replaces the F.linear call sites of the reference's denoising step
// (sduss/model_executor/modules/resnet.py:163, attention.py:73-96,259-274,411 and the
// diffusers FeedForward / AdaLayerNorm linears they wrap).
#include "../../include/sduss_b200.h"
#include "epilogue.cuh"
#include <cstdlib>

#include "host_util.h"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;
#ifndef GEMM_STAGES_128
#define GEMM_STAGES_128 5
#endif

// TWO: the CTA pair runs ONE tcgen05.mma.cta_group::2 of M = 256 per K step; each CTA keeps only its
// half of the W tile, so a stage is 16 KB + BN x 64 B and the ring holds 6-8 K steps instead of 4-5
// (these main loops are bound by operand latency x bytes in flight, not by the tensor pipe).
template <int BN, int EPI, bool TWO = false>
struct GemmCfg {
  // The residual chunk of a gate*y + residual epilogue is TMA-loaded INTO the output staging tile
  // and updated in place, so that epilogue costs no pipeline stage (a 3-stage 128 x 256 main loop
  // lost 27 % on the K = 1536 out-projections: operand latency, not the tensor pipe).
  static constexpr int kStages = TWO ? ((BN == 128) ? 8 : 6) : (BN == 256) ? 4 : (BN == 192) ? 4 : GEMM_STAGES_128;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // + 2 staging tiles (16 KB each) for the TMA epilogue
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 2 * EPI_STAGE_BYTES + 256 /*barriers*/ +
      2048 /*column sums + fp32 bias of a folded LayerNorm, one tile: [2][256] floats*/;
  // (no alignment slack: the dynamic shared memory window is declared 1024-byte aligned and the
  //  kernel traps if it is not -- 4 x 48 KB stages + staging + the two tables fill the 227 KB)
  static constexpr int kTmemCols = BN == 128 ? 256 : 512;  // two accumulator stages (power of two)
};

struct GemmShape {
  int M, N, K;
};

#ifdef GEMM_TRACE
// tools/gemm_trace.py: per launch (ring of 64) and for CTA 0, clock64 / globaltimer stamps of the phases
__device__ unsigned long long g_trace[64][10];
__device__ unsigned int g_trace_launch;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot) do { if (blockIdx.x == 0) g_trace[trace_row][slot] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif

// MC = 2: CTA pairs (cluster of 2 along M) share every W tile: each CTA fetches half of it and
// multicasts it into both CTAs' shared memory, cutting L2 -> SM operand traffic by a third
// (the 128x256x64 main loop is L2-bandwidth bound: 96 B/clk/SM of operand loads on 148 SMs).
template <int BN, int EPI, int MC, bool TWO>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
                 GemmShape s, EpiArgs e) {
  static_assert(MC == 1 || MC == 2 || (MC == 4 && TWO), "clusters: 1, a pair, or two cta_group::2 pairs");
  static_assert(!TWO || MC >= 2, "cta_group::2 needs the CTA pair");
  using Cfg = GemmCfg<BN, EPI, TWO>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();  // SW128 tiles need 1024-byte alignment (fail loudly)
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* stageC = smem + Cfg::kStages * Cfg::kStageBytes;  // [2][16 KB] residual in / output out
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageC + 2 * EPI_STAGE_BYTES);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint64_t* rfull = tempty + 2;                // [2] residual chunk landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2);
  float* colsum_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef GEMM_TRACE
  __shared__ unsigned int trace_row_s;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    trace_row_s = atomicAdd(&g_trace_launch, 1u) & 63u;
    g_trace[trace_row_s][0] = clock64();
    g_trace[trace_row_s][8] = gtimer();
  }
  __syncthreads();
  const unsigned int trace_row = blockIdx.x == 0 ? trace_row_s : 0;
#endif

  const int tiles_n = (s.N + BN - 1) / BN;
  // MC == 2: tiles are enumerated per CTA pair; rank r of the pair owns M tile 2 * pair + r
  // (a pair's second tile may lie beyond M: TMA zero-fills its loads and clips its stores).
  const int tiles_m = (s.M + BM * MC - 1) / (BM * MC);
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (s.K + BK - 1) / BK;
  // MC == 4 (TWO only): two pairs stacked along M share every W half: CTA (pair p, half r) fetches
  // quarter p of half r and multicasts it to the CTA holding half r in both pairs -- 16 + BN/4 x 128 B
  // per CTA and K step instead of 16 + BN/2 x 128 B through L2 -> SM, the bound of these main loops.
  const uint32_t rank = MC >= 2 ? cluster_ctarank() : 0;
  const int first_tile = int(blockIdx.x) / MC;
  const int tile_step = int(gridDim.x) / MC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      // released by the MMA warp of every CTA the stage is multicast to / by the pair's one MMA warp
      mbar_init(&empty[i], TWO ? MC / 2 : MC);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], TWO ? 8 : 4);  // TWO: the epilogue warps of BOTH CTAs release the leader's
      mbar_init(&rfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (TWO) {
      tmem_alloc2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (MC >= 2) cluster_sync_all();  // peer barriers are initialised before any multicast
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TRACE(1);
  pdl_launch_dependents();
  // bytes one phase of full[stage] waits for: TWO = both CTAs' A tiles and W halves, on the leader
  constexpr uint32_t kExpect = TWO ? 2u * Cfg::kStageBytes : uint32_t(Cfg::kStageBytes);
  auto load_w = [&](int stage, int kb, int n0) {
    if constexpr (TWO && MC == 4) {
      const int r = int(rank & 1), p = int(rank >> 1);
      tma_load_2d_2sm_mc(smemB + stage * Cfg::kBBytes + p * (Cfg::kBBytes / 2), &tmB, &full[stage], kb * BK,
                         n0 + r * (BN / 2) + p * (BN / 4), uint16_t(5u << r));
    } else if constexpr (TWO) {
      // my half of the W tile stays in MY shared memory; the bytes count on the leader's barrier
      tma_load_2d_2sm(smemB + stage * Cfg::kBBytes, &tmB, &full[stage], kb * BK, n0 + int(rank) * (BN / 2));
    } else if constexpr (MC == 2) {
      // my half of the W tile (BN/2 rows), delivered to both CTAs of the pair
      tma_load_2d_mc(smemB + stage * Cfg::kBBytes + rank * (Cfg::kBBytes / 2), &tmB, &full[stage],
                     kb * BK, n0 + int(rank) * (BN / 2), uint16_t(3));
    } else {
      tma_load_2d(smemB + stage * Cfg::kBBytes, &tmB, &full[stage], kb * BK, n0);
    }
  };
  // W holds weights (w_static: nothing on the stream writes it): the first ring stages of W are
  // requested BEFORE waiting for the previous kernel, so their HBM latency runs under its tail;
  // only the A loads (L2 hits: A was just written) remain after the wait.
  int npre = 0;
  if ((warp == 0 || warp == 3) && e.w_static && e.row_mask == nullptr && first_tile < num_tiles)
    npre = num_kb < Cfg::kStages ? num_kb : Cfg::kStages;
  if (warp == 0 && npre > 0) {
    if (lane == 0) {
      const int n0 = (first_tile % tiles_n) * BN;
      for (int kb = 0; kb < npre; ++kb) {
        if (!TWO || (rank & 1) == 0) mbar_expect_tx(&full[kb], kExpect);
        load_w(kb, kb, n0);
      }
    }
    __syncwarp();
  }
  pdl_wait();  // the prologue above overlapped the previous kernel; its outputs are visible now
  if (threadIdx.x == 0) TRACE(2);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const int m0 = ((t / tiles_n) * MC + int(rank)) * BM;
      const int n0 = (t % tiles_n) * BN;
      if (tile_skipped(e, m0 - int(rank) * BM, s.M)) continue;  // clean patch (same answer in both CTAs of a pair)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
          const bool pre = t == first_tile && kb < npre;  // W of this stage is already on its way
          if (!pre && (!TWO || (rank & 1) == 0)) mbar_expect_tx(&full[stage], kExpect);
          if constexpr (TWO) tma_load_2d_2sm(smemA + stage * Cfg::kABytes, &tmA, &full[stage], kb * BK, m0);
          else tma_load_2d(smemA + stage * Cfg::kABytes, &tmA, &full[stage], kb * BK, m0);
#ifndef GEMM_SPLIT_PRODUCER
          if (!pre) load_w(stage, kb, n0);
#endif
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
#ifdef GEMM_SPLIT_PRODUCER
  } else if (warp == 3) {
    // ------------------------------------------------------------ second TMA producer: the W tiles
    // (experiment: is one thread issuing both loads of a K step the serial bottleneck?)
    int stage = 0;
    uint32_t phase = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const int m0 = ((t / tiles_n) * MC + int(rank)) * BM;
      const int n0 = (t % tiles_n) * BN;
      if (tile_skipped(e, m0 - int(rank) * BM, s.M)) continue;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0 && !(t == first_tile && kb < npre)) load_w(stage, kb, n0);
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
#endif
  } else if (warp == 1 && (!TWO || (rank & 1) == 0)) {
    // ------------------------------------------------------------ MMA issuer (TWO: the leader's only)
    constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * BM : BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      if (tile_skipped(e, (t / tiles_n) * MC * BM, s.M)) continue;
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0 && t == first_tile && kb == 0) TRACE(3);
        if (lane == 0 && t == first_tile && kb == num_kb - 1) TRACE(4);
        if (lane == 0) {
          const uint64_t da = make_sdesc_sw128(smem_u32(smemA + stage * Cfg::kABytes));
          const uint64_t db = make_sdesc_sw128(smem_u32(smemB + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes (encoded >>4) per 16-element K step inside the 128B swizzle span
            if constexpr (TWO) umma_ss2(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (TWO) umma_commit2_mc(&empty[stage], uint16_t(MC == 4 ? 0xF : 3));  // every CTA that writes or is written
          else if constexpr (MC == 2) umma_commit_mc(&empty[stage], uint16_t(3));
          else umma_commit(&empty[stage]);
          if (kb == num_kb - 1) {
            if constexpr (TWO) umma_commit2_mc(&tfull[acc], uint16_t(3u << (rank & 2)));  // both CTAs' epilogues
            else umma_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      ++it;  // accumulator stages count the tiles actually computed
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int r = q * 32 + lane;          // row inside the tile == TMEM lane
    const bool leader = threadIdx.x == 128;
    const bool has_resid = (EPI == EPI_GATE_RESID) && e.resid != nullptr && !e.out_fp32;
    constexpr int NCH = BN / 64;
    uint32_t ruse0 = 0, ruse1 = 0;
    // Staging buffers alternate over ALL chunks this CTA stores, not per tile: a tile with an odd
    // number of chunks (N tail, N < BN) would otherwise be followed by a chunk that reuses the
    // buffer whose TMA store is still in flight (the store is only waited for after the next
    // chunk has been staged). Seen as corrupted tail columns on short-K GEMMs (conv_in: K = 64).
    uint32_t chunk_ctr = 0;
    int it = 0;
    if (leader) {
      tma_prefetch_desc(&tmC);
      if (has_resid) tma_prefetch_desc(&tmR);
    }
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const int acc = it & 1;
      const int m0 = ((t / tiles_n) * MC + int(rank)) * BM;
      const int n0 = (t % tiles_n) * BN;
      if (tile_skipped(e, m0 - int(rank) * BM, s.M)) continue;
      if (has_resid && leader) {  // residual chunks 0 and 1 fly while the main loop finishes
        tma_store_wait_read<0>();   // both staging tiles have left for HBM (previous tile)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < NCH && n0 + c * 64 < s.N) {
            const int rb = (chunk_ctr + c) & 1;
            mbar_expect_tx(&rfull[rb], EPI_STAGE_BYTES);
            tma_load_2d(stageC + rb * EPI_STAGE_BYTES, &tmR, &rfull[rb], n0 + c * 64, m0);
          }
        }
      }
      __syncwarp();
      const int row = m0 + r;
      // LayerNorm folded into this GEMM: the row's (mean, rstd) and the tile's column sums are
      // fetched once per tile while the main loop is still running (a global load inside the
      // per-chunk chain costs these epilogue-bound GEMMs more than the LayerNorm launch it saves)
      const float2 ln_mr = ln_row_stats(e, row, row < s.M, s.K);
      const bool tables = e.ln_colsum != nullptr || e.bias != nullptr;
      if (tables) {
        named_barrier<1, 128>();  // every thread is done with the previous tile's tables
        for (int j = int(threadIdx.x) - 128; j < BN; j += 128) {
          const bool ok = n0 + j < s.N;
          colsum_s[j] = ok && e.ln_colsum != nullptr ? e.ln_colsum[n0 + j] : 0.f;
          colsum_s[256 + j] = ok && e.bias != nullptr ? __bfloat162float(e.bias[n0 + j]) : 0.f;  // fp32 bias
        }
        named_barrier<1, 128>();
      }
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (leader && t == first_tile) TRACE(5);
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int nc = n0 + c * 64;
        if (nc >= s.N) break;
        const int b = (chunk_ctr + c) & 1;
        if (has_resid && c >= 1 && c + 1 < NCH && nc + 64 < s.N) {
          // residual of chunk c + 1 -> the tile chunk c - 1 was stored from, once that store has
          // read it; it lands while this chunk is processed
          if (leader) {
            tma_store_wait_read<0>();
            mbar_expect_tx(&rfull[b ^ 1], EPI_STAGE_BYTES);
            tma_load_2d(stageC + (b ^ 1) * EPI_STAGE_BYTES, &tmR, &rfull[b ^ 1], nc + 64, m0);
          }
          __syncwarp();
        }
        float v[64];
        tmem_ld32(t_row + c * 64, reinterpret_cast<uint32_t*>(v));
        tmem_ld32(t_row + c * 64 + 32, reinterpret_cast<uint32_t*>(v) + 32);
        tmem_wait_ld();
        if (e.out_fp32) {  // fp32 outputs keep the direct store path
          epilogue_chunk64<EPI>(e, v, row, nc, s.M, s.N);
          continue;
        }
        if (has_resid) {
          mbar_wait(&rfull[b], (b ? ruse1 : ruse0) & 1);
          if (b) ++ruse1; else ++ruse0;
        }
        epilogue_math64<EPI>(e, v, row, row < s.M, nc, s.N, has_resid ? stageC + b * EPI_STAGE_BYTES : nullptr, r, ln_mr,
                             tables ? colsum_s + c * 64 : nullptr);
        epilogue_stage64<EPI>(stageC + b * EPI_STAGE_BYTES, v, r);
        fence_proxy_async();                       // smem writes -> visible to the TMA engine
        if (leader) tma_store_wait_read<0>();      // chunk c-1 has left its staging tile
        named_barrier<1, 128>();
        if (leader) {
          tma_store_2d(&tmC, stageC + b * EPI_STAGE_BYTES, EPI == EPI_GEGLU ? (nc >> 1) : nc, m0);
          tma_store_commit();
        }
      }
      {
        const int left = (s.N - n0 + 63) / 64;
        chunk_ctr += uint32_t(left < NCH ? left : NCH);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) mbar_arrive_leader(&tempty[acc]);
        else mbar_arrive(&tempty[acc]);
      }
      ++it;
    }
    if (leader) tma_store_wait<0>();
    if (leader) TRACE(6);
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC >= 2) cluster_sync_all();  // the peer may still arrive on my barriers until here
  if (warp == 2) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
#ifdef GEMM_TRACE
  if (threadIdx.x == 64 && blockIdx.x == 0) { g_trace[trace_row][7] = clock64(); g_trace[trace_row][9] = gtimer(); }
#endif
}

template <int BN, int EPI, int MC, bool TWO = false>
static int launch_gemm_mc(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const CUtensorMap& tmR, GemmShape s, const EpiArgs& e, int num_sms,
                          cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI, TWO>;
  auto kern = gemm_bf16_kernel<BN, EPI, MC, TWO>;
  static unsigned long long configured = 0;
  if (int rc = ensure_dynamic_smem(kern, Cfg::kSmemBytes, &configured)) return rc;
  const int tiles = ((s.M + BM * MC - 1) / (BM * MC)) * ((s.N + BN - 1) / BN);
  int slots = num_sms / MC;
  if constexpr (MC == 4) {  // clusters of 4 do not pack all GPCs: ask how many are co-resident
    static int quads = 0;
    if (quads == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(num_sms / 4 * 4);
      cfg.blockDim = dim3(GEMM_THREADS);
      cfg.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) n = num_sms / 4 - 4;
      quads = n;
    }
    slots = quads;
  }
  const int grid = (tiles < slots ? tiles : slots) * MC;
  return launch_pdl_cluster(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, stream, MC, tmA, tmB,
                            tmC, tmR, s, e);
}

// mode: 0 one CTA per tile, 1 CTA pairs with W multicast (two M = 128 MMAs, round 1), 2 CTA pairs as
// cta_group::2 (one M = 256 MMA, W halves never duplicated in shared memory), 3 two such pairs
// sharing every W half (cluster of 4)
template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       const CUtensorMap& tmR, GemmShape s, const EpiArgs& e, int num_sms,
                       cudaStream_t stream, int mode) {
  switch (mode) {
    case 3: return launch_gemm_mc<BN, EPI, 4, true>(tmA, tmB, tmC, tmR, s, e, num_sms, stream);
    case 2: return launch_gemm_mc<BN, EPI, 2, true>(tmA, tmB, tmC, tmR, s, e, num_sms, stream);
    case 1: return launch_gemm_mc<BN, EPI, 2>(tmA, tmB, tmC, tmR, s, e, num_sms, stream);
    default: return launch_gemm_mc<BN, EPI, 1>(tmA, tmB, tmC, tmR, s, e, num_sms, stream);
  }
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB,
                        const CUtensorMap& tmC, const CUtensorMap& tmR, GemmShape s,
                        const EpiArgs& e, int num_sms, cudaStream_t stream, int mode) {
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BN, EPI_BIAS>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    case EPI_GELU_TANH: return launch_gemm<BN, EPI_GELU_TANH>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    case EPI_GATE_RESID: return launch_gemm<BN, EPI_GATE_RESID>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    case EPI_QK_RMSNORM: return launch_gemm<BN, EPI_QK_RMSNORM>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    case EPI_GEGLU: return launch_gemm<BN, EPI_GEGLU>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    case EPI_ROWVEC: return launch_gemm<BN, EPI_ROWVEC>(tmA, tmB, tmC, tmR, s, e, num_sms, stream, mode);
    default: return B200_ERR_INVALID;
  }
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}

}  // namespace b200

using namespace b200;

// See include/sduss_b200.h for the contract.
extern "C" int b200_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                              int epi_mode, const B200EpilogueDesc* ep, void* stream_) {
  if (!A || !W || !ep || !ep->C || M <= 0 || N <= 0 || K <= 0) return B200_ERR_INVALID;
  if ((N & 7) || (K & 7) || (lda & 7) || (ldw & 7) || (ep->ldc & 7)) return B200_ERR_INVALID;
  if (epi_mode == EPI_QK_RMSNORM && ((N & 63) || !ep->rms_wq || !ep->rms_wk)) return B200_ERR_INVALID;
  if (epi_mode == EPI_GEGLU && (N & 63)) return B200_ERR_INVALID;
  if ((epi_mode == EPI_ROWVEC && (!ep->rowvec || !ep->row_group)) ||
      (epi_mode == EPI_GATE_RESID && ep->gate && !ep->row_group))
    return B200_ERR_INVALID;
  const int sms = device_sm_count();
  if (sms <= 0) return B200_ERR_DRIVER;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);

  // 128 x 256 or 128 x 128 tiles: whichever the operand-traffic / occupancy model predicts faster
  static const bool mc_allowed0 = []() { const char* v = getenv("SDUSS_B200_NO_MULTICAST"); return !(v && v[0] == '1'); }();
  int BN = choose_tile_n((M + BM - 1) / BM, N, (K + BK - 1) / BK, sms, mc_allowed0 && M > BM);
  // One-round problems that leave a quarter of the SMs idle with 256-wide tiles (SDXL level 2:
  // M = 2560, N = 1280 -> 100 tiles) run as 128 x 192 tiles when those still fit one round (140
  // tiles): every CTA then streams 25 % fewer operand bytes through its smem window, which is
  // what bounds a single-tile CTA (load latency x bytes / window, not the tensor pipe).
  {
    static const bool no192 = []() { const char* v = getenv("SDUSS_B200_NO_BN192"); return v && v[0] == '1'; }();
    const long mt = (M + BM - 1) / BM;
    const long t256 = mt * ((N + 255) / 256), t192 = mt * ((N + 191) / 192);
    if (!no192 && BN == 256 && t256 * 4 <= 3L * sms && t192 <= sms && t192 > t256) BN = 192;
  }
  if (const char* fb = getenv("SDUSS_B200_FORCE_BN")) {  // experiment hook (tools/bench_gemm_bn.py), read per call
    const int v = atoi(fb);
    if ((v == 128 || v == 192 || v == 256) && N >= v) BN = v;
  }

  // pair CTAs (W-tile multicast) whenever there are at least two M tiles and enough pairs of
  // tiles to occupy the 74 SM pairs
  static const bool mc_allowed = []() { const char* v = getenv("SDUSS_B200_NO_MULTICAST"); return !(v && v[0] == '1'); }();
  const long pair_tiles = long((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const char* mp = getenv("SDUSS_B200_MC_MIN_PAIRS");  // experiment hook (read per call)
  const long min_pairs = mp ? atol(mp) : (sms / 2) / 2;
  const bool multicast = mc_allowed && M > BM && pair_tiles >= min_pairs;
  // SDUSS_B200_NO_2CTA=1 / SDUSS_B200_QUAD=1 are read per call: A/B runs flip them inside one process
  const char* no2 = getenv("SDUSS_B200_NO_2CTA");
  const char* qd = getenv("SDUSS_B200_QUAD");
  // cta_group::2 pays above one round of pair tiles (deeper ring, half the W bytes in shared memory);
  // a single round gains nothing from it and pays ~1 us more prologue + teardown (TMEM alloc / release
  // across the pair): those keep the W-multicast pairs (in-process A/B, profiles/r02_gemm_ab.txt)
  const long tiles_two = long((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const bool two = multicast && !(no2 && no2[0] == '1') && (tiles_two > sms / 2 || (no2 && no2[0] == '2'));
  const bool quad = two && qd && qd[0] == '1' && M > 2 * BM && ep->row_mask == nullptr;
  const int mode = quad ? 3 : two ? 2 : multicast ? 1 : 0;
  CUtensorMap tmA, tmB;
  uint64_t dA[2] = {uint64_t(K), uint64_t(M)}, sA[1] = {uint64_t(lda) * 2};
  uint32_t bA[2] = {BK, BM};
  int rc = get_tmap_bf16_sw128(&tmA, A, 2, dA, sA, bA);
  if (rc) return rc;
  uint64_t dB[2] = {uint64_t(K), uint64_t(N)}, sB[1] = {uint64_t(ldw) * 2};
  uint32_t bB[2] = {BK, uint32_t(quad ? BN / 4 : multicast ? BN / 2 : BN)};
  rc = get_tmap_bf16_sw128(&tmB, W, 2, dB, sB, bB);
  if (rc) return rc;

  EpiArgs e;
  e.C = ep->C;
  e.ldc = ep->ldc;
  e.out_fp32 = ep->out_fp32;
  e.bias = static_cast<const __nv_bfloat16*>(ep->bias);
  e.resid = static_cast<const __nv_bfloat16*>(ep->resid);
  e.ldr = ep->ldr;
  e.gate = static_cast<const __nv_bfloat16*>(ep->gate);
  e.ldg = ep->ldg;
  e.row_group = ep->row_group;
  e.rowvec = static_cast<const __nv_bfloat16*>(ep->rowvec);
  e.ldv = ep->ldv;
  e.rms_wq = static_cast<const __nv_bfloat16*>(ep->rms_wq);
  e.rms_wk = static_cast<const __nv_bfloat16*>(ep->rms_wk);
  e.rms_q_cols = ep->rms_q_cols;
  e.rms_k_cols = ep->rms_k_cols;
  e.rms_eps = ep->rms_eps;
  e.q_scale = ep->q_scale;
  e.act = ep->act;
  e.row_mask = ep->row_mask;
  e.row_mask_shift = ep->row_mask_shift;
  e.stats_out = nullptr;  // produced by the convolution kernel only
  e.ln_stats = reinterpret_cast<const float2*>(ep->ln_stats);
  e.ln_colsum = ep->ln_colsum;
  e.ln_rowpart = reinterpret_cast<const float2*>(ep->ln_rowpart);
  e.ln_nparts = ep->ln_nparts;
  e.part_ld = M;
  e.w_static = ep->w_static;
  e.ln_eps = ep->ln_eps;
  e.rowpart_out = reinterpret_cast<float2*>(ep->rowpart_out);
  if (ep->ln_colsum != nullptr && (ep->ln_stats == nullptr) == (ep->ln_rowpart == nullptr)) return B200_ERR_INVALID;
  if (ep->ln_colsum == nullptr && (ep->ln_stats != nullptr || ep->ln_rowpart != nullptr)) return B200_ERR_INVALID;
  if (ep->ln_rowpart != nullptr && ep->ln_nparts * 64 != K) return B200_ERR_INVALID;
  if (ep->rowpart_out != nullptr && (epi_mode == EPI_GEGLU || ep->out_fp32)) return B200_ERR_INVALID;
  if (ep->row_mask && ep->row_mask_shift < 8) return B200_ERR_INVALID;  // a CTA pair covers 256 rows
  GemmShape s{M, N, K};
  // output / residual tensor maps of the staged epilogue (bf16 outputs only)
  CUtensorMap tmC = tmA, tmR = tmA;
  if (!ep->out_fp32) {
    if (reinterpret_cast<uintptr_t>(ep->C) & 15) return B200_ERR_INVALID;
    const bool geglu = epi_mode == EPI_GEGLU;
    uint64_t dC[2] = {uint64_t(geglu ? N / 2 : N), uint64_t(M)}, sC[1] = {uint64_t(ep->ldc) * 2};
    uint32_t bC[2] = {geglu ? 32u : 64u, BM};
    rc = get_tmap_bf16(&tmC, ep->C, 2, dC, sC, bC, geglu ? 0 : 128);
    if (rc) return rc;
    if (epi_mode == EPI_GATE_RESID && ep->resid) {
      if ((reinterpret_cast<uintptr_t>(ep->resid) & 15) || (ep->ldr & 7)) return B200_ERR_INVALID;
      uint64_t dR[2] = {uint64_t(N), uint64_t(M)}, sR[1] = {uint64_t(ep->ldr) * 2};
      uint32_t bR[2] = {64u, BM};
      rc = get_tmap_bf16(&tmR, ep->resid, 2, dR, sR, bR, 128);
      if (rc) return rc;
    }
  }
  if (BN == 256) return dispatch_epi<256>(epi_mode, tmA, tmB, tmC, tmR, s, e, sms, stream, mode);
  if (BN == 192) return dispatch_epi<192>(epi_mode, tmA, tmB, tmC, tmR, s, e, sms, stream, mode);
  return dispatch_epi<128>(epi_mode, tmA, tmB, tmC, tmR, s, e, sms, stream, mode);
}

#ifdef GEMM_TRACE
extern "C" int b200_debug_gemm_trace(unsigned long long* out /* [64][10] */, unsigned int* n_launches) {
  cudaError_t e1 = cudaMemcpyFromSymbol(out, b200::g_trace, sizeof(unsigned long long) * 640);
  cudaError_t e2 = cudaMemcpyFromSymbol(n_launches, b200::g_trace_launch, sizeof(unsigned int));
  return (e1 == cudaSuccess && e2 == cudaSuccess) ? 0 : 1;
}
#endif
