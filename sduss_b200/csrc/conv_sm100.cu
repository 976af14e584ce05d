// 3x3 convolution (padding 1, stride 1 or 2) over packed NHWC activations of a mixed-resolution
// batch as an implicit GEMM on tcgen05 tensor cores. Replaces F.conv2d on haloed 32x32 patches
// plus the halo-exchange kernels of the reference (sduss/model_executor/modules/resnet.py:
// 102-133,262-278,350-378 and kernels/norm_silu_concat.cu:248 "get_adjacency").
//
// Layout: activations are [sum_i H_i*W_i, C] bf16 (pixel rows of all latents back to back).
// One 128-row M tile = a 16x8 (rows x cols) pixel block of ONE latent. For filter tap (dy,dx)
// and channel block c0 the A tile is ONE TMA box load {64 ch, 8 px, 16 rows} from that
// latent's tensor map at coordinates (c0, x0+dx-1, y0+dy-1); TMA zero-fills out-of-bounds
// pixels, which IS the conv padding -- no halo exchange, no padded copies, exact corners.
// K loop = 9 taps x Cin/64 channel blocks, accumulating in TMEM. The stride-2 variant views the
// input as [H/2, 2, W/2, 2, C] through a 5-D tensor map so each tap is again a dense box.
// Same warp roles / pipeline / epilogues as gemm_sm100.cu.
#include "../../include/sduss_b200.h"
#include "epilogue.cuh"
#include <cstdlib>

#include "host_util.h"

namespace b200 {

constexpr int CV_BM = 128, CV_BK = 64, CV_TW = 8, CV_TH = 16, CV_THREADS = 256;

// MT = 2: the CTA runs TWO M tiles (two entries of the tile list, any two pixel blocks) against ONE
// weight tile per k-step. A 128 x 128 tile loads 32 KB of operands per 2.1 MFLOP, more than L2
// delivers per SM (the 128-channel convolutions of the VAE decoder ran at 0.8 PFLOP/s); with two
// M tiles the ratio is that of the 128 x 256 configuration (48 KB per 4.2 MFLOP).
// TWO: a CTA pair (cluster of 2) runs one tcgen05.mma.cta_group::2 of M = 256 per K step: each CTA gathers
// ITS pixel block (one M tile) and keeps half of the weight tile (gemm_sm100.cu, same scheme): 16 KB +
// BN x 64 B per stage, 6-8 stages, and the W bytes through L2 -> SM halve.
template <int BN, int EPI, int MT = 1, bool TWO = false>
struct ConvCfg {
  static_assert(MT == 1 || BN == 128, "two M tiles: 2 x 2 x 128 TMEM columns");
  static_assert(!TWO || MT == 1, "the pair already covers two M tiles");
  // residual chunks are TMA-loaded into the output staging tiles and updated in place (see
  // gemm_sm100.cu): the residual epilogue costs no pipeline stage
  static constexpr int kStages = TWO ? (BN == 256 ? 6 : 8) : (BN == 256 || MT == 2) ? 4 : 5;
  static constexpr int kATile = CV_BM * CV_BK * 2;
  static constexpr int kABytes = MT * kATile;
  static constexpr int kBBytes = (TWO ? BN / 2 : BN) * CV_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes =
      kStages * kStageBytes + 2 * EPI_STAGE_BYTES + 1024 + 256;
  static constexpr int kTmemCols = 2 * MT * BN;
};

struct ConvArgs {
  const CUtensorMap* in_maps;  // [n_latents] device array, 64-byte aligned entries
  const CUtensorMap* out_maps; // [n_latents] output maps {Cout, Wout, Hout}, box {64, 8, 16}
  const CUtensorMap* res_maps; // [n_latents] residual maps (same geometry) or null
  const int4* tiles;           // [n_mtiles] {latent, y0, x0, 0} in OUTPUT pixel coordinates
  const int4* lat;             // [n_latents] {output row offset, Hout, Wout, 0}
  int n_mtiles, Cin, Cout, stride;
  // patch cache (SDXL variant): pixel blocks whose rows all lie in clean patches are
  // skipped (no loads, no MMA, no store: the output keeps what it had). mask: int32 per 2^shift packed
  // rows of the level the block's decision was taken on; scale = log2(rows there / output rows):
  // 0 same level, +2 a stride-2 downsampler (output one level down), -2 the upsampler's convolution.
  const int* mask; int mask_shift, mask_scale;
};

// all rows of M tile m (16 output pixel rows of one latent, whatever the columns) lie in clean patches
__device__ __forceinline__ bool conv_mtile_clean(const ConvArgs& a, int m) {
  const int4 tl = a.tiles[m];
  const int4 ld = a.lat[tl.x];
  long lo = ld.x + long(tl.y) * ld.z, hi = ld.x + long(min(tl.y + 16, ld.y)) * ld.z;  // output rows [lo, hi)
  if (a.mask_scale > 0) { lo <<= a.mask_scale; hi <<= a.mask_scale; }
  else if (a.mask_scale < 0) { lo >>= -a.mask_scale; hi = (hi + (1L << -a.mask_scale) - 1) >> -a.mask_scale; }
  for (long b = lo >> a.mask_shift; b <= ((hi - 1) >> a.mask_shift); ++b)
    if (a.mask[b] != 0) return false;
  return true;
}
// a work item (MTI consecutive M tiles starting at m_first) is skipped when all its tiles are clean:
// the same answer in every warp role and in both CTAs of a pair
template <int MTI>
__device__ __forceinline__ bool conv_item_skipped(const ConvArgs& a, int m_first) {
  if (a.mask == nullptr) return false;
#pragma unroll
  for (int i = 0; i < MTI; ++i)
    if (m_first + i < a.n_mtiles && !conv_mtile_clean(a, m_first + i)) return false;
  return true;
}

template <int BN, int EPI, int MT, bool TWO>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmW, ConvArgs a, EpiArgs e, int M_total) {
  using Cfg = ConvCfg<BN, EPI, MT, TWO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* stageC = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stageC + 2 * EPI_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* tfull = bars + 2 * Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* rfull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (a.Cout + BN - 1) / BN;
  // work item = MT consecutive M tiles x one N tile; TWO: the pair's item = 2 M tiles (one per CTA)
  constexpr int MTI = TWO ? 2 : MT;
  const int m_items = (a.n_mtiles + MTI - 1) / MTI;
  const int num_tiles = m_items * tiles_n;
  const int rank = TWO ? int(cluster_ctarank()) : 0;
  const int first_tile = TWO ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int tile_step = TWO ? int(gridDim.x >> 1) : int(gridDim.x);
  const int cblocks = a.Cin / CV_BK;
  const int num_kb = 9 * cblocks;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmW);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], TWO ? 8 : 4);  // TWO: the epilogue warps of both CTAs release the leader's
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
  if constexpr (TWO) cluster_sync_all();  // the peer's barriers exist before anything signals them
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  // bytes one phase of full[stage] waits for (TWO: on the leader, both CTAs' loads; the second CTA of
  // the last pair may have no pixel block)
  auto expect_bytes = [&](int t) -> uint32_t {
    if constexpr (TWO) {
      const bool peer = (t / tiles_n) * 2 + 1 < a.n_mtiles;
      return uint32_t((peer ? 2 : 1) * Cfg::kATile + 2 * Cfg::kBBytes);
    } else {
      return uint32_t(min(MT, a.n_mtiles - (t / tiles_n) * MT) * Cfg::kATile + Cfg::kBBytes);
    }
  };
  auto load_w = [&](int stage, int kb, int n0) {
    if constexpr (TWO) tma_load_2d_2sm(smemB + stage * Cfg::kBBytes, &tmW, &full[stage], kb * CV_BK, n0 + rank * (BN / 2));
    else tma_load_2d(smemB + stage * Cfg::kBBytes, &tmW, &full[stage], kb * CV_BK, n0);
  };
  // the weight tiles of the first ring stages are requested before the wait on the previous kernel
  // (weights are never written on the stream; see gemm_sm100.cu): their HBM latency runs under its tail
  int npre = 0;
  if (warp == 0 && first_tile < num_tiles && a.mask == nullptr) {  // (the mask is written by the previous kernel)
    npre = num_kb < Cfg::kStages ? num_kb : Cfg::kStages;
    if (lane == 0) {
      const int t = first_tile;
      const int n0 = (t % tiles_n) * BN;
      for (int kb = 0; kb < npre; ++kb) {
        if (rank == 0) mbar_expect_tx(&full[kb], expect_bytes(t));
        load_w(kb, kb, n0);
      }
    }
    __syncwarp();
  }
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const int m0 = TWO ? (t / tiles_n) * 2 + rank : (t / tiles_n) * MT;
      const int nmt = TWO ? (m0 < a.n_mtiles ? 1 : 0) : min(MT, a.n_mtiles - m0);
      const int n0 = (t % tiles_n) * BN;
      if (conv_item_skipped<MTI>(a, (t / tiles_n) * MTI)) continue;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * CV_BK;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
          const bool pre = t == first_tile && kb < npre;  // W of this stage is already on its way
          if (!pre && rank == 0) mbar_expect_tx(&full[stage], expect_bytes(t));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < nmt) {
              const int4 tl = a.tiles[m0 + mt];
              const CUtensorMap* im = a.in_maps + tl.x;
              void* dstA = smemA + stage * Cfg::kABytes + mt * Cfg::kATile;
              if (a.stride == 1) {
                if constexpr (TWO) tma_load_3d_2sm(dstA, im, &full[stage], c0, tl.z + dx, tl.y + dy);
                else tma_load_3d(dstA, im, &full[stage], c0, tl.z + dx, tl.y + dy);
              } else {
                // input pixel (2y+dy, 2x+dx) = (parity, half index): -1 -> (1, i-1); 0 -> (0, i); 1 -> (1, i)
                const int py = dy == 0 ? 0 : 1, px = dx == 0 ? 0 : 1;
                if constexpr (TWO)
                  tma_load_5d_2sm(dstA, im, &full[stage], c0, px, tl.z + (dx < 0 ? -1 : 0), py, tl.y + (dy < 0 ? -1 : 0));
                else
                  tma_load_5d(dstA, im, &full[stage], c0, px, tl.z + (dx < 0 ? -1 : 0), py, tl.y + (dy < 0 ? -1 : 0));
              }
            }
          }
          if (!pre) load_w(stage, kb, n0);
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (TWO: the leader's only)
    constexpr uint32_t idesc = make_idesc_bf16(TWO ? 2 * CV_BM : CV_BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      if (conv_item_skipped<MTI>(a, (t / tiles_n) * MTI)) continue;
      const int acc = it & 1;
      mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (MT * BN);
      const int nmt = TWO ? 1 : min(MT, a.n_mtiles - (t / tiles_n) * MT);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t db = make_sdesc_sw128(smem_u32(smemB + stage * Cfg::kBBytes));
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < nmt) {
              const uint64_t da = make_sdesc_sw128(smem_u32(smemA + stage * Cfg::kABytes + mt * Cfg::kATile));
#pragma unroll
              for (int k = 0; k < CV_BK / 16; ++k) {
                if constexpr (TWO) umma_ss2(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                else umma_ss(d_tmem + mt * BN, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              }
            }
          }
          if constexpr (TWO) {
            umma_commit2_mc(&empty[stage], uint16_t(3));
            if (kb == num_kb - 1) umma_commit2_mc(&tfull[acc], uint16_t(3));
          } else {
            umma_commit(&empty[stage]);
            if (kb == num_kb - 1) umma_commit(&tfull[acc]);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      ++it;  // accumulator stages count the items actually computed
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (TMA-staged)
    const int q = warp & 3;
    const int r = q * 32 + lane;              // row in tile: pixel (r / 8, r % 8) of the block
    const bool leader = threadIdx.x == 128;
    const bool has_resid = (EPI == EPI_GATE_RESID) && a.res_maps != nullptr;
    constexpr int NCH = BN / 64;
    uint32_t ruse0 = 0, ruse1 = 0;
    // Staging buffers alternate over ALL chunks this CTA stores, not per tile: a tile with an odd
    // number of chunks (N tail, N < BN) would otherwise be followed by a chunk that reuses the
    // buffer whose TMA store is still in flight (the store is only waited for after the next
    // chunk has been staged). Seen as corrupted tail columns on short-K GEMMs (conv_in: K = 64).
    uint32_t chunk_ctr = 0;
    int it = 0;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      if (conv_item_skipped<MTI>(a, (t / tiles_n) * MTI)) continue;
      const int acc = it & 1;
      const int m0 = TWO ? (t / tiles_n) * 2 + rank : (t / tiles_n) * MT;
      const int nmt = TWO ? (m0 < a.n_mtiles ? 1 : 0) : min(MT, a.n_mtiles - m0);
      const int n0 = (t % tiles_n) * BN;
      if (TWO && nmt == 0) {  // the last pair's second CTA without a pixel block: hand the accumulator back
        mbar_wait(&tfull[acc], (it >> 1) & 1);
      }
#pragma unroll 1
      for (int mt = 0; mt < nmt; ++mt) {
        const int4 tl = a.tiles[m0 + mt];
        const int4 ld = a.lat[tl.x];
        if (has_resid && leader) {  // residual chunks 0 and 1 fly while the main loop finishes
          tma_store_wait_read<0>();   // both staging tiles have left for HBM (previous tile)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c < NCH && n0 + c * 64 < a.Cout) {
              const int rb = (chunk_ctr + c) & 1;
              mbar_expect_tx(&rfull[rb], EPI_STAGE_BYTES);
              tma_load_3d(stageC + rb * EPI_STAGE_BYTES, a.res_maps + tl.x, &rfull[rb], n0 + c * 64, tl.z, tl.y);
            }
          }
        }
        __syncwarp();
        if (mt == 0) {
          mbar_wait(&tfull[acc], (it >> 1) & 1);
          tc_fence_after();
        }
        const int y = tl.y + r / CV_TW, x = tl.z + r % CV_TW;
        const bool row_ok = y < ld.y && x < ld.z;
        const int row = row_ok ? ld.x + y * ld.z + x : M_total;
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * (MT * BN) + mt * BN;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int nc = n0 + c * 64;
          if (nc >= a.Cout) break;
          const int b = (chunk_ctr + c) & 1;
          if (has_resid && c >= 1 && c + 1 < NCH && nc + 64 < a.Cout) {
            // residual of chunk c + 1 -> the tile chunk c - 1 was stored from, once that store has
            // read it; it lands while this chunk is processed
            if (leader) {
              tma_store_wait_read<0>();
              mbar_expect_tx(&rfull[b ^ 1], EPI_STAGE_BYTES);
              tma_load_3d(stageC + (b ^ 1) * EPI_STAGE_BYTES, a.res_maps + tl.x, &rfull[b ^ 1], nc + 64, tl.z, tl.y);
            }
            __syncwarp();
          }
          float v[64];
          tmem_ld32(t_row + c * 64, reinterpret_cast<uint32_t*>(v));
          tmem_ld32(t_row + c * 64 + 32, reinterpret_cast<uint32_t*>(v) + 32);
          tmem_wait_ld();
          if (has_resid) {
            mbar_wait(&rfull[b], (b ? ruse1 : ruse0) & 1);
            if (b) ++ruse1; else ++ruse0;
          }
          epilogue_math64<EPI>(e, v, row, row_ok, nc, a.Cout, has_resid ? stageC + b * EPI_STAGE_BYTES : nullptr, r);
          epilogue_stage64<EPI>(stageC + b * EPI_STAGE_BYTES, v, r);
          fence_proxy_async();
          if (leader) tma_store_wait_read<0>();
          named_barrier<1, 128>();
          if (leader) {
            tma_store_3d(a.out_maps + tl.x, stageC + b * EPI_STAGE_BYTES, nc, tl.z, tl.y);
            tma_store_commit();
          }
          if (e.stats_out != nullptr) {
            // GroupNorm statistics of the tensor being written, produced here instead of by a second
            // pass over it: thread (half, col) sums column col of the staged bf16 tile over its 64
            // rows in a fixed order (rows outside the latent excluded). Partials go to fixed slots
            // [(tile, half)][channel]; the finalize kernel reduces them in a fixed order, so the
            // result depends on the latent alone (batch invariance, determinism).
            const int tid = int(threadIdx.x) - 128;
            const int col = tid & 63, half = tid >> 6;
            if (nc + col < a.Cout) {
              const uint8_t* st = stageC + b * EPI_STAGE_BYTES;
              float sm = 0.f, sq = 0.f;
#pragma unroll 8
              for (int rr = half * 64; rr < half * 64 + 64; ++rr) {
                if (tl.y + rr / CV_TW < ld.y && tl.z + rr % CV_TW < ld.z) {
                  const float f = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(
                      st + sw128_off(rr, col >> 3) + (col & 7) * 2));
                  sm += f;
                  sq = fmaf(f, f, sq);
                }
              }
              e.stats_out[(size_t(m0 + mt) * 2 + half) * a.Cout + nc + col] = make_float2(sm, sq);
            }
            // the leader refills this staging tile with a residual chunk at the top of the next
            // iteration: every column reader must be done first
            if (has_resid) named_barrier<1, 128>();
          }
        }
        {
          const int left = (a.Cout - n0 + 63) / 64;
          chunk_ctr += uint32_t(left < NCH ? left : NCH);
        }
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
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // the peer may still signal my barriers / read my tiles until here
  if (warp == 2) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI, int MT, bool TWO = false>
static int launch_conv(const CUtensorMap& tmW, const ConvArgs& a, const EpiArgs& e, int M_total,
                       int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, EPI, MT, TWO>;
  auto kern = conv3x3_kernel<BN, EPI, MT, TWO>;
  static unsigned long long configured = 0;
  if (int rc = ensure_dynamic_smem(kern, Cfg::kSmemBytes, &configured)) return rc;
  if constexpr (TWO) {
    const int tiles = ((a.n_mtiles + 1) / 2) * ((a.Cout + BN - 1) / BN);
    const int grid = (tiles < num_sms / 2 ? tiles : num_sms / 2) * 2;
    return launch_pdl_cluster(kern, dim3(grid), dim3(CV_THREADS), Cfg::kSmemBytes, stream, 2, tmW, a, e, M_total);
  }
  const int tiles = ((a.n_mtiles + MT - 1) / MT) * ((a.Cout + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  return launch_pdl(kern, dim3(grid), dim3(CV_THREADS), Cfg::kSmemBytes, stream, tmW, a, e, M_total);
}

template <int BN, int MT, bool TWO = false>
static int dispatch_conv(int epi, const CUtensorMap& tmW, const ConvArgs& a, const EpiArgs& e,
                         int M_total, int sms, cudaStream_t st) {
  switch (epi) {
    case EPI_BIAS: return launch_conv<BN, EPI_BIAS, MT, TWO>(tmW, a, e, M_total, sms, st);
    case EPI_GATE_RESID: return launch_conv<BN, EPI_GATE_RESID, MT, TWO>(tmW, a, e, M_total, sms, st);
    case EPI_ROWVEC: return launch_conv<BN, EPI_ROWVEC, MT, TWO>(tmW, a, e, M_total, sms, st);
    default: return B200_ERR_UNSUPPORTED;
  }
}

}  // namespace b200

using namespace b200;

// Encodes one input tensor map per latent into `maps_host` (host memory, n_latents entries of
// 128 bytes); the caller uploads them to a 64-byte-aligned device buffer once per batch
// composition. in_desc: int32 [n][4] = {input row offset, Hin, Win, 0}.
extern "C" long long b200_conv3x3_maps_bytes(int n_latents) { return (long long)n_latents * 128; }

extern "C" int b200_conv3x3_encode_maps(const void* x, int ldx, int Cin, const int32_t* in_desc_host,
                                        int n_latents, int stride, void* maps_host) {
  if (!x || !in_desc_host || !maps_host || n_latents <= 0 || (Cin & 7) || (ldx & 7) ||
      (stride != 1 && stride != 2))
    return B200_ERR_INVALID;
  CUtensorMap* out = static_cast<CUtensorMap*>(maps_host);
  for (int i = 0; i < n_latents; ++i) {
    const int row0 = in_desc_host[4 * i], H = in_desc_host[4 * i + 1], W = in_desc_host[4 * i + 2];
    const char* base = static_cast<const char*>(x) + size_t(row0) * ldx * 2;
    int rc;
    if (stride == 1) {
      uint64_t d[3] = {uint64_t(Cin), uint64_t(W), uint64_t(H)};
      uint64_t s[2] = {uint64_t(ldx) * 2, uint64_t(ldx) * 2 * W};
      uint32_t b[3] = {CV_BK, CV_TW, CV_TH};
      rc = get_tmap_bf16_sw128(&out[i], base, 3, d, s, b);
    } else {
      if ((H & 1) || (W & 1)) return B200_ERR_INVALID;
      uint64_t d[5] = {uint64_t(Cin), 2, uint64_t(W / 2), 2, uint64_t(H / 2)};
      uint64_t s[4] = {uint64_t(ldx) * 2, uint64_t(ldx) * 4, uint64_t(ldx) * 2 * W,
                       uint64_t(ldx) * 4 * W};
      uint32_t b[5] = {CV_BK, 1, CV_TW, 1, CV_TH};
      rc = get_tmap_bf16_sw128(&out[i], base, 5, d, s, b);
    }
    if (rc) return rc;
  }
  return B200_OK;
}

extern "C" int b200_conv3x3_bf16(const void* in_maps_dev, const void* out_maps_dev,
                                 const void* resid_maps_dev, const int32_t* tiles_dev, int n_mtiles,
                                 const int32_t* out_lat_dev, int Cin, int Cout, int stride,
                                 const void* Wt, int M_total, int epi_mode,
                                 const B200EpilogueDesc* ep, void* stream_) {
  if (!in_maps_dev || !out_maps_dev || !tiles_dev || !out_lat_dev || !Wt || !ep || n_mtiles <= 0)
    return B200_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(out_maps_dev) & 63) || (reinterpret_cast<uintptr_t>(resid_maps_dev) & 63))
    return B200_ERR_INVALID;
  if ((Cin % CV_BK) || (Cout & 7) || (stride != 1 && stride != 2))
    return B200_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(in_maps_dev) & 63) != 0) return B200_ERR_INVALID;
  if (epi_mode == EPI_ROWVEC && (!ep->rowvec || !ep->row_group)) return B200_ERR_INVALID;
  const int sms = device_sm_count();
  if (sms <= 0) return B200_ERR_DRIVER;
  const int BN = choose_tile_n(n_mtiles, Cout, 9 * (Cin / CV_BK), sms, /*multicast=*/false);
  const bool bn256 = BN == 256;
  // CTA pairs as cta_group::2 (one pixel block per CTA, half of the weight tile each) once there is more
  // than one round of pair tiles; SDUSS_B200_CONV_NO_2CTA=1 (read per call, for A/B runs) keeps round 1's
  // single-CTA tiles
  const char* no2 = getenv("SDUSS_B200_CONV_NO_2CTA");
  const long pair_tiles = long((n_mtiles + 1) / 2) * ((Cout + BN - 1) / BN);
  // (256-wide tiles only: with 128 output channels the two-M-tiles-per-CTA variant below moves the same
  //  bytes per FLOP and measured 911 vs 729 TFLOP/s on the VAE's 128-channel level)
  const bool two = !(no2 && no2[0] == '1') && bn256 && n_mtiles >= 2 && pair_tiles > sms / 2;
  CUtensorMap tmW;
  uint64_t d[2] = {uint64_t(9) * Cin, uint64_t(Cout)}, s[1] = {uint64_t(9) * Cin * 2};
  uint32_t b[2] = {CV_BK, uint32_t(two ? BN / 2 : BN)};
  int rc = get_tmap_bf16_sw128(&tmW, Wt, 2, d, s, b);
  if (rc) return rc;
  ConvArgs a{static_cast<const CUtensorMap*>(in_maps_dev),
             static_cast<const CUtensorMap*>(out_maps_dev),
             static_cast<const CUtensorMap*>(resid_maps_dev), reinterpret_cast<const int4*>(tiles_dev),
             reinterpret_cast<const int4*>(out_lat_dev), n_mtiles, Cin, Cout, stride,
             ep->row_mask, ep->row_mask_shift, ep->row_mask_scale};
  if (ep->row_mask && (ep->row_mask_shift < 6 || ep->row_mask_scale < -4 || ep->row_mask_scale > 4 || ep->stats_out))
    return B200_ERR_INVALID;
  EpiArgs e;
  e.C = ep->C; e.ldc = ep->ldc; e.out_fp32 = ep->out_fp32;
  e.bias = static_cast<const __nv_bfloat16*>(ep->bias);
  e.resid = static_cast<const __nv_bfloat16*>(ep->resid); e.ldr = ep->ldr;
  e.gate = static_cast<const __nv_bfloat16*>(ep->gate); e.ldg = ep->ldg;
  e.row_group = ep->row_group;
  e.rowvec = static_cast<const __nv_bfloat16*>(ep->rowvec); e.ldv = ep->ldv;
  e.rms_wq = nullptr; e.rms_wk = nullptr; e.rms_q_cols = 0; e.rms_k_cols = 0;
  e.rms_eps = 0.f; e.q_scale = 1.f; e.act = 0; e.row_mask = nullptr; e.row_mask_shift = 0;
  e.stats_out = reinterpret_cast<float2*>(ep->stats_out);
  e.ln_stats = nullptr; e.ln_colsum = nullptr; e.ln_rowpart = nullptr; e.ln_nparts = 0; e.part_ld = 0; e.ln_eps = 0.f;
  e.rowpart_out = nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  if (two) return dispatch_conv<256, 1, true>(epi_mode, tmW, a, e, M_total, sms, st);
  if (bn256) return dispatch_conv<256, 1>(epi_mode, tmW, a, e, M_total, sms, st);
  // 128-wide tiles with several rounds of tile pairs: two M tiles per CTA share each weight tile
  // (halves the operand bytes per FLOP, see ConvCfg). Preferring this over 256-wide tiles where it
  // saves a round (Cout = 320: 3 x 128 columns in 4 rounds instead of 2 x 256 in 5) was measured
  // and is slower (SDXL level 0: 834 -> 721 TFLOP/s): not done.
  static const bool no_mt2 = []() { const char* v = getenv("SDUSS_B200_CONV_NO_MT2"); return v && v[0] == '1'; }();
  const long items2 = long((n_mtiles + 1) / 2) * ((Cout + 127) / 128);
  if (!no_mt2 && items2 >= 4L * sms) return dispatch_conv<128, 2>(epi_mode, tmW, a, e, M_total, sms, st);
  return dispatch_conv<128, 1>(epi_mode, tmW, a, e, M_total, sms, st);
}
