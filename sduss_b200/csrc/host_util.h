// Host-side helpers shared by the launchers: status codes, the driver entry point for
// cuTensorMapEncodeTiled (resolved through cudart so the library has no link-time
// dependency on libcuda and still loads on a CPU-only box), and a tensor-map cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sduss_b200.h"  // B200_OK / B200_ERR_* status codes

namespace b200 {

// Encodes a bf16 tensor map with 128B swizzle (inner box = 64 elements = 128 bytes).
// dims/strides/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
// Returns B200_OK or an error code. Results are cached by value of all arguments.
int get_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box);
// Same with an explicit swizzle: 0 = none, 64 = SWIZZLE_64B, 128 = SWIZZLE_128B.
int get_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle);

inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? B200_OK : static_cast<int>(e);
}

int device_sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel AND device: the attribute belongs to
// the current device's context, so a process that switches devices must set it again there (`done` is
// the caller's static bit mask, one bit per device ordinal; ordinals >= 64 set it on every call).
template <typename K>
inline int ensure_dynamic_smem(K kern, int bytes, unsigned long long* done) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B200_ERR_DRIVER;
  if (dev < 64 && ((*done >> dev) & 1ull)) return B200_OK;
  const cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err != cudaSuccess) return static_cast<int>(err);
  if (dev < 64) *done |= 1ull << dev;
  return B200_OK;
}

// Tile-width choice for the 128-row tcgen05 GEMM / implicit-GEMM kernels. A tile of 128 x BN does
// 2 BN tensor-pipe cycles per 64-deep K block and pulls 16 KB of A plus BN x 128 B of W through
// L2 -> SM (W halved when CTA pairs multicast it). Small-M, deep-K problems (SDXL's level-2 token
// GEMMs: M = 2560) are bound by that operand traffic, not by SM occupancy: 100 tiles of 128 x 256
// on 148 SMs beat 200 tiles of 128 x 128 (measured 58 us -> see profiles/). Returns 256 or 128.
// L2 -> SM throughput used by the model: ~3600 B/clk chip-wide (fitted on M=2560 N=1280 K=5120).
inline int choose_tile_n(long m_tiles, int N, int num_kb, int sms, bool multicast) {
  if (N < 256) return 128;
  double best = 0;
  int pick = 128;
  for (int bn : {256, 128}) {
    const long tiles = m_tiles * ((N + bn - 1) / bn);
    const long rounds = (tiles + sms - 1) / sms;
    const double t_math = double(rounds) * num_kb * (2.0 * bn);
    const double w_bytes = bn * 128.0 * ((multicast && m_tiles > 1) ? 0.5 : 1.0);
    const double t_l2 = double(tiles) * num_kb * (16384.0 + w_bytes) / 3600.0;
    const double t = (t_math > t_l2 ? t_math : t_l2) + 2048.0 * rounds;  // + per-tile epilogue / fill
    if (best == 0 || t < best) { best = t; pick = bn; }
  }
  return pick;
}

// SDUSS_B200_NO_PDL=1 disables programmatic dependent launch (plain stream order).
bool pdl_enabled();

// Launches `kern` with the programmatic-stream-serialization attribute (PDL). The kernel must
// execute pdl_wait() before its first dependent global-memory access.
template <typename... KArgs, typename... Args>
inline int launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  return err == cudaSuccess ? launch_status() : static_cast<int>(err);
}

// Same with a thread-block cluster of `cluster_x` CTAs along x (grid.x must be a multiple).
template <typename... KArgs, typename... Args>
inline int launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  int n = 1;
  if (cluster_x > 1) {
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    n = 2;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaError_t err = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  return err == cudaSuccess ? launch_status() : static_cast<int>(err);
}

}  // namespace b200
