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
