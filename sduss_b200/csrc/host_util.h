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

}  // namespace b200
