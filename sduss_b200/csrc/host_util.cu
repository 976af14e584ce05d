#include "host_util.h"

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

namespace b200 {

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SDUSS_B200_NO_PDL");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  return fn;
}

struct Key {
  uint64_t v[1 + 1 + 5 + 4 + 5 + 1];
  bool operator==(const Key& o) const { return std::memcmp(v, o.v, sizeof(v)) == 0; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};

int get_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
  return get_tmap_bf16(out, base, rank, dims, strides_bytes, box, 128);
}

int get_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  if (rank < 2 || rank > 5 || base == nullptr) return B200_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return B200_ERR_INVALID;
  Key key;
  std::memset(&key, 0, sizeof(key));
  key.v[0] = reinterpret_cast<uint64_t>(base);
  key.v[1] = static_cast<uint64_t>(rank);
  for (int i = 0; i < rank; ++i) key.v[2 + i] = dims[i];
  for (int i = 0; i < rank - 1; ++i) key.v[7 + i] = strides_bytes[i];
  for (int i = 0; i < rank; ++i) key.v[11 + i] = box[i];
  key.v[16] = static_cast<uint64_t>(swizzle);

  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return B200_OK;
    }
  }
  EncodeTiledFn enc = resolve_encode();
  if (!enc) return B200_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) {
    if (strides_bytes[i] & 15) return B200_ERR_INVALID;
    gstr[i] = strides_bytes[i];
  }
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                   : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return B200_ERR_DRIVER;
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return B200_OK;
}

}  // namespace b200
