// Patch cache ("block skip") -- the decision, on the device.
//
// Reference: CacheManager.get_sd3_mask (sduss/model_executor/modules/cache_manager.py:161-191): per
// transformer block and per 256-token patch, the feature row [block index, timestep,
// MSE(block input now, block input at the previous step)] goes through a RandomForest; a patch the
// forest calls stable is not recomputed, unless it has already been skipped `refresh` times in a
// row. The reference computes the MSE with torch ops, copies the features to the host, runs the
// (cuML) forest there and indexes tensors with the numpy mask: one device->host sync per block,
// 24 per step. Here the whole decision is one kernel per block that stays inside the captured CUDA
// graph: it streams the block input once (computing the per-patch MSE against the kept copy and
// refreshing that copy in the same pass), evaluates the flattened forest, applies the refresh rule
// and leaves an int32 mask per patch that the GEMM / LayerNorm / attention kernels consult
// (row_mask / q_mask arguments) to skip the tiles of clean patches.
#include <cuda_bf16.h>

#include "../../include/sduss_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b200 {

constexpr int PC_SPLIT = 16;     // CTAs per patch (each owns rows/16 of it): 16 x patches CTAs stream the tensor
constexpr int PC_THREADS = 256;

struct ForestDev {
  const int* feature;      // [nodes] feature index, < 0 for a leaf
  const float* threshold;  // [nodes] go left when x[feature] <= threshold (sklearn)
  const int* left;         // [nodes]
  const int* right;        // [nodes]
  const float* value;      // [nodes] leaf: probability of class "recompute"
  const int* roots;        // [n_trees]
  int n_trees;
};

struct PatchMaskArgs {
  const __nv_bfloat16* x; int ldx;     // block input, packed rows
  __nv_bfloat16* prev; int ldp;        // block input at the previous step (refreshed here)
  int n_patches, rows_per_patch, D;
  const int* patch_latent;             // [n_patches] latent of the patch
  const float* latent_t;               // [L] timestep of the latent (the model's t32 buffer)
  const float* latent_valid;           // [L] != 0: the kept copies of this latent are this request's
  int* skipped;                        // [n_patches] consecutive skips (previous_mask), in / out
  int* mask;                           // [n_patches] out: 1 = recompute
  float* mse;                          // [n_patches] out (diagnostics / predictor fitting), may be null
  float* partial;                      // [n_patches][PC_SPLIT] workspace
  unsigned* arrive;                    // [n_patches] workspace, zero on entry, left zero
  float block_index; int refresh;
  // up-block variant (CacheManager.get_mask(is_upsample=True), cache_manager.py:106-135): the feature row
  // continues with the MSE of every skip tensor of the block; those come from earlier launches of this
  // kernel in MSE-only mode (f.n_trees == 0: no decision, mask / skipped untouched).
  const float* extra; int n_extra;     // [n_extra][n_patches], n_extra <= 3
  ForestDev f;
};

__global__ void __launch_bounds__(PC_THREADS)
patch_mask_kernel(const __grid_constant__ PatchMaskArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int p = blockIdx.x / PC_SPLIT, part = blockIdx.x % PC_SPLIT;
  const int rows = a.rows_per_patch / PC_SPLIT;
  const size_t row0 = size_t(p) * a.rows_per_patch + size_t(part) * rows;
  const int vec_per_row = a.D >> 3;
  float acc = 0.f;
  // fixed assignment of vectors to threads and a fixed reduction tree: the sum does not depend on
  // anything but the data (masks must be reproducible run to run)
  for (int i = threadIdx.x; i < rows * vec_per_row; i += PC_THREADS) {
    const int r = i / vec_per_row, c = i % vec_per_row;
    const uint4 xv = *reinterpret_cast<const uint4*>(a.x + (row0 + r) * a.ldx + c * 8);
    uint4* pp = reinterpret_cast<uint4*>(a.prev + (row0 + r) * a.ldp + c * 8);
    const uint4 pv = *pp;
    *pp = xv;
    const uint32_t* xs = &xv.x;
    const uint32_t* ps = &pv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0, x1, p0, p1;
      unpack_bf16x2(xs[k], x0, x1);
      unpack_bf16x2(ps[k], p0, p1);
      const float d0 = x0 - p0, d1 = x1 - p1;
      acc = fmaf(d0, d0, acc);
      acc = fmaf(d1, d1, acc);
    }
  }
  __shared__ float red[PC_THREADS / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < PC_THREADS / 32; ++w) s += red[w];
    a.partial[p * PC_SPLIT + part] = s;
    __threadfence();
    last = atomicInc(a.arrive + p, PC_SPLIT - 1) == PC_SPLIT - 1;  // wraps back to 0 for the next launch
  }
  __syncthreads();
  if (!last || threadIdx.x != 0) return;
  __threadfence();
  // ---- the patch's decision (one thread; a forest is a few hundred node visits)
  float s = 0.f;
  for (int k = 0; k < PC_SPLIT; ++k) s += reinterpret_cast<volatile float*>(a.partial)[p * PC_SPLIT + k];
  const int lat = a.patch_latent[p];
  const bool valid = a.latent_valid[lat] != 0.f;
  // cache_manager.py:19,166: a patch without a kept input gets MSE = float(sys.maxsize)
  const float mse = valid ? s / (float(a.rows_per_patch) * float(a.D)) : 9223372036854775807.f;
  if (a.mse != nullptr) a.mse[p] = mse;
  if (a.f.n_trees == 0) return;  // MSE-only launch (a skip tensor of an up block)
  const int prev_skips = valid ? a.skipped[p] : 0;
  float feat[6] = {a.block_index, a.latent_t[lat], mse, 0.f, 0.f, 0.f};
  for (int k = 0; k < a.n_extra; ++k) feat[3 + k] = __ldcg(a.extra + size_t(k) * a.n_patches + p);
  float votes = 0.f;
  for (int t = 0; t < a.f.n_trees; ++t) {
    int n = a.f.roots[t];
    while (a.f.feature[n] >= 0) n = feat[a.f.feature[n]] <= a.f.threshold[n] ? a.f.left[n] : a.f.right[n];
    votes += a.f.value[n];
  }
  // RandomForestClassifier.predict: class 1 iff its mean probability is the larger one. A patch
  // with nothing kept is always computed; `refresh` consecutive skips force a recompute
  // (cache_manager.py:183-186).
  const bool flagged = !valid || votes > 0.5f * float(a.f.n_trees) || prev_skips == a.refresh;
  a.mask[p] = flagged ? 1 : 0;
  a.skipped[p] = (flagged || prev_skips == a.refresh) ? 0 : prev_skips + 1;
}

}  // namespace b200

using namespace b200;

extern "C" long long b200_patch_mask_workspace_bytes(int n_patches) {
  return (long long)n_patches * (PC_SPLIT * sizeof(float) + sizeof(unsigned));
}

extern "C" int b200_patch_mask_bf16(const void* x, int ldx, void* prev, int ldp, int n_patches,
                                    int rows_per_patch, int D, const int32_t* patch_latent,
                                    const float* latent_t, const float* latent_valid, int32_t* skipped,
                                    int32_t* mask, float* mse, const B200Forest* forest, int block_index,
                                    int refresh, void* workspace, void* stream) {
  if (!forest) return B200_ERR_INVALID;
  return b200_patch_mask_ex(x, ldx, prev, ldp, n_patches, rows_per_patch, D, patch_latent, latent_t,
                            latent_valid, skipped, mask, mse, forest, block_index, refresh, nullptr, 0,
                            workspace, stream);
}

extern "C" int b200_patch_mask_ex(const void* x, int ldx, void* prev, int ldp, int n_patches,
                                  int rows_per_patch, int D, const int32_t* patch_latent,
                                  const float* latent_t, const float* latent_valid, int32_t* skipped,
                                  int32_t* mask, float* mse, const B200Forest* forest, int block_index,
                                  int refresh, const float* extra_mse, int n_extra, void* workspace,
                                  void* stream) {
  if (!x || !prev || !patch_latent || !latent_t || !latent_valid || !workspace || n_patches <= 0 ||
      rows_per_patch <= 0 || (rows_per_patch % PC_SPLIT) || D <= 0 || (D & 7) || (ldx & 7) || (ldp & 7) ||
      n_extra < 0 || n_extra > 3 || (n_extra > 0 && !extra_mse))
    return B200_ERR_INVALID;
  if (forest) {
    if (!skipped || !mask || forest->n_trees <= 0 || !forest->feature || !forest->threshold ||
        !forest->left || !forest->right || !forest->value || !forest->roots)
      return B200_ERR_INVALID;
  } else if (!mse) {
    return B200_ERR_INVALID;  // an MSE-only launch without an output
  }
  PatchMaskArgs a;
  a.x = static_cast<const __nv_bfloat16*>(x); a.ldx = ldx;
  a.prev = static_cast<__nv_bfloat16*>(prev); a.ldp = ldp;
  a.n_patches = n_patches; a.rows_per_patch = rows_per_patch; a.D = D;
  a.patch_latent = patch_latent; a.latent_t = latent_t; a.latent_valid = latent_valid;
  a.skipped = skipped; a.mask = mask; a.mse = mse;
  a.partial = static_cast<float*>(workspace);
  a.arrive = reinterpret_cast<unsigned*>(a.partial + size_t(n_patches) * PC_SPLIT);
  a.block_index = float(block_index); a.refresh = refresh;
  a.extra = extra_mse; a.n_extra = n_extra;
  if (forest)
    a.f = ForestDev{forest->feature, forest->threshold, forest->left, forest->right, forest->value,
                    forest->roots, forest->n_trees};
  else
    a.f = ForestDev{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  return launch_pdl(patch_mask_kernel, dim3(n_patches * PC_SPLIT), dim3(PC_THREADS), 0,
                    reinterpret_cast<cudaStream_t>(stream), a);
}
