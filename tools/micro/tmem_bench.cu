// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (32x32b.x32), 4 or 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sduss_b200/csrc/ptx.cuh"
using namespace b200;
template <int MODE>  // 0: ld 128 cols + wait; 1: st 64 cols + wait
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[128];
  for (int i = 0; i < 128; ++i) r[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_ld32(base, r); tmem_ld32(base + 32, r + 32); tmem_ld32(base + 64, r + 64); tmem_ld32(base + 96, r + 96);
      tmem_wait_ld();
      acc += r[0] ^ r[37] ^ r[77] ^ r[127];
    } else {
      r[5] += acc;
      tmem_st32(base, r); tmem_st32(base + 32, r + 32);
      tmem_wait_st();
      acc += it;
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = acc; }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}
template <int MODE> void run(int warps, const char* name, int bytes_per_warp_iter) {
  long long* d; cudaMalloc(&d, 148 * 16);
  int iters = 2000;
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  double clk = double(h[0]) / iters;
  printf("%s warps=%d: %.1f clk/iter, %.1f B/clk/SM (%s)\n", name, warps, clk, warps * bytes_per_warp_iter / clk, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>(4, "ld 128 cols", 128 * 32 * 4); run<0>(8, "ld 128 cols", 128 * 32 * 4);
  run<1>(4, "st  64 cols", 64 * 32 * 4);  run<1>(8, "st  64 cols", 64 * 32 * 4);
  return 0;
}
