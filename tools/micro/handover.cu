// Micro-benchmark: fixed cost (cycles) of the hand-over operations a softmax warp of the attention
// kernel executes every 64-key step, measured on data that is ALREADY there: mbarrier.try_wait on
// a completed phase (+ tcgen05.fence::after_thread_sync), tcgen05.ld 64 columns + wait::ld,
// tcgen05.st 32 columns + wait::st, mbarrier.arrive by one lane after __syncwarp. One warp, 4 warps
// and 12 warps per CTA (the attention kernel runs 12 softmax warps).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handover handover.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sduss_b200/csrc/ptx.cuh"
using namespace b200;

template <int MODE>
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bar[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[64];
  for (int i = 0; i < 64; ++i) r[i] = threadIdx.x + i;
  if (lane == 0) mbar_arrive(&bar[warp]);  // phase 0 of this warp's barrier is complete from here on
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // wait on a completed barrier + fence
      mbar_wait(&bar[warp], 0);
      tc_fence_after();
      acc += it;
    } else if (MODE == 1) {   // S pull: 64 columns
      tmem_ld32(base, r); tmem_ld32(base + 32, r + 32);
      tmem_wait_ld();
      tc_fence_before();
      acc += r[1] ^ r[40];
    } else if (MODE == 2) {   // P push: 32 columns
      r[3] += acc;
      tmem_st32(base, r);
      tmem_wait_st();
      tc_fence_before();
      acc += it;
    } else if (MODE == 3) {   // arrive by one lane after a warp sync
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[12 + (warp & 3)]);
      acc += it;
    } else if (MODE == 4) {   // try_wait alone
      mbar_wait(&bar[warp], 0);
      acc += it;
    } else if (MODE == 5) {   // test_wait (non-blocking probe) alone
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&bar[warp])), "r"(0u) : "memory");
      acc += ok;
    } else {                  // tcgen05.fence::after_thread_sync alone
      tc_fence_after();
      acc += it;
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) { out[(blockIdx.x * 16 + warp) * 2] = t1 - t0; out[(blockIdx.x * 16 + warp) * 2 + 1] = acc; }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int MODE> void run(int warps, const char* name) {
  long long* d; cudaMalloc(&d, 148 * 16 * 16);
  const int iters = 4000;
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s warps/CTA=%2d: %6.1f clk per operation (%s)\n", name, warps, double(h[0]) / iters,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  for (int w : {1, 4, 12}) {
    run<0>(w, "mbarrier.try_wait (complete) + tcgen05.fence");
    run<1>(w, "tcgen05.ld 64 cols + wait::ld");
    run<2>(w, "tcgen05.st 32 cols + wait::st");
    run<3>(w, "syncwarp + mbarrier.arrive (lane 0)");
    run<4>(w, "mbarrier.try_wait (complete) alone");
    run<5>(w, "mbarrier.test_wait (complete) alone");
    run<6>(w, "tcgen05.fence::after_thread_sync alone");
  }
  return 0;
}
