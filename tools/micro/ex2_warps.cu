// MUFU.EX2 issue rate vs warps per SM sub-partition (1 CTA per SM, 32*4*w threads).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters, long long* clk) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* d; long long* c; cudaMalloc(&d, 148 * 2048 * 4); cudaMalloc(&c, 8);
  for (int w : {1, 2, 3, 4, 8}) {
    int iters = 2000;
    k<<<148, 128 * w>>>(d, iters, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("warps/SMSP=%d: %.2f clk per MUFU warp-instruction per warp, %.2f clk per instr per SMSP\n", w,
           double(h) / (iters * 16.0), double(h) / (iters * 16.0 * w));
  }
  return 0;
}
