// Micro-benchmark: MUFU ex2 throughput for f32 / f16x2 / bf16x2 operands on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  unsigned h0 = 0x38003800u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}
template <int MODE> void run(const char* name) {
  float* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  int iters = 20000;
  k<MODE><<<148 * 4, 512>>>(d, 100);
  cudaEventRecord(s); k<MODE><<<148 * 4, 512>>>(d, iters); cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double instr = 148.0 * 4 * 512 * 4.0 * iters;  // thread-level instructions
  printf("%s: %.3f ms, %.2f G thread-instr/s, %.1f thread-instr/clk/SM @1.9GHz (x2 elements for packed)\n",
         name, ms, instr / ms / 1e6, instr / (ms * 1e-3) / 148 / 1.9e9);
}
int main() { run<0>("ex2.f32   "); run<1>("ex2.f16x2 "); run<2>("ex2.bf16x2"); return 0; }
