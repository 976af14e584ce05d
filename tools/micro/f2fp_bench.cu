// Which pipe does F2FP.BF16.F32.PACK_AB use? clk per warp-instruction with 1 warp per SMSP,
// alone and interleaved with MUFU.EX2.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float a[16]; unsigned p[16];
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 1e-3f + i * 0.01f; p[i] = 0; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0 || MODE == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
      if (MODE == 1 || MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 3) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("add.f32 %0, %0, %1;" : "+f"(a[(i + 3) & 15]) : "f"(a[(i + 7) & 15])); asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[(i + 5) & 15]) : "f"(a[(i + 9) & 15])); }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(const char* n) {
  float* d; long long* c; cudaMalloc(&d, 148 * 2048 * 4); cudaMalloc(&c, 8);
  for (int w : {1, 2}) {
    k<MODE><<<148, 128 * w>>>(d, 2000, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%s warps/SMSP=%d: %.2f clk per loop element per warp\n", n, w, double(h) / (2000 * 16.0));
  }
}
int main() { run<0>("F2FP only          "); run<1>("MUFU only          "); run<2>("F2FP + MUFU        "); run<3>("MUFU + FADD + FFMA "); return 0; }
