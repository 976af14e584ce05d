// Issue-rate microbenchmark for the instructions of the attention softmax on sm_100a:
// clk per warp-instruction per SM sub-partition, for w warps per sub-partition.
// Each test body has 16 independent dependency chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define REP16(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15)

template <int MODE>
__global__ void k(float* out, int iters, long long* clk) {
  float a[16]; uint32_t p[16]; unsigned long long d[16];
  for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 1e-3f + i * 0.01f; p[i] = 0x38003800u + threadIdx.x + i; d[i] = (unsigned long long)(p[i]) << 32 | p[i]; }
  float c0 = 1.0001f, c1 = 0.001f;
  unsigned long long cc = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c0);
  unsigned long long cd = ((unsigned long long)__float_as_uint(c1) << 32) | __float_as_uint(c1);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#define EX2(i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#define EX2H(i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p[i]));
#define EX2B(i) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(p[i]));
#define CVTB(i) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
#define CVTH(i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(a[i]), "f"(a[(i + 1) & 15]));
#define FMA1(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c0), "f"(c1));
#define FADD1(i) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c1));
#define FMA2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(cc), "l"(cd));
#define FADD2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(cd));
#define FMUL2(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(cc));
#define MAX2(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]));
#define MAX3(i) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(a[(i + 2) & 15]));
#define PRMT(i) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(p[i]) : "r"(__float_as_uint(a[i])), "r"(__float_as_uint(a[(i + 1) & 15])));
#define IADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(p[i]) : "r"(p[(i + 1) & 15]));
#define IMAD(i) asm volatile("mad.lo.u32 %0, %1, 8388608, %0;" : "+r"(p[i]) : "r"(p[(i + 1) & 15]));
#define SHL(i) asm volatile("shl.b32 %0, %0, 23;" : "+r"(p[i]));
#define HFMA2(i) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(p[i]) : "r"(p[(i + 1) & 15]));
#define HADD2F(i) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(p[i]) : "r"(p[(i + 1) & 15]));
    if (MODE == 0) { REP16(EX2) }
    if (MODE == 1) { REP16(EX2H) }
    if (MODE == 2) { REP16(EX2B) }
    if (MODE == 3) { REP16(CVTB) }
    if (MODE == 4) { REP16(CVTH) }
    if (MODE == 5) { REP16(FMA1) }
    if (MODE == 6) { REP16(FADD1) }
    if (MODE == 7) { REP16(FMA2) }
    if (MODE == 8) { REP16(FADD2) }
    if (MODE == 9) { REP16(MAX2) }
    if (MODE == 10) { REP16(MAX3) }
    if (MODE == 11) { REP16(PRMT) }
    if (MODE == 12) { REP16(IADD) }
    if (MODE == 13) { REP16(IMAD) }
    if (MODE == 14) { REP16(SHL) }
    if (MODE == 15) { REP16(HFMA2) }
    // mixes: one "element group" = what the softmax issues per MUFU
#define MIX_A(i) EX2(i) FMA1(i) FADD1((i + 8) & 15)
    if (MODE == 16) { REP16(MIX_A) }                       // MUFU + FFMA + FADD
#define MIX_B(i) EX2(i) CVTB((i + 8) & 15)
    if (MODE == 17) { REP16(MIX_B) }                       // MUFU + F2FP (1:1)
#define MIX_C(i) EX2(i) PRMT((i + 8) & 15)
    if (MODE == 18) { REP16(MIX_C) }                       // MUFU + PRMT
#define MIX_D(i) EX2(i) FMA2(i) FADD2((i + 8) & 15)
    if (MODE == 19) { REP16(MIX_D) }                       // MUFU + FFMA2 + FADD2
#define MIX_E(i) EX2H(i) CVTH((i + 8) & 15)
    if (MODE == 20) { REP16(MIX_E) }                       // MUFU.f16x2 + F2FP.f16x2
#define MIX_F(i) EX2(i) FMA1(i) FMA1((i + 4) & 15) FMA1((i + 8) & 15) FADD1((i + 12) & 15) IMAD(i)
    if (MODE == 21) { REP16(MIX_F) }                       // MUFU + 5 fma-pipe ops + IMAD
    if (MODE == 22) { REP16(FMUL2) }
    if (MODE == 23) { REP16(HADD2F) }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(p[i]) + float(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int MODE> void run(const char* n, int instr_per_rep) {
  float* d; long long* c; cudaMalloc(&d, 148 * 2048 * 4); cudaMalloc(&c, 8);
  printf("%-34s", n);
  for (int w : {1, 2, 4}) {
    k<MODE><<<148, 128 * w>>>(d, 2000, c); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("  w=%d: %6.2f", w, double(h) / (2000 * 16.0 * w));
  }
  printf("   clk per group per SMSP (%d instr/group)  %s\n", instr_per_rep, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d); cudaFree(c);
}
int main() {
  run<0>("MUFU.EX2 f32", 1); run<1>("ex2.f16x2", 1); run<2>("ex2.bf16x2", 1);
  run<3>("cvt.rn.bf16x2.f32 (F2FP)", 1); run<4>("cvt.rn.f16x2.f32 (F2FP)", 1);
  run<5>("FFMA", 1); run<6>("FADD", 1); run<7>("fma.f32x2", 1); run<8>("add.f32x2", 1); run<22>("mul.f32x2", 1);
  run<9>("max.f32 (2 in)", 1); run<10>("max.f32 (3 in)", 1); run<11>("PRMT", 1); run<12>("IADD", 1);
  run<13>("IMAD imm", 1); run<14>("SHL", 1); run<15>("HFMA2", 1); run<23>("add.bf16x2", 1);
  run<16>("MUFU+FFMA+FADD", 3); run<17>("MUFU+F2FP", 2); run<18>("MUFU+PRMT", 2);
  run<19>("MUFU+FFMA2+FADD2", 3); run<20>("ex2.f16x2+F2FP.f16x2", 2); run<21>("MUFU+3FFMA+FADD+IMAD", 6);
  return 0;
}
