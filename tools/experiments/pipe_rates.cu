// Micro-benchmark (development only): issue cost per SM sub-partition of the instruction kinds the fused kernel is made of,
// at 1/2/3/4 warps per scheduler, 8 independent dependency chains per thread.  nvcc -arch=sm_100a -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
#define ITER 512
template <int K> __device__ __forceinline__ void body(f2 (&a)[8], f2 b, f2 c, float (&s)[8], float u, float v, const f2 (&x)[8], const f2 (&y)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (K == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v));
    if (K == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    if (K == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
    if (K == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b));
    if (K == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v));
    if (K == 5) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v)); }
    if (K == 6) asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(u));
    if (K == 7) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(u));
    if (K == 8) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(b)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v)); }
    if (K == 9) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[(i + 4) & 7]) : "l"(b)); }
    if (K == 10) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c)); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v)); asm volatile("set.gt.f32.f32 %0, %0, %1;" : "+f"(s[(i + 4) & 7]) : "f"(u)); }
    if (K == 11) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(b), "l"(c));  // accumulate form: two sources shared by all
    if (K == 15) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(x[i]), "l"(y[i]));   // three distinct register pairs
    if (K == 16) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x[i]));
    if (K == 17) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[i]) : "l"(x[i]), "l"(y[i])); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(u), "f"(v)); }
    if (K == 18) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(s[i]) : "f"(__uint_as_float((unsigned)x[i])), "f"(__uint_as_float((unsigned)y[i])));
    if (K == 19) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x[i])); asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(__uint_as_float((unsigned)y[i])), "f"(__uint_as_float((unsigned)x[(i+1)&7]))); }
    if (K == 20) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(__uint_as_float((unsigned)y[i]))); }
    if (K == 21) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x[i])); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[(i+4)&7]) : "l"(y[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(__uint_as_float((unsigned)y[i]))); }
    if (K == 22) { unsigned p; asm volatile("{ .reg .pred q; setp.gt.f32 q, %1, %2; selp.u32 %0, 1, 0, q; }" : "=r"(p) : "f"(s[i]), "f"(u)); s[i] = __uint_as_float(p); }
    if (K == 12) asm volatile("min.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(u));
    if (K == 13) { float t; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(s[i])); s[i] = t; }
    if (K == 14) { asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c)); float t; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(s[i])); s[i] = t; }
  }
}
template <int K> __global__ void k(float* out, long long* cyc, f2 b, f2 c, float u, float v) {
  f2 a[8]; float s[8];
  f2 x[8], y[8];
  for (int i = 0; i < 8; ++i) { a[i] = b + threadIdx.x + i; s[i] = u + threadIdx.x + i; x[i] = b + 3 * threadIdx.x + i; y[i] = c + 5 * threadIdx.x + i; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) body<K>(a, b, c, s, u, v, x, y);
  const long long t1 = clock64();
  __syncthreads();
  float r = 0; for (int i = 0; i < 8; ++i) r += s[i] + (float)a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K> void run(const char* name, int per) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  printf("%-34s", name);
  for (int w = 1; w <= 4; ++w) {
    k<K><<<148, 128 * w>>>(out, cyc, 0x3f8000003f800000ull, 0x3f0000003f000000ull, 1.0f, 0.5f);
    k<K><<<148, 128 * w>>>(out, cyc, 0x3f8000003f800000ull, 0x3f0000003f000000ull, 1.0f, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < 148; ++i) m += h[i]; m /= 148;
    printf("  %dw/sched: %.2f cyc/instr/SMSP", w, m / (double)(ITER * 8 * per * w));
  }
  printf("\n"); cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA (3 regs)", 1); run<1>("FFMA2 a=a*b+c", 1); run<11>("FFMA2 a=b*c+a", 1); run<2>("FADD2", 1); run<3>("FMUL2", 1);
  run<15>("FFMA2 a=x*y+a (3 distinct pairs)", 1); run<16>("FADD2 a=a+x (distinct)", 1); run<18>("FFMA s=x*y+s (3 distinct)", 1); run<17>("FFMA2(distinct)+FMNMX3 (per instr)", 2);
  run<19>("FADD2(distinct)+FMNMX3(distinct)", 2); run<20>("FADD2(distinct)+FMNMX", 2); run<21>("2 FADD2 + FMNMX (per instr)", 3); run<22>("FSETP+SEL (per pair)", 1);
  run<7>("FADD", 1); run<4>("FMNMX3", 1); run<12>("FMNMX", 1); run<6>("FSET", 1); run<13>("MUFU.SQRT", 1);
  run<5>("FFMA2+FMNMX3 (per instr)", 2); run<8>("FADD2+FMNMX3 (per instr)", 2); run<9>("FFMA2+FADD2 (per instr)", 2);
  run<10>("FFMA2+FMNMX3+FSET (per instr)", 3); run<14>("FFMA2+MUFU (per instr)", 2);
  return 0;
}
