// tools/microbench/lat.cu -- dependent-issue latency (one warp per SM, one chain)
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(double* out, double p, long long* cyc) {
  double a = 1.0 + threadIdx.x * 1e-3, b = 0.999, c = 1e-3; float f = 1.5f + threadIdx.x; unsigned long long v = 0x3f8000003f800000ull; unsigned u = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a = fma(a, b, c);
      if (MODE == 1) a = a * b;
      if (MODE == 2) a = a + b;
      if (MODE == 3) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(a));
      if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(v));
      if (MODE == 5) f = fmaf(f, f, f);
      if (MODE == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
      if (MODE == 7) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a) : "f"(f)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(a)); }
      if (MODE == 8) asm volatile("lop3.b32 %0, %0, %0, %0, 0x96;" : "+r"(u));
      if (MODE == 9) u = u * u + u;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + f + (double)v + u;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name, int per) {
  double* o; long long* c; cudaMalloc(&o, 8 * 148 * 32); cudaMalloc(&c, 8);
  k<MODE><<<148, 32>>>(o, 0.9999, c); k<MODE><<<148, 32>>>(o, 0.9999, c);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-22s latency %.2f clk\n", name, (double)h / (ITERS * 8.0 * per));
}
int main() { run<0>("DFMA", 1); run<1>("DMUL", 1); run<2>("DADD", 1); run<3>("MUFU.RSQ64H", 1); run<4>("FFMA2", 1); run<5>("FFMA", 1); run<6>("MUFU.LG2", 1); run<7>("F2F 32->64->32 (pair)", 1); run<8>("LOP3", 1); run<9>("IMAD", 1); return 0; }
