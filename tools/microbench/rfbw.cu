// tools/microbench/rfbw.cu -- is the register file (2 banks x one 32-lane read per clock per SM
// sub-partition) a shared limit ACROSS pipes?  Independent DFMA and FFMA2/IMAD/LOP3 streams with
// many distinct register operands are mixed; if the time is the SUM of operand reads / 2 rather
// than the max over pipes, register-file read bandwidth is the binding resource.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE> __global__ void __launch_bounds__(256) k(double* out, double p) {
  double a[4], b[4], c[4];
  unsigned long long f[4], g[4], h[4];
  unsigned x[4], y[4], z[4];
  for (int j = 0; j < 4; ++j) {
    a[j] = 1.0 + threadIdx.x * 1e-3 + j; b[j] = 0.999 + j * 1e-4 + threadIdx.x * 1e-6; c[j] = 1e-3 * (j + 1) + threadIdx.x * 1e-7;
    f[j] = 0x3f8000003f800000ull + j + threadIdx.x; g[j] = 0x3f7f00003f7f0000ull + j; h[j] = 0x3a8000003a800000ull + j;
    x[j] = threadIdx.x * 7 + j; y[j] = threadIdx.x * 13 + j; z[j] = threadIdx.x * 29 + j;
  }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 0 || MODE == 2 || MODE == 4) {                 // DFMA, 2 distinct 64-bit regs (4 reads)
        a[j] = fma(a[j], b[j], a[j]);
      }
      if (MODE == 1 || MODE == 2) {                              // FFMA2, 3 distinct 64-bit regs (6 reads)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f[j]) : "l"(g[j]), "l"(h[j]));
      }
      if (MODE == 3 || MODE == 4) {                              // 2 x LOP3 with 3 distinct regs (6 reads)
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(y[j]), "r"(z[j]));
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[j]) : "r"(z[j]), "r"(x[j]));
      }
      if (MODE == 5 || MODE == 6) {                              // DFMA, 3 distinct (6 reads)
        a[j] = fma(a[j], b[j], c[j]);
      }
      if (MODE == 6) {                                           // + FFMA2 3 distinct (6 reads)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(f[j]) : "l"(g[j]), "l"(h[j]));
      }
    }
  }
  double s = 0; for (int j = 0; j < 4; ++j) s += a[j] + b[j] + c[j] + (double)f[j] + x[j] + y[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(int sms, const char* name) {
  double* o; int grid = sms * 8, block = 256; cudaMalloc(&o, 8 * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<grid, block>>>(o, 0.9999); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = best * 1e-3 * clk * 1e3; double n = (double)grid * block / 32 * ITERS * 4;
  printf("%-52s %.2f clk per group per SMSP\n", name, cyc * sms * 4 / n);
}
int main() { int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(sms, "DFMA a,b,a (4 reads)");
  run<1>(sms, "FFMA2 f,g,h (6 reads)");
  run<2>(sms, "DFMA a,b,a + FFMA2 f,g,h (10 reads)");
  run<3>(sms, "2 x LOP3 x,y,z (6 reads)");
  run<4>(sms, "DFMA a,b,a + 2 x LOP3 x,y,z (10 reads)");
  run<5>(sms, "DFMA a,b,c (6 reads)");
  run<6>(sms, "DFMA a,b,c + FFMA2 f,g,h (12 reads)");
  return 0; }
