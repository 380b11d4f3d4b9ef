// tools/microbench/rf.cu -- does DFMA with three DISTINCT 64-bit register operands issue slower than
// fma(x,x,x)?  (register-file read ports / operand collector)
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE> __global__ void __launch_bounds__(256) k(double* out, double p, double q) {
  double a[4], b[4], c[4];
  for (int j = 0; j < 4; ++j) { a[j] = 1.0 + threadIdx.x * 1e-3 + j; b[j] = 0.999 + j * 1e-4 + threadIdx.x * 1e-6; c[j] = 1e-3 * (j + 1) + threadIdx.x * 1e-7; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 0) { a[j] = fma(a[j], a[j], a[j]); a[j] = fma(a[j], a[j], a[j]); a[j] = fma(a[j], a[j], a[j]); }
      if (MODE == 1) { a[j] = fma(a[j], b[j], c[j]); b[j] = fma(b[j], c[j], a[j]); c[j] = fma(c[j], a[j], b[j]); }   // 3 distinct regs
      if (MODE == 2) { a[j] = fma(a[j], b[j], a[j]); b[j] = fma(b[j], c[j], b[j]); c[j] = fma(c[j], a[j], c[j]); }   // 2 distinct
      if (MODE == 3) { a[j] = fma(a[j], p, c[j]); b[j] = fma(b[j], p, a[j]); c[j] = fma(c[j], q, b[j]); }           // 2 regs + const
      if (MODE == 4) { a[j] = a[j] * b[j]; b[j] = b[j] * c[j]; c[j] = c[j] * a[j]; }                               // DMUL 2 distinct
      if (MODE == 5) { a[j] = a[j] + b[j]; b[j] = b[j] + c[j]; c[j] = c[j] + a[j]; }
    }
  }
  double s = 0; for (int j = 0; j < 4; ++j) s += a[j] + b[j] + c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(int sms, const char* name) {
  double* o; int grid = sms * 8, block = 256; cudaMalloc(&o, 8 * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE><<<grid, block>>>(o, 0.9999, 1.0001); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = best * 1e-3 * clk * 1e3; double n = (double)grid * block / 32 * ITERS * 12;
  printf("%-28s %.2f clk per warp-instr per SMSP\n", name, cyc * sms * 4 / n);
}
int main() { int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(sms, "DFMA x,x,x"); run<1>(sms, "DFMA a,b,c (3 distinct)"); run<2>(sms, "DFMA a,b,a (2 distinct)"); run<3>(sms, "DFMA a,const,c"); run<4>(sms, "DMUL a,b"); run<5>(sms, "DADD a,b"); return 0; }
