// tools/microbench/disp.cu -- does a 64-bit multi-operand instruction hold the SMSP dispatch port
// while it reads its operands?  K independent single-register LOP3 (immediate operands) are mixed
// with one DFMA per group; if time = DFMA_cost + K the port is held, if time = max(DFMA_cost, 1+K) not.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE, int K> __global__ void __launch_bounds__(256) k(double* out, double p) {
  double a[4], b[4], c[4]; unsigned l[4][4];
  for (int j = 0; j < 4; ++j) { a[j] = 1.0 + threadIdx.x * 1e-3 + j; b[j] = 0.999 + j * 1e-4; c[j] = 1e-3 * (j + 1);
    for (int m = 0; m < 4; ++m) l[j][m] = threadIdx.x * 7 + j + m; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 1) a[j] = fma(a[j], a[j], a[j]);
      if (MODE == 2) a[j] = fma(a[j], b[j], a[j]);
      if (MODE == 3) { a[j] = fma(a[j], b[j], c[j]); }
      if (MODE == 4) a[j] = fma(a[j], p, a[j]);
#pragma unroll
      for (int m = 0; m < K; ++m) asm volatile("lop3.b32 %0, %0, 0x55555555, 0x33333333, 0x96;" : "+r"(l[j][m & 3]));
    }
  }
  double s = 0; for (int j = 0; j < 4; ++j) s += a[j] + b[j] + c[j] + l[j][0] + l[j][1] + l[j][2] + l[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int K> void run(int sms) {
  double* o; int grid = sms * 8, block = 256; cudaMalloc(&o, 8 * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); k<MODE, K><<<grid, block>>>(o, 0.9999); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = best * 1e-3 * clk * 1e3; double n = (double)grid * block / 32 * ITERS * 4;
  const char* names[] = {"none", "DFMA x,x,x", "DFMA a,b,a", "DFMA a,b,c", "DFMA a,const,a"};
  printf("%-16s + %d LOP3imm : %.2f clk per group per SMSP\n", names[MODE], K, cyc * sms * 4 / n);
}
int main() { int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0,1>(sms); run<0,2>(sms); run<0,4>(sms);
  run<1,0>(sms); run<1,1>(sms); run<1,2>(sms); run<1,4>(sms);
  run<2,0>(sms); run<2,1>(sms); run<2,2>(sms); run<2,4>(sms);
  run<3,0>(sms); run<3,1>(sms); run<3,2>(sms); run<3,4>(sms);
  run<4,0>(sms); run<4,2>(sms); run<4,4>(sms);
  return 0; }
