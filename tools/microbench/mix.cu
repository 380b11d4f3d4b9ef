// tools/microbench/mix.cu -- do half-rate pipes (FP64, FFMA2/IMAD, MUFU) overlap with each other and
// with full-rate ALU work inside one SMSP, or does every half-rate instruction hold the dispatch port?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define U 4
template <int ND, int NL, int NF2, int NI, int NM, int NF>
__global__ void __launch_bounds__(256) mix(double* out) {
  double d[U]; int l[U]; unsigned long long f2[U]; int im[U]; float m[U]; float ff[U];
  for (int j = 0; j < U; ++j) { d[j] = 1.0 + threadIdx.x * 1e-3 + j; l[j] = threadIdx.x + j; f2[j] = 0x3f8000003f800000ull + j; im[j] = threadIdx.x * 3 + j; m[j] = 1.5f + j; ff[j] = 1.25f + j; }
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < U; ++j) {
#pragma unroll
      for (int k = 0; k < ND; ++k) d[j] = fma(d[j], d[j], d[j]);
#pragma unroll
      for (int k = 0; k < NF2; ++k) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(f2[j]));
#pragma unroll
      for (int k = 0; k < NI; ++k) im[j] = im[j] * im[j] + im[j];
#pragma unroll
      for (int k = 0; k < NM; ++k) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(m[j]));
#pragma unroll
      for (int k = 0; k < NF; ++k) ff[j] = fmaf(ff[j], ff[j], ff[j]);
#pragma unroll
      for (int k = 0; k < NL; ++k) l[j] = (l[j] ^ (l[j] >> 3)) + 0x9e37;   // SHF/LOP3/IADD-class ALU work (2-3 instr)
    }
  }
  double s = 0; for (int j = 0; j < U; ++j) s += d[j] + l[j] + (double)f2[j] + im[j] + m[j] + ff[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ND, int NL, int NF2, int NI, int NM, int NF> void run(int sms) {
  double* o; int grid = sms * 8, block = 256; cudaMalloc(&o, 8 * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e30f;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); mix<ND, NL, NF2, NI, NM, NF><<<grid, block>>>(o); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = best * 1e-3 * clk * 1e3; double groups = (double)grid * block / 32 * ITERS * U;
  printf("DFMA=%d ALUgrp=%d FFMA2=%d IMAD=%d MUFU=%d FFMA=%d : %.2f clk per group per SMSP\n", ND, NL, NF2, NI, NM, NF, cyc * sms * 4 / groups);
  cudaFree(o);
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<1,0,0,0,0,0>(sms); run<0,1,0,0,0,0>(sms); run<0,2,0,0,0,0>(sms); run<0,4,0,0,0,0>(sms);
  run<1,1,0,0,0,0>(sms); run<1,2,0,0,0,0>(sms); run<1,4,0,0,0,0>(sms);
  run<0,0,1,0,0,0>(sms); run<1,0,1,0,0,0>(sms); run<1,0,2,0,0,0>(sms); run<2,0,1,0,0,0>(sms);
  run<0,0,0,1,0,0>(sms); run<1,0,0,1,0,0>(sms); run<0,0,1,1,0,0>(sms);
  run<0,0,0,0,1,0>(sms); run<4,0,0,0,1,0>(sms); run<4,0,4,0,1,0>(sms); run<4,4,4,0,1,0>(sms);
  run<0,0,0,0,0,1>(sms); run<1,0,0,0,0,1>(sms); run<1,0,0,0,0,2>(sms); run<1,0,0,0,0,4>(sms); run<0,0,1,0,0,2>(sms);
  run<4,2,4,0,1,4>(sms);
  return 0;
}
