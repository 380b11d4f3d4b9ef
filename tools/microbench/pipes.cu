// tools/microbench/pipes.cu -- per-SM issue rates of the instructions the path kernel leans on
// (development aid; results recorded in profiles/).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define UNROLL 8

template <class F> __global__ void __launch_bounds__(256) bench(double* out, F f, double seed) {
  double x[UNROLL];
  for (int j = 0; j < UNROLL; ++j) x[j] = seed + threadIdx.x * 1e-3 + j;
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) x[j] = f(x[j]);
  }
  double s = 0; for (int j = 0; j < UNROLL; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct OpDfmaReg { __device__ double operator()(double x) const { return fma(x, x, x); } };
struct OpDfmaConst { double a, b; __device__ double operator()(double x) const { return fma(x, a, b); } };
struct OpDmul { __device__ double operator()(double x) const { return x * x; } };
struct OpDadd { __device__ double operator()(double x) const { return x + x; } };
struct OpRcp64h { __device__ double operator()(double x) const { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; } };
struct OpRsq64h { __device__ double operator()(double x) const { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; } };
struct OpF2F { __device__ double operator()(double x) const { float f; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(x)); double y; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(f)); return y; } };
struct OpF2Fwiden { __device__ double operator()(double x) const { float f = __int_as_float(__double2loint(x) | 0x3f800000); double y; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(y) : "f"(f)); return y + x; } };
struct OpI2Fd { __device__ double operator()(double x) const { unsigned long long u = (unsigned long long)__double_as_longlong(x); double y; asm volatile("cvt.rn.f64.u64 %0, %1;" : "=d"(y) : "l"(u)); return y; } };
struct OpLg2 { __device__ double operator()(double x) const { float f = __int_as_float(__double2loint(x)); float g; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(g) : "f"(f)); return __hiloint2double(__double2hiint(x), __float_as_int(g)); } };
struct OpFfma2 { __device__ double operator()(double x) const { unsigned long long a = (unsigned long long)__double_as_longlong(x), d; asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(d) : "l"(a)); return __longlong_as_double((long long)d); } };
struct OpFfma { __device__ double operator()(double x) const { float f = __int_as_float(__double2loint(x)); f = fmaf(f, f, f); return __hiloint2double(__double2hiint(x), __float_as_int(f)); } };
struct OpImad { __device__ double operator()(double x) const { int a = __double2loint(x); a = a * a + a; return __hiloint2double(__double2hiint(x), a); } };
struct OpLop { __device__ double operator()(double x) const { int a = __double2loint(x), b = __double2hiint(x); a = (a ^ b) & 0x5555 | a; return __hiloint2double(b, a); } };

template <class F> void run(const char* name, F f, int ops_per_call, int sms) {
  double* d; int grid = sms * 8, block = 256;
  cudaMalloc(&d, sizeof(double) * grid * block);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0); bench<<<grid, block>>>(d, f, 1.000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
  }
  double warp_instr = (double)grid * block / 32 * ITERS * UNROLL * ops_per_call;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cycles = best * 1e-3 * clk * 1e3;
  printf("%-14s %8.3f ms  %.3f warp-instr/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, best,
         warp_instr / cycles / sms, cycles * sms * 4 / warp_instr);
  cudaFree(d);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run("DFMA rrr", OpDfmaReg(), 1, sms);
  run("DFMA rcc", OpDfmaConst{1.0000001, 1e-9}, 1, sms);
  run("DMUL", OpDmul(), 1, sms);
  run("DADD", OpDadd(), 1, sms);
  run("MUFU.RCP64H", OpRcp64h(), 1, sms);
  run("MUFU.RSQ64H", OpRsq64h(), 1, sms);
  run("F2F 64<->32 x2", OpF2F(), 2, sms);
  run("F2F.F64.F32+DADD", OpF2Fwiden(), 2, sms);
  run("I2F.F64.U64", OpI2Fd(), 1, sms);
  run("MUFU.LG2", OpLg2(), 1, sms);
  run("FFMA2", OpFfma2(), 1, sms);
  run("FFMA", OpFfma(), 1, sms);
  run("IMAD", OpImad(), 1, sms);
  run("LOP3 x2", OpLop(), 2, sms);
  return 0;
}
