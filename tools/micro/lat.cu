// Microbenchmarks that inform the step-kernel design: dependent-issue latency of DFMA / SHFL / LDS /
// MUFU.RCP64H, and FP64 pipe throughput as a function of warps per SM sub-partition and ILP.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_lat_dfma(double *o, int it, double a, double b, long long *clk) {
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, a, b);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  if (x == 1.2345) o[0] = x;
}
__global__ void k_lat_shfl(double *o, int it, long long *clk) {
  int x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = __shfl_sync(0xffffffffu, x, (x + 1) & 31);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  if (x == 12345) o[0] = x;
}
__global__ void k_lat_lds(double *o, int it, long long *clk) {
  __shared__ int s[64];
  s[threadIdx.x & 63] = (threadIdx.x + 1) & 31;
  __syncthreads();
  int x = threadIdx.x & 31;
  long long t0 = clock64();
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = s[x];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  if (x == 12345) o[0] = x;
}
__global__ void k_lat_rcp(double *o, int it, long long *clk) {
  double x = 1.0 + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  if (x == 1.2345) o[0] = x;
}
// throughput: ILP chains of DFMA, W warps per block, one block per SM
template <int ILP>
__global__ void k_tp_dfma(double *o, int it, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 1.2345) o[0] = s;
}
// shuffle throughput: ILP independent 32-bit shuffles
template <int ILP>
__global__ void k_tp_shfl(double *o, int it) {
  int x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) x[j] = __shfl_sync(0xffffffffu, x[j], (threadIdx.x + 1) & 31);
  }
  int s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j];
  if (s == 12345) o[0] = s;
}
// mixed: per DFMA, K other (integer) instructions, to see issue-slot sharing between FP64 and ALU
template <int ILP, int K>
__global__ void k_tp_mix(double *o, int it, double a, double b, int c) {
  double x[ILP]; int y[ILP];
#pragma unroll
  for (int j = 0; j < ILP; ++j) { x[j] = threadIdx.x + j; y[j] = threadIdx.x * j; }
  for (int i = 0; i < it; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < ILP; ++j) {
        x[j] = fma(x[j], a, b);
#pragma unroll
        for (int k = 0; k < K; ++k) y[j] = (y[j] ^ c) + (y[j] >> 3);
      }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; ++j) s += x[j] + y[j];
  if (s == 1.2345) o[0] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double *o; long long *clk; cudaMalloc(&o, 64); cudaMalloc(&clk, 64);
  long long h;
  int it = 4096;
  k_lat_dfma<<<1, 32>>>(o, it, 1.0000001, 1e-9, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("DFMA dependent latency: %.2f cycles\n", (double)h / (it * 16));
  k_lat_shfl<<<1, 32>>>(o, it, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("SHFL dependent latency: %.2f cycles\n", (double)h / (it * 16));
  k_lat_lds<<<1, 32>>>(o, it, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("LDS dependent latency: %.2f cycles\n", (double)h / (it * 16));
  k_lat_rcp<<<1, 32>>>(o, it, clk); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("MUFU.RCP64H dependent latency: %.2f cycles\n", (double)h / (it * 16));
  int dev; cudaGetDevice(&dev); cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  int sms = pr.multiProcessorCount; double ghz = pr.clockRate * 1e-6;
  printf("SMs %d clock %.3f GHz\n", sms, ghz);
  it = 20000;
#define TP(ILP, W) { float ms = timeit([&] { k_tp_dfma<ILP><<<sms, 32 * W>>>(o, it, 1.0000001, 1e-9); }); \
    double inst = (double)it * 8 * ILP * W; /* warp-instr per SM */ \
    printf("DFMA tp  ILP=%d warps/SM=%2d: %.3f warp-DFMA/clk/SM  (%.1f TFLOP/s)\n", ILP, W, inst / (ms * 1e-3 * ghz * 1e9), inst * 64 * sms / (ms * 1e-3) * 1e-12); }
  TP(1, 4) TP(2, 4) TP(4, 4) TP(8, 4) TP(1, 8) TP(2, 8) TP(4, 8) TP(8, 8) TP(1, 12) TP(2, 12) TP(4, 12) TP(1, 16) TP(2, 16) TP(4, 16) TP(1, 32) TP(2, 32)
#define TS(ILP, W) { float ms = timeit([&] { k_tp_shfl<ILP><<<sms, 32 * W>>>(o, it); }); \
    double inst = (double)it * 8 * ILP * W; \
    printf("SHFL tp  ILP=%d warps/SM=%2d: %.3f warp-SHFL/clk/SM\n", ILP, W, inst / (ms * 1e-3 * ghz * 1e9)); }
  TS(1, 8) TS(4, 8) TS(8, 8) TS(4, 16) TS(8, 32)
#define TM(ILP, K, W) { float ms = timeit([&] { k_tp_mix<ILP, K><<<sms, 32 * W>>>(o, it, 1.0000001, 1e-9, 12345); }); \
    double inst = (double)it * 8 * ILP * W; \
    printf("MIX tp  ILP=%d K=%d warps/SM=%2d: %.3f warp-DFMA/clk/SM\n", ILP, K, W, inst / (ms * 1e-3 * ghz * 1e9)); }
  TM(4, 1, 8) TM(4, 2, 8) TM(4, 1, 16) TM(4, 2, 16)
  return 0;
}
