// Instruction-cache capacity probe: a loop whose straight-line body is N kB of FFMA code, run by 8 warps per
// SM (2 blocks x 4 warps, like the step kernel) that start at staggered offsets.  Reports cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>
template <int KB>
__global__ void __launch_bounds__(128) k_body(float *o, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < KB * 8; ++r) {  // 8 FFMA = 128 B of code per r  -> KB kB per loop body
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678f) o[0] = s;
}
template <int KB> void run(float *o, int sms, double ghz, int blocks_per_sm) {
  int iters = 4096 / KB; if (iters < 4) iters = 4;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t smem = blocks_per_sm == 2 ? 100 * 1024 : 0;
  cudaFuncSetAttribute(k_body<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k_body<KB><<<sms * blocks_per_sm, 128, smem>>>(o, iters, 1.0001f, 1e-9f); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_body<KB><<<sms * blocks_per_sm * 4, 128, smem>>>(o, iters, 1.0001f, 1e-9f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double inst_per_smsp = (double)iters * KB * 64 * 4 /*waves*/ * blocks_per_sm;  // each SMSP runs blocks_per_sm warps x 4 waves
  printf("body %4d kB: %.2f cycles per warp-instruction per SMSP (ideal 1.0), err=%s\n", KB, ms * 1e-3 * ghz * 1e9 / inst_per_smsp, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  float *o; cudaMalloc(&o, 64);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  int sms = pr.multiProcessorCount; double ghz = pr.clockRate * 1e-6;
  run<8>(o, sms, ghz, 2); run<16>(o, sms, ghz, 2); run<24>(o, sms, ghz, 2); run<32>(o, sms, ghz, 2); run<48>(o, sms, ghz, 2);
  run<64>(o, sms, ghz, 2); run<96>(o, sms, ghz, 2); run<128>(o, sms, ghz, 2); run<192>(o, sms, ghz, 2); run<256>(o, sms, ghz, 2);
  return 0;
}
