// How fast can the 3-stage RHS go when ptxas has registers to interleave?  Same occupancy as the step
// kernel (2 blocks x 4 warps per SM, forced by a shared-memory pad).  Variant 0: three stages one after
// the other in a rolled loop; variant 1: unrolled (ptxas free to interleave).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ics_wt_physicsengine_b200/csrc/wt_step_core.h"
struct SmemLu {
  double *p, *cp;
  __device__ __forceinline__ void put(int slot, double x, bool mask) { if (mask) p[slot * 32] = x; }
  __device__ __forceinline__ double get(int slot) const { return p[slot * 32]; }
  __device__ __forceinline__ void cput(int k, double x) { cp[k] = x; }
  __device__ __forceinline__ double cget(int k) const { return cp[k]; }
  __device__ __forceinline__ void csync() { __syncwarp(); }
  __device__ __forceinline__ void cadd(int, int) {}
  __device__ __forceinline__ double pvget(int) const { return 0.0; }
  __device__ __forceinline__ void pvput(int, double) {}
};
template <int UNROLL>
__global__ void __launch_bounds__(128, 2) k_rhs(double *out, int iters, int n, long long *clk) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WtGroup g = wt_make_group(n);
  SmemLu st;
  st.p = smem + warp * (32 * LK_N + 4 * CK_N) + lane;
  st.cp = smem + warp * (32 * LK_N + 4 * CK_N) + 32 * LK_N + (lane / n < 3 ? lane / n : 3) * CK_N;
  double par[WTP_NPAR] = {1e-14, 4.3e-7, 4.7e-11, 3e-8, 0.002, 0.01, 0.005, 0.2, 100.0, 1000.0, 5.0, 1.0};
  double bnd[WTB_NBND] = {5.0, 7.2, 1.0, 18.0, 0.1, 0.01, 0.1, 10.0, 20.0, 1.0};
  WtConstT<SmemLu> c = wt_make_const(&st, g, 0, par, bnd);
  double y0 = 7.0 + 0.01 * lane, y1 = 2.0 + 0.01 * lane, y2 = 20.0 + 0.1 * lane;
  double a0 = 0, a1 = 0, a2 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (UNROLL) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double f0, f1, f2; bool bad;
        wt_rhs(g, c, y0 + 1e-3 * i + a0 * 1e-9, y1 + 1e-3 * i, y2 + 1e-2 * i, f0, f1, f2, bad);
        a0 += f0 * wt_rk[14 + i]; a1 += f1 * wt_rk[17 + i]; a2 += f2 * wt_rk[20 + i];
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        double f0, f1, f2; bool bad;
        wt_rhs(g, c, y0 + 1e-3 * i + a0 * 1e-9, y1 + 1e-3 * i, y2 + 1e-2 * i, f0, f1, f2, bad);
        a0 += f0 * wt_rk[14 + i]; a1 += f1 * wt_rk[17 + i]; a2 += f2 * wt_rk[20 + i];
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2;
}
int main() {
  double *o; long long *clk; cudaMalloc(&o, 8 * 148 * 8 * 128); cudaMalloc(&clk, 64);
  const int iters = 2000, n = 10;
  size_t smem = 100 * 1024;  // 2 blocks per SM
  cudaFuncSetAttribute(k_rhs<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_rhs<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int v = 0; v < 2; ++v)
    for (int rep = 0; rep < 2; ++rep) {
      if (v == 0) k_rhs<0><<<148 * 2, 128, smem>>>(o, iters, n, clk); else k_rhs<1><<<148 * 2, 128, smem>>>(o, iters, n, clk);
      cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
      if (rep) printf("variant %d: %.0f cycles per RHS evaluation per warp (8 warps/SM) err=%s\n", v, (double)h / (iters * 3), cudaGetErrorString(cudaGetLastError()));
    }
  // one warp per SMSP alone
  for (int v = 0; v < 2; ++v) {
    if (v == 0) k_rhs<0><<<148, 128, 120 * 1024>>>(o, iters, n, clk); else k_rhs<1><<<148, 128, 120 * 1024>>>(o, iters, n, clk);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("variant %d, 4 warps/SM: %.0f cycles per RHS\n", v, (double)h / (iters * 3));
  }
  return 0;
}
