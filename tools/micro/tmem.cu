// tmem.cu -- can tensor memory (TMEM) serve as a lane-private scratchpad for fp64 data of a non-MMA kernel?
//
// The step kernel keeps ~81 doubles of LU multipliers per lane in shared memory, which caps it at 2 blocks per SM.
// TMEM (256 KB per SM, 128 lanes x 512 32-bit columns) is idle in a kernel without tcgen05.mma; with the 32x32b
// shape of tcgen05.ld / tcgen05.st, thread i of a warp reads/writes consecutive columns of TMEM lane
// 32*(warp%4)+i, i.e. exactly a lane-private array.  This program checks
//   (1) correctness: every thread of every resident block stores a pattern into its columns and reads it back
//       after all blocks of the SM have written theirs (several blocks per SM, two allocations per block);
//   (2) dependent latency of tcgen05.ld (+wait) against ld.shared;
//   (3) throughput of x2 / x4 / x8 loads with 4..16 warps per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tmem tmem.cu && ./tmem
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t tmem_alloc(uint32_t *slot, int cols) {
  // one warp allocates; result lands in shared memory
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot);
  if (cols == 32) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" :: "r"(sa) : "memory");
  else if (cols == 64) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"(sa) : "memory");
  else if (cols == 128) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(sa) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" :: "r"(sa) : "memory");
  return 0;
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int cols) {
  if (cols == 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" :: "r"(addr) : "memory");
  else if (cols == 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(addr) : "memory");
  else if (cols == 128) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(addr) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t a, double x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(a), "r"(__double2loint(x)), "r"(__double2hiint(x)) : "memory");
}
__device__ __forceinline__ double tmem_ld1(uint32_t a) {
  int lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a) : "memory");
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tmem_ld4(uint32_t a, double *x) {
  int r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}
__device__ __forceinline__ void tmem_ld2(uint32_t a, double *x) {
  int r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
#pragma unroll
  for (int i = 0; i < 2; ++i) x[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}
__device__ __forceinline__ void tmem_st4(uint32_t a, const double *x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(a), "r"(__double2loint(x[0])), "r"(__double2hiint(x[0])), "r"(__double2loint(x[1])), "r"(__double2hiint(x[1])),
                  "r"(__double2loint(x[2])), "r"(__double2hiint(x[2])), "r"(__double2loint(x[3])), "r"(__double2hiint(x[3])) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- (1) correctness -----------------------------------------------------------------------------------
// WARPS warps per block; each block allocates C1 (+ C2) columns; every thread writes f(block, thread, k) into
// double slot k of its lane, spins a little (so that co-resident blocks interleave), reads back and compares.
__global__ void k_correct(int c1, int c2, int *errors, unsigned *bases, int spin) {
  __shared__ uint32_t sa[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&sa[0], c1);
    if (c2) tmem_alloc(&sa[1], c2);
    tmem_relinquish();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t b1 = sa[0], b2 = c2 ? sa[1] : 0;
  if (threadIdx.x == 0 && blockIdx.x < 64) { bases[2 * blockIdx.x] = b1; bases[2 * blockIdx.x + 1] = b2; }
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const int wq = warp >> 2;                 // warps w and w+4 share TMEM lanes: split the columns between them
  const int nq = (blockDim.x / 32 + 3) / 4; // warps per lane quarter
  const int n1 = c1 / 2 / nq, n2 = c2 / 2 / nq;  // double slots per thread in each allocation
  for (int k = 0; k < n1; ++k) tmem_st1(b1 + lane_off + 2 * (wq * n1 + k), (double)blockIdx.x * 1e6 + threadIdx.x * 1e3 + k + 0.25);
  for (int k = 0; k < n2; ++k) tmem_st1(b2 + lane_off + 2 * (wq * n2 + k), -((double)blockIdx.x * 1e6 + threadIdx.x * 1e3 + k + 0.5));
  tmem_wait_st();
  long long t0 = clock64();
  while (clock64() - t0 < spin) { }
  int bad = 0;
  for (int k = 0; k < n1; ++k) {
    double v = tmem_ld1(b1 + lane_off + 2 * (wq * n1 + k));
    tmem_wait_ld();
    bad += v != (double)blockIdx.x * 1e6 + threadIdx.x * 1e3 + k + 0.25;
  }
  for (int k = 0; k + 3 < n2; k += 4) {
    double v[4];
    tmem_ld4(b2 + lane_off + 2 * (wq * n2 + k), v);
    tmem_wait_ld();
    for (int i = 0; i < 4; ++i) bad += v[i] != -((double)blockIdx.x * 1e6 + threadIdx.x * 1e3 + k + i + 0.5);
  }
  if (bad) atomicAdd(errors, bad);
  (void)lane;
  __syncthreads();
  if (warp == 0) {
    tmem_dealloc(b1, c1);
    if (c2) tmem_dealloc(b2, c2);
  }
}

// ---- (2) dependent latency: x = fma(load(k), x, 1) chains ------------------------------------------------
__global__ void k_lat(int iters, double *out, long long *cyc, int mode) {
  __shared__ uint32_t sa;
  __shared__ double sm[64 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&sa, 128); tmem_relinquish(); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = sa + ((uint32_t)((warp & 3) * 32) << 16);
  for (int k = 0; k < 64; ++k) { tmem_st1(base + 2 * k, 1.0 + 1e-9 * k); sm[k * 32 + lane] = 1.0 + 1e-9 * k; }
  tmem_wait_st();
  __syncwarp();
  double x = 1.0;
  long long t0 = clock64();
  if (mode == 0) {        // TMEM: the slot index depends on the previous value (through a uniform-looking int)
    int k = 0;
    for (int i = 0; i < iters; ++i) {
      double v = tmem_ld1(base + 2 * k);
      tmem_wait_ld();
      x = fma(v, x, 1e-12);
      k = (k + 1 + (__double2loint(x) & 0)) & 63;  // data dependence without changing the value
      k = __shfl_sync(0xffffffffu, k, 0) * 0 + ((i + 1) & 63);
    }
  } else if (mode == 1) { // shared memory
    int k = 0;
    for (int i = 0; i < iters; ++i) {
      double v = sm[k * 32 + lane];
      x = fma(v, x, 1e-12);
      k = ((i + 1) & 63) + (__double2loint(x) & 0);
    }
  } else {                // TMEM, 4 loads in flight, one wait
    for (int i = 0; i < iters; i += 4) {
      int k = i & 63;
      double v0 = tmem_ld1(base + 2 * k), v1 = tmem_ld1(base + 2 * (k + 1)), v2 = tmem_ld1(base + 2 * (k + 2)), v3 = tmem_ld1(base + 2 * (k + 3));
      tmem_wait_ld();
      x = fma(v0, x, 1e-12); x = fma(v1, x, 1e-12); x = fma(v2, x, 1e-12); x = fma(v3, x, 1e-12);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; }
  if (x == 123.456) out[0] = x;
  __syncthreads();
  if (warp == 0) tmem_dealloc(sa, 128);
}

// ---- (3) throughput: every warp streams its 64 doubles repeatedly -----------------------------------------
__global__ void k_thr(int iters, double *out, int mode) {
  __shared__ uint32_t sa;
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&sa, 128); tmem_relinquish(); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = sa + ((uint32_t)((warp & 3) * 32) << 16);
  double *my = sm + warp * 64 * 32 + lane;
  for (int k = 0; k < 64; ++k) { tmem_st1(base + 2 * k, 1.0 + 1e-9 * k); my[k * 32] = 1.0 + 1e-9 * k; }
  tmem_wait_st();
  __syncwarp();
  double x0 = 0, x1 = 0, x2 = 0, x3 = 0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      double v[4];
      if (mode == 0) { tmem_ld4(base + 2 * k, v); tmem_wait_ld(); }
      else if (mode == 1) { tmem_ld2(base + 2 * k, v); tmem_ld2(base + 2 * k + 4, v + 2); tmem_wait_ld(); }
      else if (mode == 2) { v[0] = tmem_ld1(base + 2 * k); v[1] = tmem_ld1(base + 2 * k + 2); v[2] = tmem_ld1(base + 2 * k + 4); v[3] = tmem_ld1(base + 2 * k + 6); tmem_wait_ld(); }
      else { v[0] = my[k * 32]; v[1] = my[(k + 1) * 32]; v[2] = my[(k + 2) * 32]; v[3] = my[(k + 3) * 32]; }
      x0 += v[0]; x1 += v[1]; x2 += v[2]; x3 += v[3];
    }
  }
  if (x0 + x1 + x2 + x3 == 123.456) out[0] = x0;
  __syncthreads();
  if (warp == 0) tmem_dealloc(sa, 128);
}

int main() {
  int *d_err; unsigned *d_bases; double *d_out; long long *d_cyc;
  CK(cudaMalloc(&d_err, 4)); CK(cudaMalloc(&d_bases, 512)); CK(cudaMalloc(&d_out, 8)); CK(cudaMalloc(&d_cyc, 16));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs: %d\n", sms);
  struct { int warps, c1, c2, blocks_per_sm; } cases[] = {{4, 128, 0, 1}, {4, 128, 32, 3}, {4, 128, 0, 4}, {8, 256, 0, 2}, {6, 128, 32, 2}};
  for (auto &c : cases) {
    CK(cudaMemset(d_err, 0, 4));
    k_correct<<<sms * c.blocks_per_sm * 4, c.warps * 32>>>(c.c1, c.c2, d_err, d_bases, 20000);
    CK(cudaDeviceSynchronize());
    int err; unsigned bases[8];
    CK(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(bases, d_bases, 32, cudaMemcpyDeviceToHost));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_correct, c.warps * 32, 0));
    printf("correctness: %d warps/block, alloc %d+%d columns, grid %d (occupancy limit %d blocks/SM): %d mismatches; bases of blocks 0..3: %08x/%08x %08x/%08x %08x/%08x %08x/%08x\n",
           c.warps, c.c1, c.c2, sms * c.blocks_per_sm * 4, occ, err, bases[0], bases[1], bases[2], bases[3], bases[4], bases[5], bases[6], bases[7]);
  }
  const char *names[] = {"tcgen05.ld.x2 + wait, dependent", "ld.shared.f64, dependent", "4 x tcgen05.ld.x2 + one wait"};
  for (int mode = 0; mode < 3; ++mode) {
    const int iters = 4096;
    k_lat<<<1, 32>>>(iters, d_out, d_cyc, mode);
    CK(cudaDeviceSynchronize());
    long long cyc; CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    printf("latency  %-34s %.1f cycles per load+fma\n", names[mode], (double)cyc / iters);
  }
  const char *tn[] = {"tcgen05.ld.x8 (4 doubles)", "2 x tcgen05.ld.x4", "4 x tcgen05.ld.x2", "4 x ld.shared.f64"};
  CK(cudaFuncSetAttribute(k_thr, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 64 * 32 * 8));
  for (int wpb : {4, 8}) for (int bps : {1, 2, 3, 4}) {
    if (wpb * bps > 16 || (wpb == 8 && bps > 2)) continue;
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 2000;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      const size_t smem = (size_t)wpb * 64 * 32 * 8;
      k_thr<<<sms * bps, wpb * 32, smem>>>(10, d_out, mode);
      CK(cudaEventRecord(e0));
      k_thr<<<sms * bps, wpb * 32, smem>>>(iters, d_out, mode);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double bytes_per_sm = (double)iters * 64 * 8 * 32 * wpb * bps;
      printf("throughput %d warps/SM (%d x %d)  %-28s %.1f B/clk/SM (at 1.965 GHz)\n", wpb * bps, bps, wpb, tn[mode], bytes_per_sm / (ms * 1e-3) / 1.965e9);
    }
  }
  printf("done\n");
  return 0;
}
