// wt_kernels.cu -- sm_100a kernels and the C ABI (include/wt_b200.h).
//
// K1  wt_step_begin_kernel + wt_step_run_kernel   the plant step (wt_step_core.h): zones -> lanes, PCR solves;
//     wt_catch_up_kernel     the same step, fused, for the deferred plants of a block of steps (exact or floor mode)
//     wt_derivatives_kernel  batched RHS only
// K2  wt_calc_ph_kernel      batched charge-balance Newton-Raphson (chemistry.py:271-330)
// K3  wt_sensors_read_kernel the 7-sensor suite (wt_sensors.cuh) + maintenance / reset / statistics kernels
// K4  wt_stats_* / wt_sensor_stats_*   the payload of the one collective;  K5 diagnostics;  K6 register image
//     wt_cost_* (counting sort of the cost order), wt_apply_commands / wt_scenario_commands (orchestrator),
//     wt_dfma_peak_kernel    FP64 pipe saturation probe for the roofline denominator
//
// Nothing here is GEMM shaped, so there are no tensor-core / TMA paths: the step kernel is
// bound by the FP64 FMA pipe and touches HBM only at entry and exit.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/wt_b200.h"
#include "wt_step_core.h"
#include "wt_sensors.cuh"

static_assert((int)WT_NPAR == (int)WTP_NPAR && (int)WT_NBND == (int)WTB_NBND && (int)WT_NCNT == (int)WTC_NCNT,
              "ABI enums out of sync with the core");
static_assert(WT_ST_HALT_MASK == WTS_HALT_MASK, "status bits out of sync");

static thread_local char g_err[256] = "";
static int set_err(int code, const char *msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int cuda_err(cudaError_t e, const char *where) {
  if (e == cudaSuccess) return 0;
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return (int)e;
}

// ---------------------------------------------------------------------------------------
// per-warp store of the PCR factorizations: real multipliers in shared memory, complex ones in TENSOR MEMORY
// ---------------------------------------------------------------------------------------
// Tensor memory (256 KB per SM, 128 lanes x 512 32-bit columns) is idle in a kernel without tcgen05.mma.  With the
// 32x32b shape, thread i of warp w reads / writes consecutive columns of TMEM lane 32 (w % 4) + i: a lane-private
// array, addressed by a warp-uniform column.  The complex multipliers (12 L doubles per lane, 2/3 of the LU store)
// live there: one LDTM.x8 fetches {k1r, k1i, k2r, k2i} of a level (it is scoreboarded like a shared-memory load,
// ~18 cycles, measured in tools/micro/tmem.cu), and shared memory per warp drops from 23 KB to 9 KB, which is what
// lets a third block onto the SM.  A block allocates 128 columns (24 L <= 120 for n <= 32); 4 blocks fit an SM.
#define WT_TMEM_COLS 128
__device__ __forceinline__ void wt_tmem_ld4(uint32_t a, double *x) {
  int r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
  // the wait is tied to the loaded registers so that no use can be scheduled above it (SASS: LDTM is scoreboarded,
  // the wait itself is a NOP)
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
#pragma unroll
  for (int i = 0; i < 4; ++i) x[i] = __hiloint2double(r[2 * i + 1], r[2 * i]);
}
__device__ __forceinline__ void wt_tmem_st4(uint32_t a, const double *x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(a), "r"(__double2loint(x[0])), "r"(__double2hiint(x[0])), "r"(__double2loint(x[1])), "r"(__double2hiint(x[1])),
                  "r"(__double2loint(x[2])), "r"(__double2hiint(x[2])), "r"(__double2loint(x[3])), "r"(__double2hiint(x[3])));
}

struct SmemLu {
  double *p;   // &lu_region[lane]; slot stride = 32 doubles
  double *cp;  // constants of this lane's plant (one copy per plant, read as a broadcast)
  int *ci;     // solver path counters of this lane's plant (WTC_*): in shared memory, not in 8 registers per lane that
               // ptxas spilled; every lane of the plant adds the same increment to the same word (benign)
  uint32_t tm; // tensor-memory address of column 0 of this warp's lane quarter
  bool rmw;    // this factorization must preserve the stored factors of some plant of the warp
  __device__ __forceinline__ void czero() {
#pragma unroll
    for (int k = 0; k < WTC_NCNT; ++k) ci[k] = 0;
  }
  __device__ __forceinline__ void cadd(int k, int inc) { ci[k] += inc; }
  __device__ __forceinline__ int cval(int k) const { return ci[k]; }
  // per-plant step-control scalars (WtPlantStep::PV_*), after the constants and the counters
  __device__ __forceinline__ double pvget(int k) const { return cp[CK_N + WTC_NCNT / 2 + k]; }
  __device__ __forceinline__ void pvput(int k, double x) { cp[CK_N + WTC_NCNT / 2 + k] = x; }
  // predicated store: an `if (mask)` here becomes a branch around each group of stores, and a branch ends
  // the basic block ptxas schedules in (the six factorizations of a PCR level then run one after the other)
  __device__ __forceinline__ void put(int slot, double x, bool mask) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q st.shared.f64 [%0], %1;\n\t}"
                 :: "r"((unsigned)__cvta_generic_to_shared(p + slot * 32)), "d"(x), "r"((int)mask) : "memory");
  }
  __device__ __forceinline__ double get(int slot) const { return p[slot * 32]; }
  // complex slot space: tensor memory, two columns per double.  tcgen05.st is a warp-wide instruction and cannot be
  // predicated per lane, so a factorization that must leave another plant's factors alone (rmw) merges with the
  // stored values first; the common case (every plant of the warp factorizes, or the others hold nothing) stores.
  __device__ __forceinline__ void begin_factor(bool keep) { rmw = __any_sync(0xffffffffu, keep); }
  __device__ __forceinline__ void end_factor() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
  __device__ __forceinline__ void cx_get4(int slot, double *x) const { wt_tmem_ld4(tm + 2 * slot, x); }
  __device__ __forceinline__ void cx_put4(int slot, const double *x, bool mask) {
    double v[4] = {x[0], x[1], x[2], x[3]};
    if (rmw) {
      double o[4];
      wt_tmem_ld4(tm + 2 * slot, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = mask ? v[i] : o[i];
    }
    wt_tmem_st4(tm + 2 * slot, v);
  }
  __device__ __forceinline__ void cput(int k, double x) { cp[k] = x; }
  __device__ __forceinline__ double cget(int k) const { return cp[k]; }
  __device__ __forceinline__ void csync() { __syncwarp(); }
};

// shared-memory lane slots: the real multipliers (3 systems x 2 L) ...
__host__ __device__ inline int wt_lu_slots(int n) {
  int L = 0;
  for (int s = 1; s < n; s <<= 1) ++L;
  return 3 * 2 * L;
}
// + the lane-private constants (LK_*) and the parked per-lane solver state (WtPlantStep::PK_*) after the LU multipliers
#define WT_PARK_SLOTS 27
__host__ __device__ inline int wt_lane_slots(int n) { return wt_lu_slots(n) + LK_N + WT_PARK_SLOTS; }
// doubles of shared memory per warp: LU slots for 32 lanes + constants for (32/n + 1) plants
#define WT_PLANT_DOUBLES (CK_N + WTC_NCNT / 2 + 10)  // per-plant constants + path counters (ints) + step-control scalars (PV_N)
__host__ __device__ inline int wt_warp_smem_doubles(int n) { return wt_lane_slots(n) * 32 + (32 / n + 1) * WT_PLANT_DOUBLES; }

static_assert(WtPlantStep<SmemLu>::PV_N == 10 && WTC_NCNT % 2 == 0, "per-plant store layout out of sync with WT_PLANT_DOUBLES");
static_assert(WtPlantStep<SmemLu>::PK_N == WT_PARK_SLOTS, "parking slots out of sync with the core");
static_assert(24 * 5 <= WT_TMEM_COLS, "complex LU slots of n <= 32 zones (L <= 5) must fit the tensor-memory allocation");

struct StepArgs {
  int P, n, n_steps, bnd_stride, max_attempts;
  int ld;                // row stride of every SoA array (= P unless a column slab of a larger ensemble is stepped)
  double inv_sqrtN, inv_sqrt3N;  // 1/sqrt(3n), 1/sqrt(9n) of the RMS norms, computed by the host
  double dt;
  const double *par, *bnd;
  double *time, *y, *flow, *derived;
  uint32_t *status;
  int32_t *counters;
  const int32_t *order;  // optional: slot -> plant (e.g. plants sorted by the cost of their last step)
  int32_t *cost;         // optional: per-plant work of this launch (collocation solves + Newton iterations)
  char *ws;              // workspace (wt_step_workspace_bytes): work-queue counter + the begin -> run hand-off rows
  int n_groups;          // warps' worth of plants in this launch (ceil(P / plants per warp))
  uint32_t skip_mask;    // plants with one of these status bits are passed over (WTS_SKIP_MASK; a catch-up launch: halts only)
  const int32_t *count_dev;  // optional: number of valid entries of `order` (a device-side list shorter than P)
  const double *t_stop;  // optional: plants whose time has reached *t_stop are passed over (catch-up launches)
  double h_floor;        // > 0: floor mode (catch-up launches only): step sizes >= h_floor with forced acceptance at the floor
};

#ifndef WT_STEP_WARPS
#define WT_STEP_WARPS 4
#endif
#ifndef WT_STEP_MINBLOCKS
#define WT_STEP_MINBLOCKS 3
#endif
#ifndef WT_STEP_CARVEOUT_PCT
#define WT_STEP_CARVEOUT_PCT 100  // percent of the 228 KB: three blocks of ~63 KB (n = 10) or ~69 KB (n = 20)
#endif
#ifndef WT_BEGIN_MINBLOCKS
#define WT_BEGIN_MINBLOCKS 3
#endif
#define WT_MAX_DEVICES 64

// ---------------------------------------------------------------------------------------
// One step = TWO kernels.  The fused kernel of round 1 (5,000 SASS instructions, 80 KB) was bound by instruction
// fetch: every warp streamed the whole kernel through a ~32 KB SM instruction cache once per step, the GPC-level
// instruction cache ran at 84-95 % of its request rate, and a block kept its registers / shared memory until its
// slowest warp had finished (30 % of the resident warp slots idle).  So:
//   K1a  wt_step_begin_kernel  the once-per-step work whose control flow is the same for every plant (constants,
//        f0, select_initial_step, the first finite-difference Jacobian).  All warps of all blocks walk the same code
//        at the same time, so a fetched line serves every warp of the SM.  Leaves 22 doubles per lane + 2 per plant
//        in the hand-off rows of the workspace (coalesced 256 B rows per warp).
//   K1b  wt_step_run_kernel    the data-dependent attempt loop, as PERSISTENT warps: a warp that has finished its
//        group of plants takes the next group from a work queue (one atomicAdd), so no warp slot waits for a block
//        mate and every resident warp executes the same ~2,900 instructions of loop code.
// ---------------------------------------------------------------------------------------
enum { HO_F = 0, HO_JFAC = 3, HO_JD = 6, HO_JPT = 15, HO_JCT = 18, HO_JCP = 21, HO_SELF_H = 22, HO_FL = 23, HO_ROWS = 24 };
#define WT_WS_HEADER 256  // bytes: [0] the work-queue counter of K1b
__host__ __device__ inline size_t wt_ws_bytes(long long n_groups) { return WT_WS_HEADER + (size_t)n_groups * HO_ROWS * 32 * sizeof(double); }

// shared memory per warp of K1a: no LU slots, only the lane constants, the parking slots and the per-plant rows
__host__ __device__ inline int wt_begin_lane_slots() { return LK_N + WT_PARK_SLOTS; }
__host__ __device__ inline int wt_begin_smem_doubles(int n) { return wt_begin_lane_slots() * 32 + (32 / n + 1) * WT_PLANT_DOUBLES; }

// lane <-> plant mapping of group `wg` (both kernels use the same one)
struct LaneMap { int p, z, gi; bool in_plant; };
__device__ __forceinline__ int wt_plant_count(const StepArgs &a) {
  int Pn = a.P;
  if (a.count_dev) { const int c = *a.count_dev; Pn = c < Pn ? c : Pn; }
  return Pn;
}
// a plant this launch works on: not passed over by its status, not yet at the stop time of a catch-up launch
__device__ __forceinline__ bool wt_plant_live(const StepArgs &a, uint32_t st, double t) {
  return !(st & a.skip_mask) && !(a.t_stop && t >= *a.t_stop - 0.5 * a.dt);
}
__device__ __forceinline__ LaneMap wt_lane_map(const StepArgs &a, long long wg, int lane, int n, int gpw) {
  LaneMap m;
  m.gi = lane / n;
  const long long pl = wg * gpw + m.gi;
  m.in_plant = m.gi < gpw && pl < wt_plant_count(a);
  m.p = m.in_plant ? (a.order ? a.order[pl] : (int)pl) : 0;
  m.z = m.in_plant ? lane - m.gi * n : 0;
  return m;
}

// NZ > 0: the zone count is a compile-time constant (the BASELINE shapes n = 10 and n = 20): lane geometry, PCR level
// count and every LU slot offset fold to immediates after inlining; NZ = 0 reads n from the arguments.
template <int WARPS, int NZ>
__global__ void __launch_bounds__(WARPS * 32, WT_BEGIN_MINBLOCKS) wt_step_begin_kernel(StepArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = NZ > 0 ? NZ : a.n, gpw = 32 / n;
  const long long wg = (long long)blockIdx.x * WARPS + warp;
  if (blockIdx.x == 0 && threadIdx.x == 0) *(int *)a.ws = 0;  // the work queue of the run kernel that follows
  if (wg >= a.n_groups) return;
  const LaneMap lm = wt_lane_map(a, wg, lane, n, gpw);
  const int p = lm.p, z = lm.z;
  const size_t P = (size_t)a.ld;

  // every global load issued back to back before the first use (one round trip to HBM)
  const uint32_t st_in = a.status[p];
  double par[WTP_NPAR], bnd[WTB_NBND];
#pragma unroll
  for (int k = 0; k < WTP_NPAR; ++k) par[k] = a.par[(size_t)k * P + p];
#pragma unroll
  for (int k = 0; k < WTB_NBND; ++k) bnd[k] = a.bnd[a.bnd_stride ? (size_t)k * P + p : (size_t)k];
  const double t = a.time[p];
  double y0[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) y0[v] = a.y[((size_t)v * n + z) * P + p];

  const bool on = lm.in_plant && wt_plant_live(a, st_in, t);
  double *ho = (double *)(a.ws + WT_WS_HEADER) + (size_t)wg * HO_ROWS * 32 + lane;
  if (!__any_sync(0xffffffffu, on)) {
    ho[HO_FL * 32] = __longlong_as_double(0ll);  // nothing running in this group
    return;
  }
  SmemLu lu;
  lu.p = smem + (size_t)warp * wt_begin_smem_doubles(n) + lane;
  lu.cp = smem + (size_t)warp * wt_begin_smem_doubles(n) + wt_begin_lane_slots() * 32 + (lm.gi < gpw ? lm.gi : gpw) * WT_PLANT_DOUBLES;
  lu.ci = (int *)(lu.cp + CK_N);
  lu.tm = 0;
  lu.rmw = false;
  lu.czero();
  WtPlantStep<SmemLu> ps;
  ps.g = wt_make_group(n, a.inv_sqrtN, a.inv_sqrt3N);
  ps.lu = &lu;
  ps.pk0 = LK_N;
  ps.c = wt_make_const(&lu, ps.g, 0, par, bnd);
#pragma unroll
  for (int v = 0; v < 3; ++v) ps.y[v] = y0[v];
  ps.begin(t, a.dt, on);

  // hand-off rows (coalesced: row k of group wg is 32 consecutive doubles)
  typedef WtPlantStep<SmemLu> PS;
#pragma unroll
  for (int v = 0; v < 3; ++v) {
    ho[(HO_F + v) * 32] = ps.pk(PS::PK_F + v);
    ho[(HO_JFAC + v) * 32] = ps.pk(PS::PK_JFAC + v);
    ho[(HO_JPT + v) * 32] = ps.J.pt[v];
    ho[(HO_JCT + v) * 32] = ps.J.ct[v];
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) ho[(HO_JD + k) * 32] = ps.pk(PS::PK_JD + k);
  ho[HO_JCP * 32] = ps.J.cp;
  ho[HO_SELF_H * 32] = ps.pv(PS::PV_SELF_H);
  ho[HO_FL * 32] = __longlong_as_double((long long)ps.fl);
  if (on && z == 0 && a.counters) {
    atomicAdd(&a.counters[(size_t)WTC_NFEV * P + p], lu.cval(WTC_NFEV));
    atomicAdd(&a.counters[(size_t)WTC_NJEV * P + p], lu.cval(WTC_NJEV));
    if (lu.cval(WTC_JAC_RETRY)) atomicAdd(&a.counters[(size_t)WTC_JAC_RETRY * P + p], lu.cval(WTC_JAC_RETRY));
  }
}

// FLOOR: the floor-mode variant of the attempt loop (WtPlantStep::run<true>), launched by wt_catch_up only.
template <int WARPS, int NZ, bool FLOOR>
__global__ void __launch_bounds__(WARPS * 32, WT_STEP_MINBLOCKS) wt_step_run_kernel(StepArgs a) {
  extern __shared__ double smem[];
  static_assert(WARPS == 4, "one warp per tensor-memory lane quarter");
  __shared__ uint32_t tmem_slot;
  // (the warp index through a warp reduction: a value ptxas knows to be warp-uniform, see the queue below)
  const int lane = threadIdx.x & 31, warp = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
  // tensor memory for the complex LU multipliers: warp 0 allocates the block's columns (tcgen05.alloc writes the
  // address to shared memory) and gives up the allocation permit so that the other resident blocks can allocate
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "n"(WT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  const int n = NZ > 0 ? NZ : a.n, gpw = 32 / n;
  const size_t P = (size_t)a.ld;  // row stride; a.P bounds the plant index
  typedef WtPlantStep<SmemLu> PS;
  SmemLu lu;
  lu.p = smem + (size_t)warp * wt_warp_smem_doubles(n) + lane;
  lu.ci = nullptr;
  lu.cp = nullptr;
  lu.tm = tmem_base + ((uint32_t)(warp * 32) << 16);
  lu.rmw = false;
  PS ps;
  ps.g = wt_make_group(n, a.inv_sqrtN, a.inv_sqrt3N);
  ps.lu = &lu;
  ps.pk0 = wt_lu_slots(n) + LK_N;
  int *queue = (int *)a.ws;

  // persistent warps: the first group by position, the following ones from the queue
  long long wg = (long long)blockIdx.x * WARPS + warp;
  const long long first_wave = (long long)gridDim.x * WARPS;
#pragma unroll 1
  while (wg < a.n_groups) {
    // The optional-argument tests below are loop invariant; left visible, the compiler "unswitches" the loop into
    // one full copy of the 5,000-instruction body per combination.  Laundering them through an empty asm keeps ONE copy.
    const int32_t *order = a.order;
    int32_t *counters = a.counters, *cost = a.cost;
    double *derived = a.derived, *flow = a.flow;
    int bnd_stride = a.bnd_stride, or_status = a.n_steps;
    asm volatile("" : "+l"(order), "+l"(counters), "+l"(cost), "+l"(derived), "+l"(flow), "+r"(bnd_stride), "+r"(or_status));
    LaneMap lm;
    {
      lm.gi = lane / n;
      const long long pl = wg * gpw + lm.gi;
      lm.in_plant = lm.gi < gpw && pl < wt_plant_count(a);
      lm.p = lm.in_plant ? (order ? order[pl] : (int)pl) : 0;
      lm.z = lm.in_plant ? lane - lm.gi * n : 0;
    }
    const int p = lm.p, z = lm.z;
    const double *ho = (const double *)(a.ws + WT_WS_HEADER) + (size_t)wg * HO_ROWS * 32 + lane;
    const int fl_in = (int)__double_as_longlong(ho[HO_FL * 32]);
    const uint32_t st_in = a.status[p];
    double t = a.time[p];
    const bool live = lm.in_plant && wt_plant_live(a, st_in, t);
    if (__any_sync(0xffffffffu, live)) {
      // all loads of the group issued back to back
      double par[WTP_NPAR], bnd[WTB_NBND], hrow[HO_SELF_H + 1];
#pragma unroll
      for (int k = 0; k < WTP_NPAR; ++k) par[k] = a.par[(size_t)k * P + p];
#pragma unroll
      for (int k = 0; k < WTB_NBND; ++k) bnd[k] = a.bnd[bnd_stride ? (size_t)k * P + p : (size_t)k];
      double y0[3];
#pragma unroll
      for (int v = 0; v < 3; ++v) y0[v] = a.y[((size_t)v * n + z) * P + p];
#pragma unroll
      for (int k = 0; k <= HO_SELF_H; ++k) hrow[k] = ho[k * 32];

      lu.cp = smem + (size_t)warp * wt_warp_smem_doubles(n) + wt_lane_slots(n) * 32 + (lm.gi < gpw ? lm.gi : gpw) * WT_PLANT_DOUBLES;
      lu.ci = (int *)(lu.cp + CK_N);
      __syncwarp();  // the previous group's per-plant rows are dead
      lu.czero();
      // reactor.py:500: flow_rate = inlet + acid + chlorine flow, parked with the plant's constants until the epilogue
      lu.cput(CK_flow, bnd[WTB_INLET_FLOW] + bnd[WTB_ACID_FLOW] + bnd[WTB_CL_FLOW]);
      ps.c = wt_make_const(&lu, ps.g, wt_lu_slots(n), par, bnd);
#pragma unroll
      for (int v = 0; v < 3; ++v) ps.y[v] = y0[v];
      // solver state as begin() left it
      ps.reset(t, a.dt, live);
      ps.fl = fl_in;
      ps.pvset(PS::PV_SELF_H, hrow[HO_SELF_H]);
      {
        const bool all = true;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          ps.pkset(PS::PK_F + v, hrow[HO_F + v], all);
          ps.pkset(PS::PK_JFAC + v, hrow[HO_JFAC + v], all);
          ps.J.pt[v] = hrow[HO_JPT + v];
          ps.J.ct[v] = hrow[HO_JCT + v];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) ps.pkset(PS::PK_JD + k, hrow[HO_JD + k], all);
        ps.J.cp = hrow[HO_JCP];
      }
      double yin[3] = {ps.y[0], ps.y[1], ps.y[2]};
      ps.template run<FLOOR>(t, a.max_attempts, a.h_floor);
      double der[3];
      bool adv;
      const int sb = wt_finish_step(ps, yin, der, adv);
      if (live) {
#pragma unroll
        for (int v = 0; v < 3; ++v) a.y[((size_t)v * n + z) * P + p] = ps.y[v];
        if (derived && adv) {
#pragma unroll
          for (int v = 0; v < 3; ++v) derived[((size_t)v * n + z) * P + p] = der[v];
        }
        if (z == 0) {
          if (adv) a.time[p] = t + a.dt;
          // (the deferred mark survives: only wt_defer_rejoin takes it off)
          a.status[p] = (uint32_t)sb | (or_status > 0 ? (st_in & ~(uint32_t)WTS_SKIP_MASK) : 0u) | (st_in & (uint32_t)WTS_DEFERRED);
          if (adv && flow) flow[p] = lu.cget(CK_flow);
          if (counters) {
            // fire-and-forget reductions: a load-add-store here made the whole warp wait for eight loads
#pragma unroll
            for (int k = 0; k < WTC_NCNT; ++k)
              if (lu.cval(k)) atomicAdd(&counters[(size_t)k * P + p], lu.cval(k));
          }
          if (cost) cost[p] = lu.cval(WTC_NSTEPS) + lu.cval(WTC_NREJECT) + lu.cval(WTC_NNEWTON_FAIL) + lu.cval(WTC_NNEWTON);
        }
      }
    }
    // next group of plants: one atomic per warp
    // (the warp reduction returns its result in a uniform register: ptxas then KNOWS the loop is warp-uniform and
    // keeps the body free of WARPSYNC / divergence handling, which otherwise doubles its size)
    unsigned nx = 0;
    if (lane == 0) nx = (unsigned)atomicAdd(queue, 1);
    wg = first_wave + (long long)__reduce_max_sync(0xffffffffu, nx);
  }
  // every warp is done with its tensor-memory columns: the allocating warp frees them
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(WT_TMEM_COLS) : "memory");
}

// K1c  wt_catch_up_kernel  the deferred plants of one block of steps, ALL their steps in ONE launch.
// A warp takes one group of listed plants and runs begin() + run() step after step until every plant of the group has
// reached *t_stop (or halted).  The list is short (1e-4 .. 1e-2 of the ensemble), so what matters here is not the
// instruction cache (this kernel is the fused 7,000-instruction form the step kernels were split out of) but the number
// of launches: 2 x steps-per-block launch pairs of the two step kernels on the side stream cost more than the plants in
// them (each had to wait for a gap between the persistent main launches).  One launch per block of steps instead.
template <int WARPS, int NZ, bool FLOOR>
__global__ void __launch_bounds__(WARPS * 32, WT_STEP_MINBLOCKS) wt_catch_up_kernel(StepArgs a) {
  extern __shared__ double smem[];
  static_assert(WARPS == 4, "one warp per tensor-memory lane quarter");
  __shared__ uint32_t tmem_slot;
  const int lane = threadIdx.x & 31, warp = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
  const int n = NZ > 0 ? NZ : a.n, gpw = 32 / n;
  const long long groups = ((long long)wt_plant_count(a) + gpw - 1) / gpw;
  if ((long long)blockIdx.x * WARPS >= groups) return;  // (block-uniform) nothing listed for this block
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "n"(WT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  const size_t P = (size_t)a.ld;
  typedef WtPlantStep<SmemLu> PS;
  const long long wg = (long long)blockIdx.x * WARPS + warp;
  if (wg < groups) {
    const LaneMap lm = wt_lane_map(a, wg, lane, n, gpw);
    const int p = lm.p, z = lm.z;
    const uint32_t st_in = a.status[p];
    double t = a.time[p];
    double par[WTP_NPAR], bnd[WTB_NBND], y[3], der[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < WTP_NPAR; ++k) par[k] = a.par[(size_t)k * P + p];
#pragma unroll
    for (int k = 0; k < WTB_NBND; ++k) bnd[k] = a.bnd[a.bnd_stride ? (size_t)k * P + p : (size_t)k];
#pragma unroll
    for (int v = 0; v < 3; ++v) y[v] = a.y[((size_t)v * n + z) * P + p];
    bool on = lm.in_plant && !(st_in & a.skip_mask);
    SmemLu lu;
    lu.p = smem + (size_t)warp * wt_warp_smem_doubles(n) + lane;
    lu.cp = smem + (size_t)warp * wt_warp_smem_doubles(n) + wt_lane_slots(n) * 32 + (lm.gi < gpw ? lm.gi : gpw) * WT_PLANT_DOUBLES;
    lu.ci = (int *)(lu.cp + CK_N);
    lu.tm = tmem_base + ((uint32_t)(warp * 32) << 16);
    lu.rmw = false;
    PS ps;
    ps.g = wt_make_group(n, a.inv_sqrtN, a.inv_sqrt3N);
    ps.lu = &lu;
    ps.pk0 = wt_lu_slots(n) + LK_N;
    ps.c = wt_make_const(&lu, ps.g, wt_lu_slots(n), par, bnd);
    const double flow = bnd[WTB_INLET_FLOW] + bnd[WTB_ACID_FLOW] + bnd[WTB_CL_FLOW];  // reactor.py:500
    const double t_stop = a.t_stop ? *a.t_stop : 1e300;
    uint32_t st_acc = st_in;
    bool advanced = false;
#pragma unroll 1
    for (int s = 0; s < a.n_steps; ++s) {
      on = on && !(t >= t_stop - 0.5 * a.dt);
      if (!__any_sync(0xffffffffu, on)) break;
      __syncwarp();
      lu.czero();
      const double yin[3] = {y[0], y[1], y[2]};
#pragma unroll
      for (int v = 0; v < 3; ++v) ps.y[v] = y[v];
      ps.begin(t, a.dt, on);
      ps.template run<FLOOR>(t, a.max_attempts, a.h_floor);
      double d[3];
      bool adv;
      const int sb = wt_finish_step(ps, yin, d, adv);
      if (on) {
#pragma unroll
        for (int v = 0; v < 3; ++v) y[v] = ps.y[v];
        if (adv) {
          t += a.dt;
          advanced = true;
#pragma unroll
          for (int v = 0; v < 3; ++v) der[v] = d[v];
        }
        // the non-halting bits of the earlier steps of this launch stay (as wt_advance ORs them); the deferred mark survives
        st_acc = (uint32_t)sb | (s > 0 ? (st_acc & ~(uint32_t)WTS_SKIP_MASK) : 0u) | (st_in & (uint32_t)WTS_DEFERRED);
        if (z == 0 && a.counters) {
#pragma unroll
          for (int k = 0; k < WTC_NCNT; ++k)
            if (lu.cval(k)) atomicAdd(&a.counters[(size_t)k * P + p], lu.cval(k));
        }
        if (sb & (int)WTS_HALT_MASK) on = false;
      }
    }
    if (lm.in_plant && !(st_in & a.skip_mask) && (advanced || st_acc != st_in)) {
#pragma unroll
      for (int v = 0; v < 3; ++v) a.y[((size_t)v * n + z) * P + p] = y[v];
      if (a.derived && advanced) {
#pragma unroll
        for (int v = 0; v < 3; ++v) a.derived[((size_t)v * n + z) * P + p] = der[v];
      }
      if (z == 0) {
        a.time[p] = t;
        a.status[p] = st_acc;
        if (advanced && a.flow) a.flow[p] = flow;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(WT_TMEM_COLS) : "memory");
}

__global__ void wt_derivatives_kernel(int P, int n, const double *par_, const double *bnd_, int bnd_stride,
                                      const double *y, double *dy, int32_t *bad_out) {
  const int lane = threadIdx.x & 31;
  const int gpw = 32 / n;
  const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int gi = lane / n;
  const long long pl = wg * gpw + gi;
  const bool in_plant = gi < gpw && pl < P;
  const int p = in_plant ? (int)pl : 0;
  const int z = in_plant ? lane - gi * n : 0;
  WtGroup g = wt_make_group(n);
  __shared__ double cs[4][17 * CK_N + LK_N * 32];
  SmemLu st;
  st.p = &cs[threadIdx.x >> 5][17 * CK_N + lane];
  st.cp = &cs[threadIdx.x >> 5][(gi < gpw ? gi : gpw) * CK_N];
  st.ci = nullptr;  // the RHS alone counts nothing
  st.tm = 0;
  st.rmw = false;
  double par[WTP_NPAR], bnd[WTB_NBND];
#pragma unroll
  for (int k = 0; k < WTP_NPAR; ++k) par[k] = par_[(size_t)k * P + p];
#pragma unroll
  for (int k = 0; k < WTB_NBND; ++k) bnd[k] = bnd_[bnd_stride ? (size_t)k * P + p : (size_t)k];
  WtConstT<SmemLu> c = wt_make_const(&st, g, 0, par, bnd);
  double yy[3], d[3];
#pragma unroll
  for (int v = 0; v < 3; ++v) yy[v] = y[((size_t)v * n + z) * P + p];
  bool bad;
  wt_rhs(g, c, yy[0], yy[1], yy[2], d[0], d[1], d[2], bad);
  bool gbad = wt_gany(g, bad);
  if (in_plant) {
#pragma unroll
    for (int v = 0; v < 3; ++v) dy[((size_t)v * n + z) * P + p] = d[v];
    if (z == 0 && bad_out) bad_out[p] = gbad ? 1 : 0;
  }
}

__device__ int wt_ph_queues[64];
// chemistry.py:193-330.  The Newton-Raphson iteration counts of independent buffer systems range from 1 to 100
// (BASELINE configs[3]: mode 6-14, 15 % run into the 100-iteration limit and hold half of all iterations), so one thread
// per system leaves 70 % of the lanes of a warp idle (ncu, round 2: 9.6 of 32 threads active per instruction, FP64 pipe
// 61 % busy on them).  Here the warps are persistent and every lane that finishes a solve takes the next system from a
// global queue (one warp-aggregated atomicAdd per refill): lanes stay busy until the queue is empty and the tail is one
// solve.  The arithmetic of a solve is unchanged.
__global__ void __launch_bounds__(128) wt_calc_ph_kernel(int P, const double *alk, const double *ct, const double *temp,
                                                           const double *guess, double *ph, int32_t *iters, int32_t *status,
                                                           int *queue) {
  const int lane = threadIdx.x & 31;
  bool busy = false, drained = false;
  int i = 0, it = 0;
  double pH = 0.0, Kw = 0.0, Ka1 = 0.0, Ka2 = 0.0, C_T = 0.0, alk_eq = 0.0;
  for (;;) {
    // ---- idle lanes take the next systems of the queue.  The set-up of a solve (three transcendentals) is as long as
    // two or three iterations, so it only runs when at least half of the lanes are idle (or none is busy).
    const unsigned need = __ballot_sync(0xffffffffu, !busy);
    const bool refill = !drained && (__popc(need) >= 16 || need == 0xffffffffu);
    if (refill) {
      int base = 0;
      if (lane == 0) base = atomicAdd(queue, __popc(need));
      base = __shfl_sync(0xffffffffu, base, 0);
      drained = base + __popc(need) >= P;
      if (!busy) {
        const int idx = base + __popc(need & ((1u << lane) - 1u));
        if (idx < P) {
          const double tc = temp[idx];
          if (tc < 0.0 || tc > 100.0) {  // thermodynamics.py:146-157 via chemistry.py:118
            ph[idx] = nan("");
            iters[idx] = 0;
            status[idx] = 3;
          } else {
            const double TK = tc + 273.15;
            Kw = 1.0e-14 * exp((55900.0 / 8.314) * (1.0 / 298.15 - 1.0 / TK));
            Ka1 = exp10(-(6.35 + (-0.008) * (tc - 25.0)));
            Ka2 = exp10(-(10.33 + (-0.008) * (tc - 25.0)));
            C_T = ct[idx] / 1000.0;
            alk_eq = alk[idx] / 50000.0;
            pH = guess[idx];
            i = idx;
            it = 0;
            busy = true;
          }
        }
      }
    }
    if (!__any_sync(0xffffffffu, busy)) {
      if (drained) break;
      continue;
    }
    // ---- one Newton-Raphson iteration of every busy lane (chemistry.py:291-330)
    if (busy) {
      const double H = exp10(-pH);
      const double OH = Kw / H;
      const double D = H * H + Ka1 * H + Ka1 * Ka2;
      const double a1 = (Ka1 * H) / D;
      const double a2 = (Ka1 * Ka2) / D;
      const double f = H - OH + a1 * C_T + 2.0 * (a2 * C_T) - alk_eq;
      const double dH = -WT_LN10 * H;
      const double dOH = -(Kw / (H * H)) * dH;
      const double dD = 2.0 * H + Ka1;
      const double da1 = Ka1 * (D - H * dD) / (D * D);
      const double da2 = -Ka1 * Ka2 * dD / (D * D);
      const double df = dH - dOH + C_T * da1 * dH + 2.0 * (C_T * da2 * dH);
      int st = -1;
      if (fabs(df) < 1e-15) st = 1;
      else {
        const double delta = -f / df;
        double pn = pH + delta;
        pn = pn != pn ? pn : fmin(fmax(pn, 0.0), 14.0);  // np.clip keeps NaN
        pH = pn;
        if (fabs(delta) < 1e-6) st = 0;
      }
      ++it;
      if (st < 0 && it >= 100) st = 2;
      if (st >= 0) {
        ph[i] = pH;
        iters[i] = it;
        status[i] = st;
        busy = false;
      }
    }
  }
}


// ---------------------------------------------------------------------------------------
// K4: ensemble statistics (the payload of the one collective).
// Stage 1: each block reduces a slab of plants to per-block partials (coalesced rows of the
// zone-major state).  Stage 2: one block sums the partials in a fixed order -> deterministic.
//   out[0] live plants, out[1] halted plants, out[2] outlet Cl < thr[0],
//   out[3] outlet pH outside [thr[1], thr[2]], out[4] outlet T > thr[3], out[5] live plants whose last step was a
//   floor-mode continuation (WTS_DEGRADED), out[6] of the "halted": plants being caught up (WTS_DEFERRED) or waiting
//   for it (WTS_WORK_LIMIT), out[7] reserved,
//   out[8 + 2*(v*n+z)] = sum (x - shift[v]), out[9 + 2*(v*n+z)] = sum (x - shift[v])^2   (live plants)
// ---------------------------------------------------------------------------------------
#define WT_STATS_HDR 8
#define WT_STATS_TPB 256

__device__ __forceinline__ double block_sum(double v, double *sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WT_STATS_TPB / 32; ++i) r += sh[i];
  }
  __syncthreads();
  return r;  // valid on thread 0
}

__global__ void __launch_bounds__(WT_STATS_TPB) wt_stats_partial_kernel(int P, int n, const double *y, const uint32_t *status,
                                                                          const double *shift_thr, double *partial) {
  __shared__ double sh[WT_STATS_TPB / 32];
  const int nstat = WT_STATS_HDR + 6 * n;
  double *out = partial + (size_t)blockIdx.x * nstat;
  const int per_block = (P + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per_block, hi = min(P, lo + per_block);
  const double c0 = shift_thr[0], c1 = shift_thr[1], c2 = shift_thr[2];
  const double t_cl = shift_thr[3], t_ph_lo = shift_thr[4], t_ph_hi = shift_thr[5], t_T = shift_thr[6];
  double live = 0, halted = 0, e0 = 0, e1 = 0, e2 = 0, degraded = 0, pending = 0;
  for (int p = lo + threadIdx.x; p < hi; p += WT_STATS_TPB) {
    const uint32_t sw = status[p];
    const bool h = (sw & WTS_SKIP_MASK) != 0;   // halted, or deferred (being caught up on the side stream)
    if (h) { halted += 1.0; pending += (sw & (WTS_DEFERRED | WTS_WORK_LIMIT)) ? 1.0 : 0.0; continue; }
    live += 1.0;
    degraded += (sw & WTS_DEGRADED) ? 1.0 : 0.0;
    const double ph = y[((size_t)0 * n + (n - 1)) * P + p], cl = y[((size_t)1 * n + (n - 1)) * P + p],
                 T = y[((size_t)2 * n + (n - 1)) * P + p];
    e0 += cl < t_cl ? 1.0 : 0.0;
    e1 += (ph < t_ph_lo || ph > t_ph_hi) ? 1.0 : 0.0;
    e2 += T > t_T ? 1.0 : 0.0;
  }
  double r;
  r = block_sum(live, sh); if (threadIdx.x == 0) out[0] = r;
  r = block_sum(halted, sh); if (threadIdx.x == 0) out[1] = r;
  r = block_sum(e0, sh); if (threadIdx.x == 0) out[2] = r;
  r = block_sum(e1, sh); if (threadIdx.x == 0) out[3] = r;
  r = block_sum(e2, sh); if (threadIdx.x == 0) { out[4] = r; out[7] = 0; }
  r = block_sum(degraded, sh); if (threadIdx.x == 0) out[5] = r;
  r = block_sum(pending, sh); if (threadIdx.x == 0) out[6] = r;
  for (int row = 0; row < 3 * n; ++row) {
    const double c = row < n ? c0 : (row < 2 * n ? c1 : c2);
    double s1 = 0, s2 = 0;
    for (int p = lo + threadIdx.x; p < hi; p += WT_STATS_TPB) {
      if (status[p] & WTS_SKIP_MASK) continue;
      const double d = y[(size_t)row * P + p] - c;
      s1 += d;
      s2 += d * d;
    }
    r = block_sum(s1, sh); if (threadIdx.x == 0) out[WT_STATS_HDR + 2 * row] = r;
    r = block_sum(s2, sh); if (threadIdx.x == 0) out[WT_STATS_HDR + 2 * row + 1] = r;
  }
}

__global__ void wt_stats_final_kernel(int nblocks, int nstat, const double *partial, double *out, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nstat) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nstat + k];
  out[k] = accumulate ? out[k] + s : s;
}

// ---------------------------------------------------------------------------------------
// K5: per-plant diagnostics: one thread per plant walks its zone rows
// (coalesced across the plants of a warp).  HBM-bound: reads 3n (+n) doubles, writes WT_NDIAG (+ n-1).
// Two-pass mean / population std as numpy's mean / std (reactor.py:570-611, transport.py:338-384,
// spatial.py:322-379, 440-477).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) wt_diagnostics_kernel(int P, int n, const double *par, const double *y,
                                                             const double *hc, double *out, double *n2, int32_t *bad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  const double zh = par[(size_t)WTP_ZH * Pz + p], volume = par[(size_t)WTP_VOLUME * Pz + p];
  const bool strat = par[(size_t)WTP_STRAT * Pz + p] != 0.0;
  const double zone_volume = volume / n, inv_n = 1.0 / n;
  auto Y = [&](int v, int z) { return y[((size_t)v * n + z) * Pz + p]; };
  auto O = [&](int k, double x) { out[(size_t)k * Pz + p] = x; };
  // validate_conservation
  const double T0 = Y(2, 0);
  if (bad) bad[p] = (T0 < 0.0 || T0 > 100.0) ? 1 : 0;
  const double Kw = 1.0e-14 * exp((55900.0 / 8.314) * (1.0 / 298.15 - 1.0 / (T0 + 273.15)));
  double sH = 0.0, sOH = 0.0;
  for (int z = 0; z < n; ++z) {
    const double H = hc ? hc[(size_t)z * Pz + p] : exp10(-Y(0, z));
    sH += H;
    sOH += Kw / H;
  }
  const double tH = sH * zone_volume / 1000, tOH = sOH * zone_volume / 1000;
  O(WT_DG_TOTAL_H_MOL, tH);
  O(WT_DG_TOTAL_OH_MOL, tOH);
  O(WT_DG_CHARGE_BALANCE_MOL, tH - tOH);
  // spatial statistics of the three variables
  for (int v = 0; v < 3; ++v) {
    double s = 0.0, mx = -INFINITY, mn = INFINITY, gmax = -1.0, gsum = 0.0, prev = 0.0, sdev = 0.0;
    int gloc = 0;
    for (int z = 0; z < n; ++z) {
      const double x = Y(v, z);
      s += x;
      mx = fmax(mx, x);  // np.max / np.min propagate NaN; a NaN state is reported by the step's status word
      mn = fmin(mn, x);
      if (z > 0) {
        const double g = fabs((x - prev) / zh);
        gsum += g;
        if (g > gmax) { gmax = g; gloc = z - 1; }  // np.argmax: first maximum
      }
      prev = x;
    }
    const double mean = s * inv_n;
    double thermal = 0.0;
    for (int z = 0; z < n; ++z) {
      const double x = Y(v, z), d = x - mean;
      sdev += d * d;
      thermal += x - 20.0;
    }
    const double sd = sqrt(sdev * inv_n);
    const int k = WT_DG_GRAD0 + 8 * v;
    O(k + 0, mean); O(k + 1, sd); O(k + 2, mx); O(k + 3, mn); O(k + 4, mx - mn);
    O(k + 5, gmax); O(k + 6, gsum / (n - 1)); O(k + 7, (double)gloc);
    if (v == 1) {  // chlorine: total mass + calculate_mixing_quality
      O(WT_DG_TOTAL_CL_MG, s * zone_volume);
      O(WT_DG_CL_CV, mean > 0.0 ? sd / mean : 0.0);
      const double vs = mean * mean;
      O(WT_DG_CL_SEGREGATION, vs > 0.0 ? fmin(fmax((sd * sd) / vs, 0.0), 1.0) : 0.0);
    }
    if (v == 2) {  // temperature: thermal energy relative to 20 C, thermocline
      O(WT_DG_THERMAL_ENERGY_KJ, 998.2 * 4184 * (volume / 1000) * (thermal * inv_n) / 1000);
      O(WT_DG_THERMOCLINE_DEPTH, (strat && gmax > 0.5) ? zh * n - (gloc + 0.5) * zh : nan(""));
    }
  }
  // Brunt-Vaisala N^2 per interface
  double n2max = -INFINITY, n2min = INFINITY, rho_prev = 0.0;
  for (int z = 0; z < n; ++z) {
    const double T = Y(2, z);
    const double d4 = T - 4.0;
    const double rho = T <= 8.0 ? 999.97 - 0.008 * (d4 * d4) : 998.2 - 2.1e-4 * 998.2 * (T - 20.0);
    if (z > 0) {
      const double v = -(9.81 / (0.5 * (rho_prev + rho))) * ((rho - rho_prev) / zh);
      if (n2) n2[(size_t)(z - 1) * Pz + p] = v;
      n2max = fmax(n2max, v);
      n2min = fmin(n2min, v);
    }
    rho_prev = rho;
  }
  O(WT_DG_N2_MAX, n2max);
  O(WT_DG_N2_MIN, n2min);
}

// ---------------------------------------------------------------------------------------
// K6: Modbus input-register image of selected plants (__main__.py:166-224,
// modbus/protocols.py:34-58).  One thread per selected plant; byte work, bound by the 7 gathered reads.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void wt_put_f32(uint16_t *row, int addr, double v) {
  const uint32_t b = __float_as_uint(__double2float_rn(v));  // struct.pack(">f", v): round to nearest even
  row[addr] = (uint16_t)(b >> 16);                           // high word first (big-endian word order)
  row[addr + 1] = (uint16_t)(b & 0xffffu);
}
__global__ void wt_register_image_kernel(int K, const int32_t *sel, int P, const double *value, const int32_t *fault,
                                         double sim_time, uint16_t *ir, uint8_t *di, uint8_t *ok) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int p = sel[k];
  uint16_t *row = ir + (size_t)k * WT_WIRE_NIR;
  for (int a = 0; a < WT_WIRE_NIR; ++a) row[a] = 0;
  uint8_t *bits = di + (size_t)k * WT_WIRE_NDI;
  bits[0] = bits[1] = bits[2] = 0;
  const bool in = p >= 0 && p < P;
  double v[WT_NSENS];
  int f[WT_NSENS];
  bool good = in && sim_time >= -1e9 && sim_time <= 1e9;
  for (int s = 0; s < WT_NSENS; ++s) {
    double x = in ? value[(size_t)s * P + p] : 0.0;
    if (x != x || x == INFINITY || x == -INFINITY) x = 0.0;  // safe_value (__main__.py:178-186)
    v[s] = x;
    f[s] = in ? fault[(size_t)s * P + p] : 0;
    good = good && x >= -1e9 && x <= 1e9;                     // slave.py:146-148
  }
  if (ok) ok[k] = good ? 1 : 0;
  if (!good) return;
  const int addr[WT_NSENS] = {0, 4, 6, 8, 10, 12, 14};       // register_map.py:119-215
  bool any = false;
  for (int s = 0; s < WT_NSENS; ++s) {
    wt_put_f32(row, addr[s], v[s]);
    any = any || f[s] != 0;
  }
  wt_put_f32(row, 100, sim_time);
  row[102] = any ? 1 : 0;
  bits[0] = f[0] != 0;
  bits[1] = f[1] != 0;
  bits[2] = (f[2] != 0) || (f[3] != 0);
}

// BaseSensor.get_statistics over a window of history rows (base_sensor.py:809-856): one thread per plant,
// two-pass mean / population std over the finite values like numpy.
__global__ void wt_sensor_window_stats_kernel(int P, int m, const double *hist, const int32_t *rows, int sensor, double *out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  double s = 0.0, mn = INFINITY, mx = -INFINITY;
  int nf = 0;
  for (int i = 0; i < m; ++i) {
    const double v = hist[((size_t)rows[i] * WT_NSENS + sensor) * Pz + p];
    if (isfinite(v)) { s += v; mn = fmin(mn, v); mx = fmax(mx, v); ++nf; }
  }
  double mean = nan(""), sd = nan("");
  if (nf > 0) {
    mean = s / nf;
    double q = 0.0;
    for (int i = 0; i < m; ++i) {
      const double v = hist[((size_t)rows[i] * WT_NSENS + sensor) * Pz + p];
      if (isfinite(v)) q += (v - mean) * (v - mean);
    }
    sd = sqrt(q / nf);
  }
  const bool none = m == 0;
  out[0 * Pz + p] = none ? 0.0 : mean;
  out[1 * Pz + p] = none ? 0.0 : sd;
  out[2 * Pz + p] = none ? 0.0 : (nf > 0 ? mn : nan(""));
  out[3 * Pz + p] = none ? 0.0 : (nf > 0 ? mx : nan(""));
  out[4 * Pz + p] = (double)m;
  out[5 * Pz + p] = 0.0;
  out[6 * Pz + p] = none ? 0.0 : (nf > 0 ? (double)(m - nf) / m : 1.0);
}

// ---------------------------------------------------------------------------------------
// Deferral of budget-exhausted plants.  The reference never drops a plant for work (solve_ivp runs to the end,
// reactor.py:476-490); the engine budgets the collocation solves of a plant-step.  Instead of
// halting for good, plants that ran out of budget are COLLECTED into a device-side list, marked WTS_DEFERRED (ordinary
// launches pass over them), caught up by launches over that list with a larger budget on a side stream while the
// ensemble moves on, and REJOINED at the next block boundary.  Only a plant that exhausts the larger budget too (or
// raises in the reference's sense, WTS_T_RANGE) stays halted.
// ---------------------------------------------------------------------------------------
__global__ void wt_defer_collect_kernel(int P, uint32_t *status, int32_t *list, int32_t *count, int cap, double *t_stop,
                                        double t_stop_inc) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0 && t_stop) *t_stop += t_stop_inc;   // the end time of the block the catch-up that follows works towards
  if (p >= P) return;
  const uint32_t st = status[p];
  if ((st & WTS_WORK_LIMIT) && !(st & WTS_DEFERRED)) {
    const int slot = atomicAdd(count, 1);
    if (slot < cap) {   // (a full list leaves the plant halted)
      list[slot] = p;
      status[p] = (st & ~(uint32_t)WTS_WORK_LIMIT) | (uint32_t)WTS_DEFERRED;
    }
  }
}
__global__ void wt_defer_rejoin_kernel(uint32_t *status, const int32_t *list, int32_t *count, int cap) {
  int c = *count;
  c = c < cap ? c : cap;
  for (int i = threadIdx.x; i < c; i += blockDim.x) status[list[i]] &= ~(uint32_t)WTS_DEFERRED;
  __syncthreads();
  if (threadIdx.x == 0) *count = 0;
}

// ---------------------------------------------------------------------------------------
// Orchestrator: the reference's main loop turns three actuator commands per plant into
// boundary conditions through two layers of zero-trust clamps (__main__.py:57-63 validate_flow_rate, :227-252
// read_modbus_commands, :255-271 apply_boundary_conditions).  One thread per plant, in place on the boundary SoA.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double wt_validate_flow_rate(double v, double vmax) {  // __main__.py:57-63
  if (v != v) return 0.0;                 // NaN
  return fmax(0.0, fmin(v, vmax));        // max(0.0, min(float(value), max_value)): +inf -> max, -inf -> 0
}
__device__ __forceinline__ void wt_apply_commands_one(double acid, double chlor, double inlet, double *bnd, size_t P, int p) {
  // read_modbus_commands' clamps, then apply_boundary_conditions' (defence in depth)
  acid = wt_validate_flow_rate(wt_validate_flow_rate(acid, 2.0), 2.0);
  chlor = wt_validate_flow_rate(wt_validate_flow_rate(chlor, 1.0), 1.0);
  inlet = wt_validate_flow_rate(inlet, 20.0);
  bnd[(size_t)WTB_ACID_FLOW * P + p] = acid;
  bnd[(size_t)WTB_CL_FLOW * P + p] = chlor;
  if (inlet > 0.1) bnd[(size_t)WTB_INLET_FLOW * P + p] = wt_validate_flow_rate(inlet, 20.0);   // only a significant command
}
__global__ void wt_apply_commands_kernel(int P, const double *acid, const double *chlor, const double *inlet, double *bnd) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  wt_apply_commands_one(acid[p], chlor[p], inlet[p], bnd, (size_t)P, p);
}
// Scenario scripting: S scripts of K piecewise-constant command triplets (breakpoints times[K], ascending), every plant
// follows script sid[p].  The time comes from the device clock of the sensor suite / step loop (clock[0]), so a whole
// scripted run replays from a CUDA graph with no per-step host-to-device traffic.  Before the first breakpoint the
// boundary is left as it is.
__global__ void wt_scenario_commands_kernel(int P, int K, int S, const double *times, const double *cmd, const int32_t *sid,
                                            const double *clock, double t_host, double *bnd) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double t = clock ? clock[0] : t_host;
  int k = -1;
  for (int i = 0; i < K; ++i) if (times[i] <= t) k = i;   // K is small; every thread takes the same path
  if (k < 0) return;
  int s = sid ? sid[p] : 0;
  s = s < 0 ? 0 : (s >= S ? S - 1 : s);
  const double *c = cmd + ((size_t)s * K + k) * 3;
  wt_apply_commands_one(c[0], c[1], c[2], bnd, (size_t)P, p);
}

// 8 independent DFMA chains per thread: saturates the FP64 pipe without memory traffic
__global__ void wt_dfma_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 12345.678) out[0] = s;  // never true; keeps the loop alive
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int wt_abi_version(void) { return WT_ABI_VERSION; }
const char *wt_last_error(void) { return g_err; }
int wt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int check_common(int P, int n) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (n < 2 || n > WT_MAX_ZONES) return set_err(WT_ERR_BAD_ARG, "n_zones must be in [2, 32]");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  return 0;
}

struct StepKernels { void (*begin)(StepArgs); void (*run)(StepArgs); void (*run_floor)(StepArgs); void (*catch_up)(StepArgs); void (*catch_up_floor)(StepArgs); };
static StepKernels step_kernels(int n) {
  if (getenv("WT_B200_GENERIC_N")) n = 0;  // A/B runs
  // (the catch-up kernel handles a few hundred plants: only the headline shape n = 10 gets its own instantiation)
  if (n == 10) return {wt_step_begin_kernel<WT_STEP_WARPS, 10>, wt_step_run_kernel<WT_STEP_WARPS, 10, false>, wt_step_run_kernel<WT_STEP_WARPS, 10, true>,
                       wt_catch_up_kernel<WT_STEP_WARPS, 10, false>, wt_catch_up_kernel<WT_STEP_WARPS, 10, true>};
  if (n == 20) return {wt_step_begin_kernel<WT_STEP_WARPS, 20>, wt_step_run_kernel<WT_STEP_WARPS, 20, false>, wt_step_run_kernel<WT_STEP_WARPS, 20, true>,
                       wt_catch_up_kernel<WT_STEP_WARPS, 0, false>, wt_catch_up_kernel<WT_STEP_WARPS, 0, true>};
  return {wt_step_begin_kernel<WT_STEP_WARPS, 0>, wt_step_run_kernel<WT_STEP_WARPS, 0, false>, wt_step_run_kernel<WT_STEP_WARPS, 0, true>,
          wt_catch_up_kernel<WT_STEP_WARPS, 0, false>, wt_catch_up_kernel<WT_STEP_WARPS, 0, true>};
}

// per-device launch facts, set up once per device under a lock (several host threads, one per device, may call in)
struct DevInfo { bool done; int sms; };
static int device_info(DevInfo *out) {
  static std::mutex mu;
  static DevInfo info[WT_MAX_DEVICES];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= WT_MAX_DEVICES) return cuda_err(e != cudaSuccess ? e : cudaErrorInvalidDevice, "cudaGetDevice");
  std::lock_guard<std::mutex> lock(mu);
  if (!info[dev].done) {
    // the attributes belong to the (kernel, device) pair
    const int shapes[3] = {0, 10, 20};
    int pct = WT_STEP_CARVEOUT_PCT;
    if (const char *ev = getenv("WT_B200_CARVEOUT_PCT")) pct = atoi(ev);  // tuning runs
    for (int i = 0; i < 3; ++i) {
      StepKernels k = step_kernels(shapes[i]);
      e = cudaFuncSetAttribute(k.run, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.begin, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      // shared memory for the resident blocks, the rest of the 256 KB stays L1 (it backs the register spills)
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.run, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.run_floor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.run_floor, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.catch_up, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k.catch_up_floor, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return cuda_err(e, "cudaFuncSetAttribute");
    }
    e = cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_err(e, "cudaDeviceGetAttribute");
    info[dev].done = true;
  }
  *out = info[dev];
  return 0;
}

static int launch_step(StepArgs a, cudaStream_t s) {
  a.inv_sqrtN = 1.0 / sqrt((double)(3 * a.n));
  a.inv_sqrt3N = 1.0 / sqrt((double)(9 * a.n));
  const int gpw = 32 / a.n;
  const long long groups = ((long long)a.P + gpw - 1) / gpw;
  a.n_groups = (int)groups;
  const long long blocks = (groups + WT_STEP_WARPS - 1) / WT_STEP_WARPS;
  DevInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  int per_sm = WT_STEP_MINBLOCKS;
  if (const char *ev = getenv("WT_B200_RUN_BLOCKS_PER_SM")) per_sm = atoi(ev);  // tuning runs
  long long run_blocks = (long long)di.sms * per_sm;
  if (run_blocks > blocks) run_blocks = blocks;
  const size_t smem_run = (size_t)WT_STEP_WARPS * wt_warp_smem_doubles(a.n) * sizeof(double);
  const size_t smem_begin = (size_t)WT_STEP_WARPS * wt_begin_smem_doubles(a.n) * sizeof(double);
  const StepKernels k = step_kernels(a.n);
  const int n_steps = a.n_steps;
  for (int i = 0; i < n_steps; ++i) {
    a.n_steps = i;  // > 0: OR the non-halting status bits of the earlier steps into this step's word
    k.begin<<<(unsigned)blocks, WT_STEP_WARPS * 32, smem_begin, s>>>(a);
    (a.h_floor > 0.0 ? k.run_floor : k.run)<<<(unsigned)run_blocks, WT_STEP_WARPS * 32, smem_run, s>>>(a);
  }
  return cuda_err(cudaGetLastError(), "wt_step kernels launch");
}

size_t wt_step_workspace_bytes(int P, int n) {
  if (P <= 0 || n < 2 || n > WT_MAX_ZONES) return 0;
  const int gpw = 32 / n;
  return wt_ws_bytes(((long long)P + gpw - 1) / gpw);
}

int wt_advance(int P, int n, int n_steps, double dt, const double *par, const double *bnd, int bnd_stride,
               double *time, double *y, double *flow, double *derived, uint32_t *status, int32_t *counters,
               int max_attempts, const int32_t *order, int32_t *cost, void *workspace, void *stream) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!(dt > 0.0)) return set_err(WT_ERR_BAD_ARG, "dt must be positive");
  if (n_steps < 1) return set_err(WT_ERR_BAD_ARG, "n_steps must be >= 1");
  if (!par || !bnd || !time || !y || !status) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  if (!workspace) return set_err(WT_ERR_BAD_ARG, "null workspace (wt_step_workspace_bytes)");
  if (bnd_stride != 0 && bnd_stride != P) return set_err(WT_ERR_BAD_ARG, "bnd_stride must be 0 or P");
  StepArgs a;
  a.P = P; a.ld = P; a.n = n; a.n_steps = n_steps; a.bnd_stride = bnd_stride; a.max_attempts = max_attempts;
  a.dt = dt; a.par = par; a.bnd = bnd; a.time = time; a.y = y; a.flow = flow; a.derived = derived;
  a.status = status; a.counters = counters; a.order = order; a.cost = cost; a.ws = (char *)workspace;
  a.skip_mask = WTS_SKIP_MASK; a.count_dev = nullptr; a.t_stop = nullptr; a.h_floor = 0.0;
  return launch_step(a, (cudaStream_t)stream);
}

int wt_catch_up(int cap, int ld, int n, int n_steps, double dt, const double *par, const double *bnd, int bnd_stride,
                double *time, double *y, double *flow, double *derived, uint32_t *status, int32_t *counters,
                int max_attempts, int floor_div, const int32_t *list, const int32_t *count, const double *t_stop,
                void *workspace, void *stream) {
  int rc = check_common(cap, n);
  if (floor_div < 0) return set_err(WT_ERR_BAD_ARG, "floor_div must be >= 0");
  if (rc) return rc;
  if (!(dt > 0.0) || n_steps < 1 || ld < 1) return set_err(WT_ERR_BAD_ARG, "bad dt, n_steps or ld");
  (void)workspace;  // (ABI v3 ran the two step kernels per step and needed their hand-off rows; the fused kernel does not)
  if (!par || !bnd || !time || !y || !status || !list || !count || !t_stop)
    return set_err(WT_ERR_BAD_ARG, "null device pointer");
  if (bnd_stride != 0 && bnd_stride != ld) return set_err(WT_ERR_BAD_ARG, "bnd_stride must be 0 or ld");
  StepArgs a;
  a.P = cap; a.ld = ld; a.n = n; a.n_steps = n_steps; a.bnd_stride = bnd_stride; a.max_attempts = max_attempts;
  a.dt = dt; a.par = par; a.bnd = bnd; a.time = time; a.y = y; a.flow = flow; a.derived = derived;
  a.status = status; a.counters = counters; a.order = list; a.cost = nullptr; a.ws = (char *)workspace;
  a.skip_mask = WTS_HALT_MASK;   // the listed plants carry WTS_DEFERRED: that is what this launch is for
  a.count_dev = count; a.t_stop = t_stop;
  a.h_floor = floor_div > 0 ? dt / (double)floor_div : 0.0;
  // ONE launch for all n_steps of the listed plants (wt_catch_up_kernel); the workspace is not needed by it
  a.inv_sqrtN = 1.0 / sqrt((double)(3 * n));
  a.inv_sqrt3N = 1.0 / sqrt((double)(9 * n));
  const int gpw = 32 / n;
  const long long groups = ((long long)cap + gpw - 1) / gpw;
  a.n_groups = (int)groups;
  DevInfo di;
  rc = device_info(&di);
  if (rc) return rc;
  const StepKernels k = step_kernels(n);
  const size_t smem_run = (size_t)WT_STEP_WARPS * wt_warp_smem_doubles(n) * sizeof(double);
  (a.h_floor > 0.0 ? k.catch_up_floor : k.catch_up)<<<(unsigned)((groups + WT_STEP_WARPS - 1) / WT_STEP_WARPS), WT_STEP_WARPS * 32, smem_run,
                                                       (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "wt_catch_up_kernel launch");
}

int wt_defer_collect(int P, uint32_t *status, int32_t *list, int32_t *count, int cap, double *t_stop, double t_stop_inc,
                     void *stream) {
  if (P <= 0 || cap <= 0) return set_err(WT_ERR_BAD_ARG, "P and cap must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!status || !list || !count) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_defer_collect_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, status, list, count, cap, t_stop, t_stop_inc);
  return cuda_err(cudaGetLastError(), "wt_defer_collect_kernel launch");
}

int wt_defer_rejoin(uint32_t *status, const int32_t *list, int32_t *count, int cap, void *stream) {
  if (cap <= 0) return set_err(WT_ERR_BAD_ARG, "cap must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!status || !list || !count) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_defer_rejoin_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(status, list, count, cap);
  return cuda_err(cudaGetLastError(), "wt_defer_rejoin_kernel launch");
}

int wt_step(int P, int n, double dt, const double *par, const double *bnd, int bnd_stride, double *time,
            double *y, double *flow, double *derived, uint32_t *status, int32_t *counters, int max_attempts,
            void *workspace, void *stream) {
  return wt_advance(P, n, 1, dt, par, bnd, bnd_stride, time, y, flow, derived, status, counters, max_attempts, nullptr,
                    nullptr, workspace, stream);
}

int wt_derivatives(int P, int n, const double *par, const double *bnd, int bnd_stride, const double *y,
                   double *dy, int32_t *bad, void *stream) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!par || !bnd || !y || !dy) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  const int gpw = 32 / n;
  const long long warps = ((long long)P + gpw - 1) / gpw;
  const int wpb = 4;
  const long long blocks = (warps + wpb - 1) / wpb;
  wt_derivatives_kernel<<<(unsigned)blocks, wpb * 32, 0, (cudaStream_t)stream>>>(P, n, par, bnd, bnd_stride, y, dy, bad);
  return cuda_err(cudaGetLastError(), "wt_derivatives_kernel launch");
}

int wt_calc_ph(int P, const double *alk, const double *ct, const double *temp, const double *guess, double *ph,
               int32_t *iters, int32_t *status, void *stream) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!alk || !ct || !temp || !guess || !ph || !iters || !status) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  const int tpb = 128;
  DevInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  // persistent warps: 8 blocks per SM (fewer when there are fewer systems than lanes); the queue counter is a
  // stream-ordered allocation, so concurrent calls on different streams do not share it
  long long blocks = (long long)di.sms * 8;
  const long long want = ((long long)P / 2 + tpb - 1) / tpb;   // at least ~2 systems per lane
  if (blocks > want) blocks = want < 1 ? 1 : want;
  // the queue counter: one of 64 device words handed out round robin, so that calls in flight on different streams
  // do not share one (a stream-ordered allocation per call cost 12 ms with the default pool settings)
  static std::atomic<unsigned> next_slot{0};
  int *queues = nullptr;
  cudaError_t e = cudaGetSymbolAddress((void **)&queues, wt_ph_queues);
  if (e != cudaSuccess) return cuda_err(e, "cudaGetSymbolAddress");
  int *queue = queues + (next_slot.fetch_add(1) % 64);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(queue, 0, sizeof(int), st);
  wt_calc_ph_kernel<<<(unsigned)blocks, tpb, 0, st>>>(P, alk, ct, temp, guess, ph, iters, status, queue);
  return cuda_err(cudaGetLastError(), "wt_calc_ph_kernel launch");
}


int wt_diagnostics(int P, int n, const double *par, const double *y, const double *h, double *out, double *n2,
                   int32_t *bad, void *stream) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!par || !y || !out) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_diagnostics_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, n, par, y, h, out, n2, bad);
  return cuda_err(cudaGetLastError(), "wt_diagnostics_kernel launch");
}

int wt_register_image(int K, const int32_t *sel, int P, const double *value, const int32_t *fault, double sim_time,
                      uint16_t *ir, uint8_t *di, uint8_t *ok, void *stream) {
  if (K <= 0 || P <= 0) return set_err(WT_ERR_BAD_ARG, "K and P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!sel || !value || !fault || !ir || !di) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_register_image_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(K, sel, P, value, fault, sim_time, ir, di, ok);
  return cuda_err(cudaGetLastError(), "wt_register_image_kernel launch");
}

int wt_sensor_window_stats(int P, int m, const double *hist, const int32_t *rows, int sensor, double *out, void *stream) {
  if (P <= 0 || m < 0 || sensor < 0 || sensor >= WT_NSENS) return set_err(WT_ERR_BAD_ARG, "bad P, m or sensor index");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!out || (m > 0 && (!hist || !rows))) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_sensor_window_stats_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, m, hist, rows, sensor, out);
  return cuda_err(cudaGetLastError(), "wt_sensor_window_stats_kernel launch");
}

int wt_apply_commands(int P, const double *acid, const double *chlor, const double *inlet, double *bnd_soa, void *stream) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!acid || !chlor || !inlet || !bnd_soa) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_apply_commands_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, acid, chlor, inlet, bnd_soa);
  return cuda_err(cudaGetLastError(), "wt_apply_commands_kernel launch");
}

int wt_scenario_commands(int P, int K, int S, const double *times, const double *cmd, const int32_t *sid, const double *clock,
                         double t, double *bnd_soa, void *stream) {
  if (P <= 0 || K <= 0 || S <= 0) return set_err(WT_ERR_BAD_ARG, "P, K and S must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!times || !cmd || !bnd_soa) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_scenario_commands_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, K, S, times, cmd, sid, clock, t, bnd_soa);
  return cuda_err(cudaGetLastError(), "wt_scenario_commands_kernel launch");
}

int wt_stats_size(int n) { return WT_STATS_HDR + 6 * n; }
int wt_stats_scratch_doubles(int n) { return 1024 * (WT_STATS_HDR + 6 * n); }

int wt_sum_rows(int rows, int n, const double *in, double *out, void *stream) {
  if (rows <= 0 || n <= 0) return set_err(WT_ERR_BAD_ARG, "rows and n must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!in || !out) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_stats_final_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rows, n, in, out, 0);
  return cuda_err(cudaGetLastError(), "wt_sum_rows launch");
}

int wt_stats(int P, int n, const double *y, const uint32_t *status, const double *shift_thr, double *out,
             double *scratch, int accumulate, void *stream) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!y || !status || !shift_thr || !out || !scratch) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  int blocks = (P + 4 * WT_STATS_TPB - 1) / (4 * WT_STATS_TPB);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  const int nstat = wt_stats_size(n);
  wt_stats_partial_kernel<<<blocks, WT_STATS_TPB, 0, (cudaStream_t)stream>>>(P, n, y, status, shift_thr, scratch);
  wt_stats_final_kernel<<<(nstat + 127) / 128, 128, 0, (cudaStream_t)stream>>>(blocks, nstat, scratch, out, accumulate);
  return cuda_err(cudaGetLastError(), "wt_stats launch");
}


// ---- K3: sensor suite ---------------------------------------------------------------------------
__global__ void wt_sensors_calibrate_kernel(int P, int sensor, double t, const double *ref, double ref_scalar,
                                            double *sens, int *sens_i) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  double *S = sens + (size_t)sensor * Pz + p;
  const double r = ref ? ref[p] : ref_scalar;
  // BaseSensor.calibrate (base_sensor.py:701-755): offset = reference - current_value, timers reset
  S[(size_t)WT_SF_CALOFF * WT_NSENS * Pz] = r - S[(size_t)WT_SF_CUR * WT_NSENS * Pz];
  S[(size_t)WT_SF_TCAL * WT_NSENS * Pz] = t;
  S[(size_t)WT_SF_TPOWER * WT_NSENS * Pz] = t;
  sens_i[(size_t)sensor * Pz + p] = SS_NORMAL;
  sens_i[((size_t)WT_SI_FAULT * WT_NSENS + sensor) * Pz + p] = SFLT_NONE;
  sens_i[((size_t)WT_SI_FLAGS * WT_NSENS + sensor) * Pz + p] &= ~1;   // calibration_history.append(record)
}

// BaseSensor.reset (base_sensor.py:858-878) for sensor `sensor` of every plant, with the simulated time t where the
// reference stamps time.monotonic(): value to mid-range, offset 0, histories cleared (every later read reports
// CALIBRATION_EXPIRED until calibrate()), status / fault cleared, warm-up restarted, the sensor's sample line emptied
// (the pH and temperature sensors of one side share the line object, sensors/__init__.py:62-67).
__global__ void wt_sensors_reset_kernel(int P, int sensor, double t, const double *cfg_flow, double *sens, int *sens_i, int *ring_i) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  const int type = wt_sensor_type(sensor);
  double *S = sens + (size_t)sensor * Pz + p;
  const double lo = type == ST_TEMP_RTD ? -10.0 : 0.0;
  const double hi = type == ST_PH ? 14.0 : (type == ST_TEMP_RTD ? 110.0 : (type == ST_FLOW_MAG ? cfg_flow[p] * 2.0 : 10.0));
  S[(size_t)WT_SF_CUR * WT_NSENS * Pz] = (lo + hi) / 2.0;
  S[(size_t)WT_SF_CALOFF * WT_NSENS * Pz] = 0.0;
  S[(size_t)WT_SF_TCAL * WT_NSENS * Pz] = t;
  S[(size_t)WT_SF_TPOWER * WT_NSENS * Pz] = t;
  S[(size_t)WT_SF_LASTVAL * WT_NSENS * Pz] = nan("");
  sens_i[(size_t)sensor * Pz + p] = SS_NORMAL;
  sens_i[((size_t)WT_SI_FAULT * WT_NSENS + sensor) * Pz + p] = SFLT_NONE;
  sens_i[((size_t)WT_SI_HIST * WT_NSENS + sensor) * Pz + p] = 0;
  sens_i[((size_t)WT_SI_FLAGS * WT_NSENS + sensor) * Pz + p] |= 1;
  const int line = (sensor == 0 || sensor == 5) ? 0 : ((sensor == 1 || sensor == 6) ? 1 : -1);
  if (line >= 0) { ring_i[((size_t)line * 2 + 0) * Pz + p] = 0; ring_i[((size_t)line * 2 + 1) * Pz + p] = 0; }
}

// device clock of a sensor suite {t, t_prev, read_index, t0, dt}: advanced in stream order after every read of a
// captured (CUDA graph) step, so that replays need no new launch arguments.  t = t0 + k dt, as the host loop computes it.
__global__ void wt_clock_tick_kernel(double *clock) {
  const double k = clock[2] + 1.0;
  clock[1] = clock[0];
  clock[2] = k;
  clock[0] = clock[3] + k * clock[4];
}

// Per-sensor ensemble statistics of the LAST suite read (the sensor half of the all-reduce payload, SURVEY.md 8e;
// reference analogue BaseSensor.get_statistics, base_sensor.py:809-856, here across plants instead of across time):
//   out[s * WT_SSTAT_N + 0] readings with a finite value   [1] sum (value - shift_s)   [2] sum (value - shift_s)^2
//   [3 .. 14] SensorStatus histogram   [15 .. 21] SensorFault histogram          (s = sensor 0..6; halted plants skipped)
// One block per (sensor, slab) -> partials summed in a fixed order by the final kernel: bitwise reproducible.
#define WT_SSTAT_N 22
__global__ void __launch_bounds__(WT_STATS_TPB) wt_sensor_stats_partial_kernel(int P, const double *value, const int32_t *st,
                                                                                 const int32_t *ft, const uint32_t *plant_status,
                                                                                 const double *shift7, double *partial) {
  __shared__ double sh[WT_STATS_TPB / 32];
  __shared__ int hist[12 + 7];
  const int s = blockIdx.y;
  const int per_block = (P + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * per_block, hi = min(P, lo + per_block);
  if (threadIdx.x < 19) hist[threadIdx.x] = 0;
  __syncthreads();
  const double c = shift7[s];
  double nv = 0, s1 = 0, s2 = 0;
  for (int p = lo + threadIdx.x; p < hi; p += WT_STATS_TPB) {
    if (plant_status[p] & WTS_SKIP_MASK) continue;
    const double v = value[(size_t)s * P + p];
    if (isfinite(v)) { const double d = v - c; nv += 1.0; s1 += d; s2 += d * d; }
    const int a = st[(size_t)s * P + p], b = ft[(size_t)s * P + p];
    if (a >= 0 && a < 12) atomicAdd(&hist[a], 1);
    if (b >= 0 && b < 7) atomicAdd(&hist[12 + b], 1);
  }
  double *out = partial + ((size_t)blockIdx.x * WT_NSENS + s) * WT_SSTAT_N;
  double r;
  r = block_sum(nv, sh); if (threadIdx.x == 0) out[0] = r;
  r = block_sum(s1, sh); if (threadIdx.x == 0) out[1] = r;
  r = block_sum(s2, sh); if (threadIdx.x == 0) out[2] = r;
  if (threadIdx.x < 19) out[3 + threadIdx.x] = (double)hist[threadIdx.x];
}

// Counting sort of the plants by the work of their last step, most expensive first (scheduling only: stragglers
// start first and plants with similar solver paths share a warp).  Keys are small integers (collocation solves +
// Newton iterations), so three tiny kernels replace a general radix sort: histogram, exclusive scan over the
// descending keys, scatter.  The order inside a key is whatever the atomics give -- results never depend on it.
#define WT_COST_BINS 1024
__global__ void wt_cost_hist_kernel(int P, const int32_t *cost, int32_t *bins) {
  __shared__ int sh[WT_COST_BINS];
  for (int i = threadIdx.x; i < WT_COST_BINS; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    int c = cost[p];
    c = c < 0 ? 0 : (c >= WT_COST_BINS ? WT_COST_BINS - 1 : c);
    atomicAdd(&sh[c], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WT_COST_BINS; i += blockDim.x) if (sh[i]) atomicAdd(&bins[i], sh[i]);
}
__global__ void wt_cost_scan_kernel(int32_t *bins) {  // one block of WT_COST_BINS threads: bins[c] <- first slot of key c, keys descending
  __shared__ int sh[WT_COST_BINS];
  const int i = threadIdx.x;
  sh[i] = bins[WT_COST_BINS - 1 - i];  // descending
  __syncthreads();
  for (int o = 1; o < WT_COST_BINS; o <<= 1) {
    const int v = i >= o ? sh[i - o] : 0;
    __syncthreads();
    sh[i] += v;
    __syncthreads();
  }
  bins[WT_COST_BINS - 1 - i] = sh[i] - (i == 0 ? sh[0] : sh[i] - sh[i - 1]);  // exclusive
}
// Scatter: almost all plants share a handful of keys, so one global atomic per plant serialises on ~10 addresses
// (0.15 ms per 262,144 plants, 4 % of a step).  A block ranks its plants per key in shared memory, reserves one range
// per (block, key) with a single global atomic, and scatters: ~40x fewer global atomics, none of them hot.
__global__ void __launch_bounds__(WT_COST_BINS) wt_cost_scatter_kernel(int P, const int32_t *cost, int32_t *cursor, int32_t *order) {
  __shared__ int cnt[WT_COST_BINS], base[WT_COST_BINS];
  cnt[threadIdx.x] = 0;
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  int c = 0, rank = 0;
  if (p < P) {
    c = cost[p];
    c = c < 0 ? 0 : (c >= WT_COST_BINS ? WT_COST_BINS - 1 : c);
    rank = atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  if (cnt[threadIdx.x]) base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]);
  __syncthreads();
  if (p < P) order[base[c] + rank] = p;
}

// Maintenance operations of the reference sensors, for sensor `sensor` of every plant:
//   op 0  pHSensor.calibrate_two_point(b1, b2, m1, m2, t)   ph_sensor.py:338-393
//   op 1  pHSensor.clean_electrode(method, t)               ph_sensor.py:395-434   a0 = 0 water_rinse, 1 acid_clean, 2 pepsin_clean
//   op 2  ChlorineSensor.replace_membrane(t)                chlorine_sensor.py:486-509
//   op 3  ChlorineSensor.replace_reagent(t)                 chlorine_sensor.py:511-537
// (slope_percentage is overwritten by every read, ph_sensor.py:256-262, and glass_etching is not used on the
// read path, so neither is device state.)
__global__ void wt_sensors_maintain_kernel(int P, int sensor, int op, double t, double a0, double a1, double *sens, int *sens_i) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  double *S = sens + (size_t)sensor * Pz + p;
#define MF(f) S[(size_t)(f) * WT_NSENS * Pz]
  double reference = 0.0;
  bool recal = true;
  if (op == 0) {                       // two-point: junction cleaned, single-point calibration at the mid buffer
    MF(WT_SF_AUX1) = 0.0;
    reference = (a0 + a1) / 2.0;
  } else if (op == 1) {                // cleaning: fouling removed, warm-up restarts, calibration untouched
    const int m = (int)a0;
    MF(WT_SF_AUX0) *= (m == 0 ? 0.5 : (m == 1 ? 0.1 : 0.2));
    MF(WT_SF_AUX2) = 0.0;
    MF(WT_SF_TPOWER) = t;
    recal = false;
  } else if (op == 2) {                // new membrane
    MF(WT_SF_AUX0) = 0.0;
    MF(WT_SF_AUX1) = 0.0;
  } else {                             // new reagent
    MF(WT_SF_AUX0) = 1.0;
    MF(WT_SF_AUX1) = 0.0;
    MF(WT_SF_AUX2) = 0.0;
  }
  if (recal) {                         // BaseSensor.calibrate(reference, t), base_sensor.py:701-755
    MF(WT_SF_CALOFF) = reference - MF(WT_SF_CUR);
    MF(WT_SF_TCAL) = t;
    MF(WT_SF_TPOWER) = t;
    sens_i[(size_t)sensor * Pz + p] = SS_NORMAL;
    sens_i[((size_t)WT_SI_FAULT * WT_NSENS + sensor) * Pz + p] = SFLT_NONE;
    sens_i[((size_t)WT_SI_FLAGS * WT_NSENS + sensor) * Pz + p] &= ~1;
  }
#undef MF
}

int wt_sensors_maintain(int P, int sensor, int op, double t, double a0, double a1, double *sens, int32_t *sens_i, void *stream) {
  if (P <= 0 || sensor < 0 || sensor >= WT_NSENS) return set_err(WT_ERR_BAD_ARG, "bad P or sensor index");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!sens || !sens_i) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  const int type = wt_sensor_type(sensor);
  // the reference raises ValueError for these (ph-only methods exist only on pHSensor)
  if ((op == 0 || op == 1) && type != ST_PH) return set_err(WT_ERR_BAD_ARG, "calibrate_two_point / clean_electrode: not a pH sensor");
  if (op == 1 && !(a0 == 0.0 || a0 == 1.0 || a0 == 2.0)) return set_err(WT_ERR_BAD_ARG, "Unknown cleaning method");
  if (op == 2 && type != ST_CL_AMP) return set_err(WT_ERR_BAD_ARG, "Only amperometric sensors have membranes");
  if (op == 3 && type != ST_CL_DPD) return set_err(WT_ERR_BAD_ARG, "Only DPD sensors have reagent");
  if (op < 0 || op > 3) return set_err(WT_ERR_BAD_ARG, "unknown maintenance operation");
  wt_sensors_maintain_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, sensor, op, t, a0, a1, sens, sens_i);
  return cuda_err(cudaGetLastError(), "wt_sensors_maintain_kernel launch");
}

int wt_sensors_init(int P, double t0, const double *cfg_flow, const double *cfg_cl, const double *cfg_T, double *sens,
                    int32_t *sens_i, int32_t *ring_i, void *stream) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!cfg_flow || !cfg_cl || !cfg_T || !sens || !sens_i || !ring_i) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_sensors_init_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, t0, cfg_flow, cfg_cl, cfg_T, sens, sens_i, ring_i);
  return cuda_err(cudaGetLastError(), "wt_sensors_init_kernel launch");
}

int wt_sensors_calibrate(int P, int sensor, double t, const double *ref_dev, double ref_scalar, double *sens,
                         int32_t *sens_i, void *stream) {
  if (P <= 0 || sensor < 0 || sensor >= WT_NSENS) return set_err(WT_ERR_BAD_ARG, "bad P or sensor index");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!sens || !sens_i) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_sensors_calibrate_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, sensor, t, ref_dev, ref_scalar, sens, sens_i);
  return cuda_err(cudaGetLastError(), "wt_sensors_calibrate_kernel launch");
}

int wt_sensors_read(int P, int n, long long plant0, unsigned read_index, double t, double t_prev, const double *y,
                    const double *flow, const double *cfg_flow, const double *cfg_cl, const double *cfg_T, double *sens,
                    int32_t *sens_i, double *ring, int32_t *ring_i, double *out, int32_t *out_status, int32_t *out_fault,
                    const double *suite8, uint64_t seed, const double *clock_dev, void *stream) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!y || !flow || !cfg_flow || !cfg_cl || !cfg_T || !sens || !sens_i || !ring || !ring_i || !out || !out_status ||
      !out_fault || !suite8)
    return set_err(WT_ERR_BAD_ARG, "null pointer");
  const double *suite6 = suite8;
  if (!(suite8[6] >= 0.0 && suite8[6] <= 3.0) || !(suite8[7] == 0.0 || suite8[7] == 1.0))
    return set_err(WT_ERR_BAD_ARG, "unknown temperature / flow sensor type");
  SensorArgs a;
  a.P = P; a.n = n; a.plant0 = plant0; a.read_index = read_index; a.t = t; a.t_prev = t_prev; a.clock = clock_dev;
  a.s.temp_kind = (int)suite8[6]; a.s.flow_kind = (int)suite8[7];
  a.y = y; a.flow = flow; a.cfg_flow = cfg_flow; a.cfg_cl = cfg_cl; a.cfg_T = cfg_T;
  a.sens = sens; a.sens_i = sens_i; a.ring = ring; a.ring_i = ring_i; a.out = out; a.out_status = out_status;
  a.out_fault = out_fault;
  a.s.flow_velocity = suite6[0]; a.s.bubble_per_min = suite6[1]; a.s.grounding = suite6[2]; a.s.vibration_g = suite6[3];
  a.s.ambient_temp = suite6[4]; a.s.line_delay_s = suite6[5];
  a.s.seed_lo = (uint32_t)seed; a.s.seed_hi = (uint32_t)(seed >> 32);
  wt_sensors_read_kernel<<<dim3((P + 127) / 128, 5), 128, 0, (cudaStream_t)stream>>>(a);
  return cuda_err(cudaGetLastError(), "wt_sensors_read_kernel launch");
}

int wt_sensors_reset(int P, int sensor, double t, const double *cfg_flow, double *sens, int32_t *sens_i, int32_t *ring_i,
                     void *stream) {
  if (P <= 0 || sensor < 0 || sensor >= WT_NSENS) return set_err(WT_ERR_BAD_ARG, "bad P or sensor index");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!cfg_flow || !sens || !sens_i || !ring_i) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  wt_sensors_reset_kernel<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, sensor, t, cfg_flow, sens, sens_i, ring_i);
  return cuda_err(cudaGetLastError(), "wt_sensors_reset_kernel launch");
}

int wt_clock_tick(double *clock_dev, void *stream) {
  if (!clock_dev) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  wt_clock_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(clock_dev);
  return cuda_err(cudaGetLastError(), "wt_clock_tick_kernel launch");
}

int wt_sensor_stats_size(void) { return WT_NSENS * WT_SSTAT_N; }
int wt_sensor_stats_scratch_doubles(void) { return 256 * WT_NSENS * WT_SSTAT_N; }
int wt_sensor_stats(int P, const double *out_value, const int32_t *out_status, const int32_t *out_fault,
                    const uint32_t *plant_status, const double *shift7, double *stats, double *scratch, int accumulate,
                    void *stream) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!out_value || !out_status || !out_fault || !plant_status || !shift7 || !stats || !scratch)
    return set_err(WT_ERR_BAD_ARG, "null device pointer");
  int blocks = (P + 8 * WT_STATS_TPB - 1) / (8 * WT_STATS_TPB);
  if (blocks > 256) blocks = 256;
  const int nstat = WT_NSENS * WT_SSTAT_N;
  wt_sensor_stats_partial_kernel<<<dim3(blocks, WT_NSENS), WT_STATS_TPB, 0, (cudaStream_t)stream>>>(P, out_value, out_status, out_fault,
                                                                                              plant_status, shift7, scratch);
  wt_stats_final_kernel<<<(nstat + 127) / 128, 128, 0, (cudaStream_t)stream>>>(blocks, nstat, scratch, stats, accumulate);
  return cuda_err(cudaGetLastError(), "wt_sensor_stats launch");
}

int wt_cost_order(int P, const int32_t *cost_dev, int32_t *order_dev, int32_t *bins_dev, void *stream) {
  if (P <= 0) return set_err(WT_ERR_BAD_ARG, "P must be positive");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  if (!cost_dev || !order_dev || !bins_dev) return set_err(WT_ERR_BAD_ARG, "null device pointer");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(bins_dev, 0, WT_COST_BINS * sizeof(int32_t), s);
  if (e != cudaSuccess) return cuda_err(e, "cudaMemsetAsync");
  int blocks = (P + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  wt_cost_hist_kernel<<<blocks, 256, 0, s>>>(P, cost_dev, bins_dev);
  wt_cost_scan_kernel<<<1, WT_COST_BINS, 0, s>>>(bins_dev);
  wt_cost_scatter_kernel<<<(P + WT_COST_BINS - 1) / WT_COST_BINS, WT_COST_BINS, 0, s>>>(P, cost_dev, bins_dev, order_dev);
  return cuda_err(cudaGetLastError(), "wt_cost_order launch");
}

// Host-buffer path: per device, one workspace grown on demand, one copy-in stream, one copy-out stream, three
// compute streams, a pool of events, and the shape whose constants are resident.  One call at a time per device
// (the context is locked for the call).
enum { WT_HOST_NCOMPUTE = 3, WT_HOST_MAX_SLABS = 64 };
struct HostCtx {
  std::mutex mu;
  size_t cap_bytes;
  char *dev;
  cudaStream_t s_in, s_out, s_k[WT_HOST_NCOMPUTE];
  cudaEvent_t ev_in[WT_HOST_MAX_SLABS], ev_k[WT_HOST_MAX_SLABS];
  bool ready;
  int res_P, res_n;
};
static HostCtx g_host[WT_MAX_DEVICES];

// Slab schedule of one call: plants per slab.  A pipeline over EQUAL slabs pays the upload of its first slab and the
// download of its last one in the open, and small slabs make short launches (a launch of 32,768 plants runs the step
// kernels at 0.73 of the rate of a 262,144-plant launch: VERDICT round 1).  So the slabs ramp up from `lo` plants
// (a 16,384-plant upload takes ~0.1 ms), double up to `hi`, stay there, and ramp down again at the end.
static int wt_host_slab_plan(int P, int *sizes) {
  int lo = 16384, hi = P / 4;  // (a 131,072-plant shard of an 8-GPU run: 16k, 3 x 32k, 16k)
  if (hi > 131072) hi = 131072;
  if (hi < 2 * lo) hi = 2 * lo;
  if (const char *e = getenv("WT_B200_HOST_SLAB_MIN")) lo = atoi(e);
  if (const char *e = getenv("WT_B200_HOST_SLAB_MAX")) hi = atoi(e);
  if (lo < 96) lo = 96;
  lo = (lo + 31) & ~31;
  if (hi < lo) hi = lo;
  if (const char *e = getenv("WT_B200_HOST_SLABS")) {  // tuning runs: this many equal slabs
    int k = atoi(e);
    if (k < 1) k = 1;
    if (k > WT_HOST_MAX_SLABS) k = WT_HOST_MAX_SLABS;
    const int per = ((P + k - 1) / k + 31) & ~31;
    int c = 0;
    for (int p0 = 0; p0 < P; p0 += per) sizes[c++] = (P - p0 < per) ? P - p0 : per;
    return c;
  }
  int head[16], nh = 0;
  long long ramp = 0;
  for (int sz = lo; sz < hi && nh < 16 && 2 * (ramp + sz) <= P / 2; sz *= 2) { head[nh++] = sz; ramp += sz; }
  const long long mid = (long long)P - 2 * ramp;
  int nm = (int)((mid + hi - 1) / hi);
  if (nm < 1) nm = 1;
  while (2 * nh + nm > WT_HOST_MAX_SLABS) ++hi, nm = (int)((mid + hi - 1) / hi);
  const int per = (int)(((mid + nm - 1) / nm + 31) & ~31ll);
  int c = 0;
  for (int i = 0; i < nh; ++i) sizes[c++] = head[i];
  long long left = mid;
  for (int i = 0; i < nm && left > 0; ++i) { const int w = left < per ? (int)left : per; sizes[c++] = w; left -= w; }
  for (int i = nh - 1; i >= 0; --i) sizes[c++] = head[i];
  return c;
}

int wt_step_host_plan(int P, int *sizes, int cap) {
  if (P <= 0) return 0;
  int plan[WT_HOST_MAX_SLABS];
  const int c = wt_host_slab_plan(P, plan);
  for (int i = 0; i < c && sizes && i < cap; ++i) sizes[i] = plan[i];
  return c;
}

int wt_step_host(int P, int n, double dt, const double *par, const double *bnd, int bnd_stride, double *time,
                 double *y, double *flow, uint32_t *status, int max_attempts, int flags) {
  int rc = check_common(P, n);
  if (rc) return rc;
  if (!par || !bnd || !time || !y || !status) return set_err(WT_ERR_BAD_ARG, "null host pointer");
  if (bnd_stride != 0 && bnd_stride != P) return set_err(WT_ERR_BAD_ARG, "bnd_stride must be 0 or P");
  if (!(dt > 0.0)) return set_err(WT_ERR_BAD_ARG, "dt must be positive");
  // Pipelined over column slabs of the SoA arrays (a slab of plants is a column range of every row: 2-D copies with
  // the row pitch P; plants are independent, so slabs need no ordering among themselves).  Three stages on their own
  // streams, chained per slab by events:
  //   copy-in stream   every H2D copy, slab after slab (one DMA engine direction, always busy)
  //   compute streams  slab c's two step kernels on stream c % 3 once its upload has landed; consecutive slabs sit on
  //                    different streams, so the persistent warps of the next launch take over the SMs as the blocks
  //                    of the previous one drain
  //   copy-out stream  every D2H copy, as the kernels finish (PCIe is full duplex: the other DMA direction)
  // No stage waits for a later one, so nothing but the first upload and the last download is exposed.
  int sizes[WT_HOST_MAX_SLABS];
  const int slabs = wt_host_slab_plan(P, sizes);
  int per = 0;
  for (int c = 0; c < slabs; ++c) per = sizes[c] > per ? sizes[c] : per;
  const size_t Pz = (size_t)P;
  const size_t b_par = WT_NPAR * Pz * 8, b_bnd = WT_NBND * (bnd_stride ? Pz : 1) * 8, b_t = Pz * 8,
               b_y = 3 * (size_t)n * Pz * 8, b_f = Pz * 8, b_s = Pz * 4;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t b_ws = al(wt_step_workspace_bytes(per, n));  // one step workspace per compute stream
  const size_t total = al(b_par) + al(b_bnd) + al(b_t) + al(b_y) + al(b_f) + al(b_s) + WT_HOST_NCOMPUTE * b_ws;
  int dev = 0;
  {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= WT_MAX_DEVICES) return cuda_err(e != cudaSuccess ? e : cudaErrorInvalidDevice, "cudaGetDevice");
  }
  HostCtx &g_ws = g_host[dev];
  std::lock_guard<std::mutex> lock(g_ws.mu);
  if (total > g_ws.cap_bytes) {
    flags &= ~1;
    if (g_ws.dev) cudaFree(g_ws.dev);
    g_ws.dev = nullptr;
    g_ws.cap_bytes = 0;
    cudaError_t e = cudaMalloc((void **)&g_ws.dev, total);
    if (e != cudaSuccess) { cuda_err(e, "cudaMalloc workspace"); return WT_ERR_ALLOC; }
    g_ws.cap_bytes = total;
  }
  char *q = g_ws.dev;
  double *d_par = (double *)q; q += al(b_par);
  double *d_bnd = (double *)q; q += al(b_bnd);
  double *d_t = (double *)q; q += al(b_t);
  double *d_y = (double *)q; q += al(b_y);
  double *d_f = (double *)q; q += al(b_f);
  uint32_t *d_s = (uint32_t *)q; q += al(b_s);
  char *d_ws = q;
  // WT_HOST_PARAMS_RESIDENT: the derived constants uploaded by the previous call (same P, n) are reused
  const bool up_par = !(flags & 1) || g_ws.res_P != P || g_ws.res_n != n;
  g_ws.res_P = P; g_ws.res_n = n;

  if (!g_ws.ready) {
    cudaError_t e = cudaStreamCreateWithFlags(&g_ws.s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g_ws.s_out, cudaStreamNonBlocking);
    for (int i = 0; i < WT_HOST_NCOMPUTE && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&g_ws.s_k[i], cudaStreamNonBlocking);
    for (int i = 0; i < WT_HOST_MAX_SLABS && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&g_ws.ev_in[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g_ws.ev_k[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return cuda_err(e, "wt_step_host: stream / event creation");
    g_ws.ready = true;
  }
  const cudaStream_t s_in = g_ws.s_in, s_out = g_ws.s_out;
  const size_t pitch = Pz * 8;
  if (!bnd_stride) cudaMemcpyAsync(d_bnd, bnd, b_bnd, cudaMemcpyHostToDevice, s_in);  // one broadcast row, ahead of slab 0
  int p0 = 0;
  for (int c = 0; c < slabs; ++c) {
    const int w = sizes[c];
    const size_t wb = (size_t)w * 8;
    if (up_par) cudaMemcpy2DAsync(d_par + p0, pitch, par + p0, pitch, wb, WT_NPAR, cudaMemcpyHostToDevice, s_in);
    if (bnd_stride) cudaMemcpy2DAsync(d_bnd + p0, pitch, bnd + p0, pitch, wb, WT_NBND, cudaMemcpyHostToDevice, s_in);
    cudaMemcpyAsync(d_t + p0, time + p0, wb, cudaMemcpyHostToDevice, s_in);
    cudaMemcpy2DAsync(d_y + p0, pitch, y + p0, pitch, wb, 3 * (size_t)n, cudaMemcpyHostToDevice, s_in);
    cudaMemcpyAsync(d_s + p0, status + p0, (size_t)w * 4, cudaMemcpyHostToDevice, s_in);
    if (flow) cudaMemcpyAsync(d_f + p0, flow + p0, wb, cudaMemcpyHostToDevice, s_in);
    cudaEventRecord(g_ws.ev_in[c], s_in);
    const cudaStream_t s = g_ws.s_k[c % WT_HOST_NCOMPUTE];
    cudaStreamWaitEvent(s, g_ws.ev_in[c], 0);
    StepArgs a;
    a.P = w; a.ld = P; a.n = n; a.n_steps = 1; a.bnd_stride = bnd_stride; a.max_attempts = max_attempts;
    a.dt = dt; a.par = d_par + p0; a.bnd = bnd_stride ? d_bnd + p0 : d_bnd; a.time = d_t + p0; a.y = d_y + p0;
    a.flow = flow ? d_f + p0 : nullptr; a.derived = nullptr; a.status = d_s + p0; a.counters = nullptr;
    a.order = nullptr; a.cost = nullptr; a.ws = d_ws + (size_t)(c % WT_HOST_NCOMPUTE) * b_ws;
    a.skip_mask = WTS_SKIP_MASK; a.count_dev = nullptr; a.t_stop = nullptr; a.h_floor = 0.0;
    rc = launch_step(a, s);
    if (rc) return rc;
    cudaEventRecord(g_ws.ev_k[c], s);
    cudaStreamWaitEvent(s_out, g_ws.ev_k[c], 0);
    cudaMemcpyAsync(time + p0, d_t + p0, wb, cudaMemcpyDeviceToHost, s_out);
    cudaMemcpy2DAsync(y + p0, pitch, d_y + p0, pitch, wb, 3 * (size_t)n, cudaMemcpyDeviceToHost, s_out);
    cudaMemcpyAsync(status + p0, d_s + p0, (size_t)w * 4, cudaMemcpyDeviceToHost, s_out);
    if (flow) cudaMemcpyAsync(flow + p0, d_f + p0, wb, cudaMemcpyDeviceToHost, s_out);
    p0 += w;
  }
  // every kernel is followed by copies on the copy-out stream, which therefore finishes last
  rc = cuda_err(cudaStreamSynchronize(s_out), "wt_step_host");
  if (rc) return rc;
  rc = cuda_err(cudaStreamSynchronize(s_in), "wt_step_host");
  for (int i = 0; i < WT_HOST_NCOMPUTE && !rc; ++i) rc = cuda_err(cudaStreamSynchronize(g_ws.s_k[i]), "wt_step_host");
  return rc;
}

int wt_measure_fp64_peak(double *tflops_out, int iters) {
  if (!tflops_out || iters <= 0) return set_err(WT_ERR_BAD_ARG, "bad argument");
  if (wt_device_count() <= 0) return set_err(WT_ERR_NO_DEVICE, "no CUDA device");
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double *out = nullptr;
  cudaError_t e = cudaMalloc((void **)&out, 8);
  if (e != cudaSuccess) { cuda_err(e, "cudaMalloc"); return WT_ERR_ALLOC; }
  const int tpb = 256, blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  wt_dfma_peak_kernel<<<blocks, tpb>>>(out, iters / 8 + 1, 1.0000001, 1e-9);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    wt_dfma_peak_kernel<<<blocks, tpb>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8.0 * (double)iters * (double)tpb * (double)blocks;
    double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops_out = best;
  return cuda_err(cudaGetLastError(), "wt_measure_fp64_peak");
}

}  // extern "C"
