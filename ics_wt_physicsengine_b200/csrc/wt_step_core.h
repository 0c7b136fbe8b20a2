// wt_step_core.h -- the fused plant step: one IntegratedCSTR.step(dt, boundary) per plant.
//
// What it replaces (reference, src/wt_simulator/core/reactor.py):
//   step()                      :450-509   one scipy solve_ivp(method="Radau") over [t, t+dt]
//   derivatives()               :272-448   the ODE right-hand side
//   _update_derived_state()     :511-524
//   _enforce_physical_bounds()  :526-541
// and, from scipy 1.18.1 (scipy/integrate/_ivp/), the pieces step() executes:
//   radau.py:48-136 (simplified Newton), :139-176 (step factor), :295-347 (setup),
//   :405-545 (_step_impl), :547-578 (dense output); common.py:63-134 (norm, initial step),
//   :260-382 (finite-difference Jacobian num_jac); base.py:179-210; ivp.py:659-666.
//
// B200 mapping (not a translation of the above):
//   * zones -> lanes, floor(32/n) plants per warp, everything in registers; neighbours by SHFL.
//   * Every RHS row i depends only on zones i-1, i, i+1 and dT depends on T only, dpH on
//     (pH, T), dCl on (pH_i, Cl, T).  In the ordering (T, pH, Cl) the Newton matrix
//     gamma*I - J is block lower triangular with TRIDIAGONAL blocks, so LAPACK's dense
//     real + complex LU becomes six scalar tridiagonal factorizations, done across lanes by
//     parallel cyclic reduction (PCR).  Algebraically the same linear solves.
//   * The finite-difference Jacobian is rebuilt from perturbed evaluations of the affected
//     rows only (rows outside the 3-zone stencil difference to an exact 0 in the reference).
//     scipy's column logic (first arg-max row, retry with 10x step, factor adaptation) is
//     applied unchanged on the lane that owns the column.
//   * Plants sharing a warp run in lockstep; per-plant decisions are masks.
//
// Written against wt_simt.h only, so the same source is the sm_100a kernel body and the
// CPU lane-emulation test build.
#pragma once

#include "wt_simt.h"

// parameter / boundary / status / counter indices (shared with include/wt_b200.h)
enum {
  WTP_KW = 0, WTP_KA1, WTP_KA2, WTP_KACL, WTP_CT, WTP_KX, WTP_V, WTP_ZH, WTP_VZL, WTP_VOLUME,
  WTP_AT, WTP_STRAT, WTP_NPAR
};
enum {
  WTB_INLET_FLOW = 0, WTB_INLET_PH, WTB_INLET_CL, WTB_INLET_T, WTB_ACID_FLOW, WTB_ACID_CONC,
  WTB_CL_FLOW, WTB_CL_CONC, WTB_AMBIENT_T, WTB_HEAT_LOSS, WTB_NBND
};
enum {
  WTS_SOLVER_FAILED = 1, WTS_T_RANGE = 2, WTS_CLIP_PH = 4, WTS_CLIP_CL = 8, WTS_CLIP_T = 16,
  WTS_NONFINITE = 32, WTS_T_RANGE_DERIVED = 64, WTS_WORK_LIMIT = 128, WTS_DEFERRED = 256, WTS_DEGRADED = 512
};
#define WTS_HALT_MASK (WTS_T_RANGE | WTS_WORK_LIMIT)
#define WTS_SKIP_MASK (WTS_HALT_MASK | WTS_DEFERRED)   // what an ordinary launch passes over
enum {
  WTC_NFEV = 0, WTC_NJEV, WTC_NLU, WTC_NSTEPS, WTC_NNEWTON, WTC_NREJECT, WTC_NNEWTON_FAIL,
  WTC_JAC_RETRY, WTC_NCNT
};

// The three stage evaluations of a Newton iteration stay a ROLLED loop: the kernel is bound by instruction
// fetch (SM instruction-cache hit rate ~80 %, GPC-level cache at 92 % of its request rate), so one copy of the
// RHS in the hot loop beats three (measured); -DWT_UNROLL_STAGES unrolls for experiments.
#ifdef WT_UNROLL_STAGES
#define WT_STAGE_UNROLL WT_UNROLL
#else
#define WT_STAGE_UNROLL WT_NOUNROLL
#endif
#define WT_RTOL 1e-6  // reactor.py:482
#define WT_ATOL 1e-8  // reactor.py:483
#define WT_LN10 2.302585092994046
#define WT_EPS 2.220446049250313e-16

// radau.py:11-45.  On the GPU the tables live in __constant__ memory so that DFMA/DMUL take them
// as c[bank][offset] operands instead of materialising 64-bit immediates (two UMOV each).
#define WT_S6 2.449489742783178
#define WT_RADAU_TABLE {                                                                      \
    (4.0 - WT_S6) / 10.0, (4.0 + WT_S6) / 10.0,                                  /* 0 C0 C1 */ \
    (-13.0 - 7.0 * WT_S6) / 3.0, (-13.0 + 7.0 * WT_S6) / 3.0, -1.0 / 3.0,        /* 2 E */     \
    3.6378342527444957, 2.6810828736277523, -3.0504301992474105,                 /* 5 MU_REAL, MU_COMPLEX re, im */ \
    0.09443876248897524, -0.14125529502095421, 0.03002919410514742,              /* 8 T row 0 */ \
    0.25021312296533332, 0.20412935229379994, -0.38294211275726192,              /* 11 T row 1 */ \
    4.17871859155190428, 0.32768282076106237, 0.52337644549944951,               /* 14 TI row 0 */ \
    -4.17871859155190428, -0.32768282076106237, 0.47662355450055044,             /* 17 TI row 1 */ \
    0.50287263494578682, -2.57192694985560522, 0.59603920482822492,              /* 20 TI row 2 */ \
    13.0 / 3.0 + 7.0 * WT_S6 / 3.0, -23.0 / 3.0 - 22.0 * WT_S6 / 3.0, 10.0 / 3.0 + 5.0 * WT_S6,   /* 23 P row 0 */ \
    13.0 / 3.0 - 7.0 * WT_S6 / 3.0, -23.0 / 3.0 + 22.0 * WT_S6 / 3.0, 10.0 / 3.0 - 5.0 * WT_S6,   /* 26 P row 1 */ \
    1.0 / 3.0, -8.0 / 3.0, 10.0 / 3.0,                                           /* 29 P row 2 */ \
    1.0, 1.0, 0.0 }                                                              /* 32 T row 2 (stage-indexed reads of T: 8 + 3 i for i < 2, 32 for i = 2) */
#ifdef WT_EMU
static const double wt_rk[35] = WT_RADAU_TABLE;
#else
__constant__ double wt_rk[35] = WT_RADAU_TABLE;
#endif
#define WT_C0 wt_rk[0]
#define WT_C1 wt_rk[1]
#define WT_E0 wt_rk[2]
#define WT_E1 wt_rk[3]
#define WT_E2 wt_rk[4]
#define WT_MU_REAL wt_rk[5]   // 3 + 3**(2/3) - 3**(1/3)
#define WT_MU_CRE wt_rk[6]    // 3 + 0.5*(3**(1/3) - 3**(2/3))
#define WT_MU_CIM wt_rk[7]    // -0.5*(3**(5/6) + 3**(7/6))
#define WT_T00 wt_rk[8]
#define WT_T01 wt_rk[9]
#define WT_T02 wt_rk[10]
#define WT_T10 wt_rk[11]
#define WT_T11 wt_rk[12]
#define WT_T12 wt_rk[13]
#define WT_TI00 wt_rk[14]
#define WT_TI01 wt_rk[15]
#define WT_TI02 wt_rk[16]
#define WT_TI10 wt_rk[17]
#define WT_TI11 wt_rk[18]
#define WT_TI12 wt_rk[19]
#define WT_TI20 wt_rk[20]
#define WT_TI21 wt_rk[21]
#define WT_TI22 wt_rk[22]
#define WT_P00 wt_rk[23]
#define WT_P01 wt_rk[24]
#define WT_P02 wt_rk[25]
#define WT_P10 wt_rk[26]
#define WT_P11 wt_rk[27]
#define WT_P12 wt_rk[28]
#define WT_P20 wt_rk[29]
#define WT_P21 wt_rk[30]
#define WT_P22 wt_rk[31]
#define WT_NEWTON_MAXITER 6
#define WT_HARD_MAX_ATTEMPTS 2000000
#define WT_NEWTON_TOL 1e-3  // max(10*EPS/rtol, min(0.03, sqrt(rtol))), radau.py:315

// ----------------------------------------------------------------------------------------
// lane <-> (plant, zone) geometry of one warp
// ----------------------------------------------------------------------------------------
struct WtGroup {
  int n;       // zones per plant (warp-uniform)
  vi z;        // zone index of this lane inside its plant
  vi base;     // first lane of this lane's plant
  vi gmask;    // bit mask of the plant's lanes
  vb first;    // z == 0
  vb last;     // z == n-1
  int L;       // PCR levels = ceil(log2 n)
  double inv_sqrtN, inv_sqrt3N;  // 1/sqrt(3n), 1/sqrt(9n) for the RMS norms (common.py:63-65)
  vi lane;     // this lane
  vi last_lane;  // last lane of this lane's plant
  vi src_dn1, src_up1;  // neighbour zones' lanes, clamped to the plant
};

// inv_sqrtN / inv_sqrt3N = 1/sqrt(3n), 1/sqrt(9n): warp-uniform, computed by the host (a double sqrt and a
// division with their slow paths are ~100 instructions of once-per-step code in a kernel bound by instruction fetch)
WT_DEV WtGroup wt_make_group(int n, double inv_sqrtN, double inv_sqrt3N) {
  WtGroup g;
  g.n = n;
  vi lane = lane_id();
  int gpw = WT_WARP / n;
  // lanes past the last whole plant form harmless one-lane pseudo groups
  vb in = lane < gpw * n;
  vi q = vdivi(lane, n);
  g.base = seli(in, q * n, lane);
  g.z = seli(in, lane - g.base, 0);
  uint32_t full = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  g.gmask = seli(in, vshl_i((int)full, g.base), vlanebit());
  g.first = g.z == 0;
  g.last = selb(in, g.z == (n - 1), vbroadcast_b(true));
  g.L = 0;
  for (int s = 1; s < n; s <<= 1) ++g.L;
  g.inv_sqrtN = inv_sqrtN;
  g.inv_sqrt3N = inv_sqrt3N;
  g.lane = lane;
  g.last_lane = seli(in, g.base + (n - 1), lane);
  g.src_dn1 = vmaxi(lane - 1, g.base);
  g.src_up1 = vmini(lane + 1, g.last_lane);
  return g;
}
WT_DEV WtGroup wt_make_group(int n) {
  return wt_make_group(n, 1.0 / sqrt((double)(3 * n)), 1.0 / sqrt((double)(9 * n)));
}

// neighbour access inside the plant; `dflt` outside
WT_DEV vd wt_dn(const WtGroup &g, vd x, double dflt) { return sel(g.first, dflt, shfl_up(x, 1)); }
WT_DEV vd wt_up(const WtGroup &g, vd x, double dflt) { return sel(g.last, dflt, shfl_down(x, 1)); }
// Same neighbours through source lanes clamped to the plant (edge zones read themselves): for use
// where the coefficient that multiplies a non-existent neighbour is exactly 0, so no select is needed
// and no value of another plant (possibly NaN) is ever touched.
WT_DEV vd wt_dnc(const WtGroup &g, vd x) { return shfl_idx(x, g.src_dn1); }
WT_DEV vd wt_upc(const WtGroup &g, vd x) { return shfl_idx(x, g.src_up1); }
WT_DEV vd wt_dn_s(const WtGroup &g, vd x, int s) { return sel(g.z >= s, shfl_up(x, s), 0.0); }
WT_DEV vd wt_up_s(const WtGroup &g, vd x, int s) { return sel((g.z + s) < g.n, shfl_down(x, s), 0.0); }
// Source lanes at stride s CLAMPED to the plant: a lane never reads another plant's values (no
// foreign NaNs), and in the PCR recurrences the coefficient multiplying a clamped (i.e.
// non-existent) neighbour is exactly zero, so no select is needed on the fetched value.
WT_DEV vi wt_src_dn(const WtGroup &g, int s) { return vmaxi(g.lane - s, g.base); }
WT_DEV vi wt_src_up(const WtGroup &g, int s) { return vmini(g.lane + s, g.last_lane); }

// sum over the plant's lanes, result replicated on all of them
WT_DEV vd wt_gsum(const WtGroup &g, vd x) {
  WT_NOUNROLL
  for (int s = 1; s < g.n; s <<= 1) x = x + wt_up_s(g, x, s);
  return shfl_idx(x, g.base);
}
WT_DEV vb wt_gany(const WtGroup &g, vb c) { return vmask_any(vballot(c), g.gmask); }

// ----------------------------------------------------------------------------------------
// per-plant constants and boundary, replicated per lane
// ----------------------------------------------------------------------------------------
// The 17 numeric constants live in the per-warp store (shared memory on the GPU: one copy
// per plant, read as a broadcast) instead of 34 registers per lane; the flags stay in registers.
enum {
  CK_Kw = 0, CK_Ka1, CK_Ka12, CK_KaCl, CK_CT2303, CK_Kx, CK_zh, CK_Ri_thr, CK_QV, CK_Hin, CK_dHd,
  CK_cl_dose, CK_inCl, CK_inT, CK_hlA, CK_amb, CK_inv_hl_den, CK_flow, CK_N
};
// per-LANE constants (zone-position dependent), kept in lane-private slots of the store after the LU
// multipliers: the "zone 0 only" / "last zone only" terms of reactor.py:336-395 become multiplications
// by a constant that is 0 on the other zones (x + 0 and 0 * finite are exact), not selects
enum { LK_KX_UP = 0, LK_KX_DN, LK_QV_FIRST, LK_QV_LAST, LK_DHD_FIRST, LK_DOSE_FIRST, LK_N };
template <class Store>
struct WtConstT {
  Store *st;
  int lk0;  // first lane-constant slot
  vb strat, v_ok;
  WT_DEV vd lk(int k) const { return st->get(lk0 + k); }
  WT_DEV vd Kw() const { return st->cget(CK_Kw); }
  WT_DEV vd Ka1() const { return st->cget(CK_Ka1); }
  WT_DEV vd Ka12() const { return st->cget(CK_Ka12); }
  WT_DEV vd KaCl() const { return st->cget(CK_KaCl); }
  WT_DEV vd CT2303() const { return st->cget(CK_CT2303); }
  WT_DEV vd Kx() const { return st->cget(CK_Kx); }
  WT_DEV vd zh() const { return st->cget(CK_zh); }
  WT_DEV vd Ri_thr() const { return st->cget(CK_Ri_thr); }
  WT_DEV vd QV() const { return st->cget(CK_QV); }
  WT_DEV vd Hin() const { return st->cget(CK_Hin); }
  WT_DEV vd dHd() const { return st->cget(CK_dHd); }
  WT_DEV vd cl_dose() const { return st->cget(CK_cl_dose); }
  WT_DEV vd inCl() const { return st->cget(CK_inCl); }
  WT_DEV vd inT() const { return st->cget(CK_inT); }
  WT_DEV vd hlA() const { return st->cget(CK_hlA); }
  WT_DEV vd amb() const { return st->cget(CK_amb); }
  WT_DEV vd inv_hl_den() const { return st->cget(CK_inv_hl_den); }
};

// par / bnd hold this lane's plant values (already loaded); boundary-derived terms follow
// reactor.py:336, 349-368, 388-395, 420, 426-443
template <class Store>
WT_DEV WtConstT<Store> wt_make_const(Store *st, const WtGroup &g, int lk0, const vd *par, const vd *bnd) {
  WtConstT<Store> c;
  c.st = st;
  c.lk0 = lk0;
  st->cput(CK_Kw, par[WTP_KW]);
  st->cput(CK_Ka1, par[WTP_KA1]);
  st->cput(CK_Ka12, par[WTP_KA1] * par[WTP_KA2]);
  st->cput(CK_KaCl, par[WTP_KACL]);
  st->cput(CK_CT2303, 2.303 * par[WTP_CT]);
  st->cput(CK_Kx, par[WTP_KX]);
  st->cput(CK_zh, par[WTP_ZH]);
  st->cput(CK_Ri_thr, 0.25 * (par[WTP_V] * par[WTP_V]));
  c.v_ok = par[WTP_V] > 1e-6;
  c.strat = par[WTP_STRAT] != 0.0;
  const vd QV = wt_div(wt_div(bnd[WTB_INLET_FLOW], 60.0), par[WTP_VOLUME]);
  st->cput(CK_QV, QV);
  st->cput(CK_Hin, vexp10(-bnd[WTB_INLET_PH]));
  const vd dHd = sel(bnd[WTB_ACID_FLOW] > 0.0, wt_div(wt_div(bnd[WTB_ACID_FLOW], 60.0) * bnd[WTB_ACID_CONC], par[WTP_VZL]), 0.0);
  st->cput(CK_dHd, dHd);
  const vd dose = sel(bnd[WTB_CL_FLOW] > 0.0, wt_div(wt_div(bnd[WTB_CL_FLOW], 60.0) * bnd[WTB_CL_CONC], par[WTP_VZL]), 0.0);
  st->cput(CK_cl_dose, dose);
  st->cput(CK_inCl, bnd[WTB_INLET_CL]);
  st->cput(CK_inT, bnd[WTB_INLET_T]);
  st->cput(CK_hlA, sel(bnd[WTB_HEAT_LOSS] > 0.0, bnd[WTB_HEAT_LOSS] * par[WTP_AT], 0.0));  // reactor.py:426: only if > 0
  const vb all = vbroadcast_b(true);
  st->put(lk0 + LK_KX_UP, sel(g.last, 0.0, par[WTP_KX]), all);
  st->put(lk0 + LK_KX_DN, sel(g.first, 0.0, par[WTP_KX]), all);
  st->put(lk0 + LK_QV_FIRST, sel(g.first, QV, 0.0), all);
  st->put(lk0 + LK_QV_LAST, sel(g.last, QV, 0.0), all);
  st->put(lk0 + LK_DHD_FIRST, sel(g.first, dHd, 0.0), all);
  st->put(lk0 + LK_DOSE_FIRST, sel(g.first, dose, 0.0), all);
  st->cput(CK_amb, bnd[WTB_AMBIENT_T]);
  st->cput(CK_inv_hl_den, wt_rcp((998.2 * 4184.0) * wt_div(par[WTP_VOLUME], 1000.0)));
  st->csync();
  return c;
}

// ----------------------------------------------------------------------------------------
// right-hand-side building blocks
// ----------------------------------------------------------------------------------------

// spatial.py:175-195
WT_DEV vd wt_density(vd T) {
  vd d4 = T - 4.0;
  vd cold = WT_PC_RHO_A + WT_PC_RHO_B * (d4 * d4);
  vd warm = WT_PC_RHO_W + WT_PC_RHO_WS * (T - 20.0);
  return sel(T <= 8.0, cold, warm);
}

// spatial.py:266-275, 293, 313-316: exchange multiplier of the interface between a zone
// (rho_lo) and the zone above it (rho_hi)
template <class Store>
WT_DEV vd wt_suppression(const WtConstT<Store> &c, vd rho_lo, vd rho_hi) {
  vd drho = rho_hi - rho_lo;
  vd ravg = 0.5 * (rho_lo + rho_hi);
  // Ri = (g drho zh)/(ravg v^2) > 0.25, cross-multiplied (ravg v^2 > 0): same decision except
  // within an ulp of the threshold; v <= 1e-6 -> Ri = +inf
  vb stable = (!c.v_ok) | (((WT_PC_G * drho) * c.zh()) > ravg * c.Ri_thr());
  return sel(c.strat & stable, 0.5, 1.0);
}

// thermodynamics.py:187-193
WT_DEV vd wt_arrhenius(vd T) {
  vd TK = T + WT_PC_T0K;
  vd e = WT_PC_EA_R * (wt_rcp(TK) - WT_PC_ITREF);
  return WT_PC_KREF * vexp(e);
}
WT_DEV vb wt_t_out_of_range(vd T) { return (T < 0.0) | (T > 100.0); }  // thermodynamics.py:146-157

// chemistry.py:422-437 (+ :181-191): beta(pH) * ln(10).  Split into "denominators", "reciprocals" and "the rest"
// so that the RHS can take the three reciprocals of a zone (1/H, 1/D, 1/(H + Ka_HOCl)) in one interleaved
// wt_rcp_n<3> while num_jac, which needs them one at a time, runs the very same arithmetic.
template <class Store>
WT_DEV vd wt_beta_den(const WtConstT<Store> &c, vd H) { return H * H + c.Ka1() * H + c.Ka12(); }
template <class Store>
WT_DEV vd wt_beta_ln10_r(const WtConstT<Store> &c, vd H, vd iH, vd iD, vb &bpos) {
  vd bw = WT_PC_2303 * (H + c.Kw() * iH);
  vd HH = H * H;
  // the three alphas share one reciprocal (<= 1 ulp from three divisions)
  vd a0 = HH * iD;
  vd a1 = (c.Ka1() * H) * iD;
  vd a2 = c.Ka12() * iD;
  vd bc = c.CT2303() * (a0 * a1 + (4.0 * a1) * a2 + a0 * a2);
  vd beta = bw + bc;
  bpos = beta > 0.0;
  return beta * WT_PC_LN10;
}
template <class Store>
WT_DEV vd wt_beta_ln10(const WtConstT<Store> &c, vd H, vb &bpos) {
  return wt_beta_ln10_r(c, H, wt_rcp(H), wt_rcp(wt_beta_den(c, H)), bpos);
}

// chemistry.py:510-523
template <class Store>
WT_DEV vd wt_decay_factor_r(const WtConstT<Store> &c, vd H, vd iden) {
  return (H * iden) * 1.0 + (c.KaCl() * iden) * WT_PC_OCL;
}
template <class Store>
WT_DEV vd wt_decay_factor(const WtConstT<Store> &c, vd H) { return wt_decay_factor_r(c, H, wt_rcp(H + c.KaCl())); }

struct WtMix { vd off_dn, off_up, diag; };

// reactor.py:318-337: row z of the stratification-scaled exchange matrix
template <class Store>
WT_DEV WtMix wt_mix_row(const WtGroup &g, const WtConstT<Store> &c, vd s_dn, vd s_up) {
  WtMix m;
  m.off_up = c.lk(LK_KX_UP) * s_up;
  m.off_dn = c.lk(LK_KX_DN) * s_dn;
  m.diag = -(m.off_dn + m.off_up) - c.lk(LK_QV_LAST);
  return m;
}
WT_DEV vd wt_mix(const WtMix &m, vd xdn, vd x, vd xup) { return (m.off_dn * xdn + m.diag * x) + m.off_up * xup; }

// reactor.py:349-368: the zone-0-only acid dosing + inlet terms of dpH (0 elsewhere)
// (ibl = 1 / (beta ln10): the up-to-three quotients of zone 0 share one reciprocal)
template <class Store>
// The reference guards each of the three terms with the same `beta > 0`; the guard is applied once,
// to the sum, in wt_dph.
WT_DEV vd wt_dph_inlet(const WtGroup &g, const WtConstT<Store> &c, vd H, vd ibl, vb bpos) {
  vd t1 = -c.lk(LK_DHD_FIRST) * ibl;
  vd dHin = c.lk(LK_QV_FIRST) * (c.Hin() - H);
  vd t2 = -dHin * ibl;
  return (0.0 + t1) + t2;
}
// reactor.py:371-376
WT_DEV vd wt_dph(vd t12, vd mixH, vd ibl, vb bpos) { return sel(bpos, t12 + (-mixH * ibl), 0.0); }
// reactor.py:388-411
template <class Store>
WT_DEV vd wt_dcl(const WtGroup &g, const WtConstT<Store> &c, vd Cl, vd mixCl, vd kf) {
  vd r = c.lk(LK_DOSE_FIRST) + c.lk(LK_QV_FIRST) * (c.inCl() - Cl);
  r = r + mixCl;
  return r - kf * Cl;
}
// reactor.py:420-443
template <class Store>
WT_DEV vd wt_dt(const WtGroup &g, const WtConstT<Store> &c, vd T, vd mixT) {
  vd r = c.lk(LK_QV_FIRST) * (c.inT() - T) + mixT;
  return r - (c.hlA() * (T - c.amb())) * c.inv_hl_den();  // hlA = 0 when the coefficient is not > 0
}

// Full RHS for this lane's zone.  `bad` is set where the reference would raise ValueError.
template <class Store>
WT_DEV void wt_rhs(const WtGroup &g, const WtConstT<Store> &c, vd pH, vd Cl, vd T, vd &dpH, vd &dCl, vd &dT,
                   vb &bad) {
  vd rho = wt_density(T);
  vd s_up = wt_suppression(c, rho, shfl_down(rho, 1));
  vd s_dn = shfl_up(s_up, 1);
  WtMix m = wt_mix_row(g, c, s_dn, s_up);
  vd H, ke;
  wt_h_and_arrh(pH, T, H, ke);  // H = 10^-pH and the Arrhenius exponential, chains interleaved
  vd den[3] = {H, wt_beta_den(c, H), H + c.KaCl()}, inv[3];
  wt_rcp_n<3>(den, inv);  // three Newton chains interleaved
  vb bpos;
  vd ibl = wt_rcp(wt_beta_ln10_r(c, H, inv[0], inv[1], bpos));
  vd mixH = wt_mix(m, wt_dnc(g, H), H, wt_upc(g, H));
  dpH = wt_dph(wt_dph_inlet(g, c, H, ibl, bpos), mixH, ibl, bpos);
  vd kf = (WT_PC_KREF * ke) * wt_decay_factor_r(c, H, inv[2]);
  dCl = wt_dcl(g, c, Cl, wt_mix(m, wt_dnc(g, Cl), Cl, wt_upc(g, Cl)), kf);
  dT = wt_dt(g, c, T, wt_mix(m, wt_dnc(g, T), T, wt_upc(g, T)));
  bad = wt_t_out_of_range(T);
}

// ----------------------------------------------------------------------------------------
// tridiagonal systems across lanes: parallel cyclic reduction
// ----------------------------------------------------------------------------------------
// LuStore (backend-specific, see the kernel / the emulation harness) keeps, per lane, the elimination multipliers
// of each level and the final reciprocal pivot, in two slot spaces:
//   real     void put(int slot, vd x, vb mask);            vd get(int slot);
//   complex  void cx_put4(int slot, const vd *x, vb mask);  void cx_get4(int slot, vd *x);   (4 consecutive slots)
// plus begin_factor(keep) / end_factor() around a factorization (`keep` = lanes whose stored factors must survive it).
// On the GPU the real slots live in shared memory and the complex ones in TENSOR MEMORY (one tcgen05.ld.x8 fetches
// the four doubles of a level), which takes the store out of the shared-memory budget that capped the occupancy.
//
// Rows: a x[z-1] + b x[z] + c x[z+1] = d, with a = 0 on the first and c = 0 on the last zone.
// At the level of stride s the multipliers that would touch a neighbour outside the plant are
// exactly 0 (a stays 0 on zones < s, c on zones >= n-s), so fetched values need no masking.
// The LAST level (stride s = 2^(L-1), 2 s >= n) couples every zone with at most ONE partner (z - s or z + s),
// so it keeps one multiplier and one exchange per zone instead of two (same bits: the other product was an exact 0).
// Slots per system: real 2 L  = {k1, k2} x (L-1), k, pivot;  complex 4 L = {k1r, k1i, k2r, k2i} x (L-1), kr, ki, pr, pi.

WT_DEV int wt_pcr_levels(int n) { int L = 0; for (int s = 1; s < n; s <<= 1) ++L; return L; }
// partner lane of the last PCR level (clamped to the plant: a zone without partner has a zero multiplier)
WT_DEV vi wt_src_last(const WtGroup &g, int s) { return seli(g.z >= s, g.lane - s, vmini(g.lane + s, g.last_lane)); }

// ----------------------------------------------------------------------------------------
// finite-difference Jacobian rows of this lane's zone (index 0: column zone z-1, 1: z, 2: z+1)
// ----------------------------------------------------------------------------------------
// The diagonal blocks d(dT)/dT, d(dpH)/dpH, d(dCl)/dCl (3 x 3 doubles per lane) are only needed by the
// factorization; num_jac parks them in the store (WtPlantStep::PK_JD).  The coupling blocks are used by every solve:
struct WtJac {
  vd pt[3];  // d(dpH_z)/dT   (non-zero only when a perturbation flips a Richardson switch)
  vd ct[3];  // d(dCl_z)/dT
  vd cp;     // d(dCl_z)/dpH_z
};

// One plant-step worth of solver state, per lane.
template <class LuStore>
struct WtPlantStep {
  WtGroup g;
  WtConstT<LuStore> c;
  LuStore *lu;
  // state vector of this zone: 0 pH, 1 Cl, 2 T
  vd y[3];
  WtJac J;   // tt / pp / cc are only filled inside num_jac and factor (parked in between), pt / ct / cp stay in registers
  // PARKED per-lane state: live for the whole step but touched once or twice per attempt, never inside the Newton
  // loop.  It lives in lane-private slots of the store (shared memory), not in 54 registers per lane: with it parked
  // the hot loops fit the 168 registers of three resident blocks per SM.
  //   f = dy/dt at (t, y)           3      jfac (num_jac step factors)    3
  //   Q dense output (radau.py:547) 9      yold                           3      J.tt / J.pp / J.cc   9
  enum { PK_F = 0, PK_JFAC = 3, PK_Q = 6, PK_YOLD = 15, PK_JD = 18, PK_N = 27 };
  int pk0;  // first parking slot
  WT_DEV vd pk(int k) const { return lu->get(pk0 + k); }
  WT_DEV void pkset(int k, vd x, vb m) { lu->put(pk0 + k, x, m); }
  // Per-plant decisions of the solver, one bit each in ONE register (as separate bools ptxas kept them in
  // byte lanes of several registers and spilled those to local memory, which misses L1 here: long_sb stalls)
  enum { F_RUNNING = 1, F_NEED_JAC = 2, F_CURRENT_JAC = 4, F_LU_VALID = 8, F_NEW_STEP = 16, F_HAVE_OLD = 32,
         F_REJECTED = 64, F_HAVE_JFAC = 128, F_HAVE_SOL = 256, F_TRANGE = 512, F_FAILED = 1024,
         F_WORKLIMIT = 2048, F_SELF_HAVE_OLD = 4096, F_DEGRADED = 8192 };
  vi fl;
  WT_DEV vb fget(int bit) const { return (fl & bit) != 0; }
  WT_DEV void fset(int bit, vb m) { fl = fl | seli(m, bit, 0); }      // set where m
  WT_DEV void fclr(int bit, vb m) { fl = fl & seli(m, ~bit, -1); }    // clear where m
  WT_DEV vb trange() const { return fget(F_TRANGE); }      // the reference would have raised ValueError inside the solve
  WT_DEV vb failed() const { return fget(F_FAILED); }      // TOO_SMALL_STEP
  WT_DEV vb degraded() const { return fget(F_DEGRADED); }    // engine policy: floor mode forced an acceptance (not reference behaviour)
  WT_DEV vb worklimit() const { return fget(F_WORKLIMIT); }  // engine policy: attempt budget exhausted (not reference behaviour)
  vd W[3][3];     // W[k][var]
  // Per-plant step-control scalars live in the per-plant store (shared memory), not replicated in two
  // registers per lane for the whole step: they are touched a few times per attempt, never in the Newton loop.
  enum { PV_SELF_H = 0, PV_SELF_H_OLD, PV_SELF_ERR_OLD, PV_H_OLD, PV_ERR_OLD, PV_MIN_STEP, PV_SOL_TOLD, PV_SOL_H,
         PV_T_BOUND, PV_MAX_STEP, PV_N };
  WT_DEV vd pv(int k) const { return lu->pvget(k); }
  WT_DEV void pvset(int k, vd x) { lu->pvput(k, x); }                          // every lane of the plant stores the same value
  WT_DEV void pvset(int k, vd x, vb m) { lu->pvput(k, sel(m, x, lu->pvget(k))); }  // ... where m

  WT_DEV int slot_real(int sys) const { return sys * (2 * g.L); }   // real slot space (3 * 2 L slots)
  WT_DEV int slot_cplx(int sys) const { return sys * (4 * g.L); }   // complex slot space (3 * 4 L slots)

  // -------------------------------------------------------------------------------------
  // linear algebra on the block-triangular structure.  System order: 0 = T, 1 = pH, 2 = Cl.
  // -------------------------------------------------------------------------------------
  // All six factorizations (real + complex of T, pH, Cl) share one sweep over the PCR levels:
  // six independent dependency chains per level instead of six sweeps back to back.
  // The three systems (T, pH, Cl) go through ONE rolled copy of the factorization (real and complex chains of a
  // system interleaved): a third of the code of the all-systems-at-once version, which matters more than its
  // extra instruction-level parallelism in a kernel bound by instruction fetch.
  WT_DEV void factor(vd h, vb mask, vb keep) {
    lu->begin_factor(keep);
    const vd ih = wt_rcp(h);
    const vd mr = WT_MU_REAL * ih, gr = WT_MU_CRE * ih, gi = WT_MU_CIM * ih;
    const int s_last = 1 << (g.L - 1);
    WT_NOUNROLL
    for (int q = 0; q < 3; ++q) {
      // rows of (mu/h I - J_qq): J.tt, J.pp, J.cc parked by num_jac
      const vd jdg = pk(PK_JD + 3 * q + 1);
      vd a = -pk(PK_JD + 3 * q), b = mr - jdg, c_ = -pk(PK_JD + 3 * q + 2);                 // real
      vd ar = a, ai = vbroadcast(0.0), br = gr - jdg, bi = gi, cr = c_, ci = vbroadcast(0.0);  // complex
      const int sr = slot_real(q), sc = slot_cplx(q);
      int l = 0;
      WT_NOUNROLL
      for (int s = 1; s < s_last; s <<= 1, ++l) {
        const vi sd = wt_src_dn(g, s), su = wt_src_up(g, s);
        vd den[2] = {b, br * br + bi * bi}, inv[2];  // real pivot, |complex pivot|^2
        wt_rcp_n<2>(den, inv);
        const vd rr = br * inv[1], ri = -(bi * inv[1]);  // 1 / complex pivot
        const vd k1 = a * shfl_idx(inv[0], sd), k2 = c_ * shfl_idx(inv[0], su);
        const vd rdr = shfl_idx(rr, sd), rdi = shfl_idx(ri, sd), rur = shfl_idx(rr, su), rui = shfl_idx(ri, su);
        vd kc[4];  // k1r, k1i, k2r, k2i
        kc[0] = ar * rdr - ai * rdi; kc[1] = ar * rdi + ai * rdr;
        kc[2] = cr * rur - ci * rui; kc[3] = cr * rui + ci * rur;
        const vd a_dn = shfl_idx(a, sd), c_dn = shfl_idx(c_, sd), a_up = shfl_idx(a, su), c_up = shfl_idx(c_, su);
        b = b - c_dn * k1 - a_up * k2;
        a = -(a_dn * k1);
        c_ = -(c_up * k2);
        const vd adr = shfl_idx(ar, sd), adi = shfl_idx(ai, sd), cdr = shfl_idx(cr, sd), cdi = shfl_idx(ci, sd);
        const vd aur = shfl_idx(ar, su), aui = shfl_idx(ai, su), cur = shfl_idx(cr, su), cui = shfl_idx(ci, su);
        br = wt_cmsub_re(wt_cmsub_re(br, cdr, cdi, kc[0], kc[1]), aur, aui, kc[2], kc[3]);
        bi = wt_cmsub_im(wt_cmsub_im(bi, cdr, cdi, kc[0], kc[1]), aur, aui, kc[2], kc[3]);
        ar = -(adr * kc[0] - adi * kc[1]);
        ai = -(adr * kc[1] + adi * kc[0]);
        cr = -(cur * kc[2] - cui * kc[3]);
        ci = -(cur * kc[3] + cui * kc[2]);
        lu->put(sr + 2 * l, k1, mask);
        lu->put(sr + 2 * l + 1, k2, mask);
        lu->cx_put4(sc + 4 * l, kc, mask);
      }
      {
        // last level: one partner per zone (z - s for the upper zones, z + s for the lower ones), then the pivots
        const vb upper = g.z >= s_last;
        const vi sl = wt_src_last(g, s_last);
        vd den[2] = {b, br * br + bi * bi}, inv[2];
        wt_rcp_n<2>(den, inv);
        const vd rr = br * inv[1], ri = -(bi * inv[1]);
        const vd e = sel(upper, a, c_), er = sel(upper, ar, cr), ei = sel(upper, ai, ci);
        const vd k = e * shfl_idx(inv[0], sl);
        const vd rpr = shfl_idx(rr, sl), rpi = shfl_idx(ri, sl);
        vd kc[4];  // kr, ki, pivot re, pivot im
        kc[0] = er * rpr - ei * rpi; kc[1] = er * rpi + ei * rpr;
        const vd ep = shfl_idx(e, sl), epr = shfl_idx(er, sl), epi = shfl_idx(ei, sl);
        b = b - ep * k;
        br = wt_cmsub_re(br, epr, epi, kc[0], kc[1]);
        bi = wt_cmsub_im(bi, epr, epi, kc[0], kc[1]);
        den[0] = b; den[1] = br * br + bi * bi;
        wt_rcp_n<2>(den, inv);
        kc[2] = br * inv[1];
        kc[3] = -(bi * inv[1]);
        lu->put(sr + 2 * l, k, mask);
        lu->put(sr + 2 * l + 1, inv[0], mask);
        lu->cx_put4(sc + 4 * l, kc, mask);
      }
    }
    lu->end_factor();
  }
  WT_DEV vd tri_mv(const vd *row, vd xdn, vd x, vd xup) const { return (row[0] * xdn + row[1] * x) + row[2] * xup; }
  // one system, real and complex right-hand sides in the same sweep (three independent chains); `cplx` false
  // (warp-uniform): the real one only (a closing pass)
  WT_DEV void solve_sys3(int q, vd &d, vd &dr, vd &di, bool cplx) {
    const int sr = slot_real(q), sc = slot_cplx(q);
    int l = 0;
    const int s_last = 1 << (g.L - 1);
    WT_NOUNROLL
    for (int s = 1; s < s_last; s <<= 1, ++l) {
      const vi sd = wt_src_dn(g, s), su = wt_src_up(g, s);
      vd k1 = lu->get(sr + 2 * l), k2 = lu->get(sr + 2 * l + 1);
      vd dd = shfl_idx(d, sd), du = shfl_idx(d, su);
      d = d - dd * k1 - du * k2;
      if (cplx) {
        vd kc[4];  // k1r, k1i, k2r, k2i
        lu->cx_get4(sc + 4 * l, kc);
        vd ddr = shfl_idx(dr, sd), ddi = shfl_idx(di, sd), dur = shfl_idx(dr, su), dui = shfl_idx(di, su);
        vd nr = wt_cmsub_re(wt_cmsub_re(dr, ddr, ddi, kc[0], kc[1]), dur, dui, kc[2], kc[3]);
        vd ni = wt_cmsub_im(wt_cmsub_im(di, ddr, ddi, kc[0], kc[1]), dur, dui, kc[2], kc[3]);
        dr = nr;
        di = ni;
      }
    }
    {
      const vi sl = wt_src_last(g, s_last);
      vd k = lu->get(sr + 2 * l), piv = lu->get(sr + 2 * l + 1);
      vd dp = shfl_idx(d, sl);
      d = (d - dp * k) * piv;
      if (cplx) {
        vd kc[4];  // kr, ki, pivot re, pivot im
        lu->cx_get4(sc + 4 * l, kc);
        vd dpr = shfl_idx(dr, sl), dpi = shfl_idx(di, sl);
        vd nr = wt_cmsub_re(dr, dpr, dpi, kc[0], kc[1]);
        vd ni = wt_cmsub_im(di, dpr, dpi, kc[0], kc[1]);
        dr = nr * kc[2] - ni * kc[3];
        di = nr * kc[3] + ni * kc[2];
      }
    }
  }
  // the real and the complex collocation systems of one Newton iteration together
  WT_DEV void solve_newton(vd *b, vd *br, vd *bi, bool cplx) {
    vd xT = b[2], tr = br[2], ti = bi[2];
    solve_sys3(0, xT, tr, ti, cplx);
    vd xTd = wt_dnc(g, xT), xTu = wt_upc(g, xT);
    vd xp = b[0] + tri_mv(J.pt, xTd, xT, xTu);
    vd pr = br[0], pi = bi[0];
    vd trd = tr, tru = tr, tid = ti, tiu = ti;
    if (cplx) {
      trd = wt_dnc(g, tr); tru = wt_upc(g, tr); tid = wt_dnc(g, ti); tiu = wt_upc(g, ti);
      pr = pr + tri_mv(J.pt, trd, tr, tru);
      pi = pi + tri_mv(J.pt, tid, ti, tiu);
    }
    solve_sys3(1, xp, pr, pi, cplx);
    vd xc = b[1] + tri_mv(J.ct, xTd, xT, xTu) + J.cp * xp;
    vd qr = br[1], qi = bi[1];
    if (cplx) {
      qr = qr + tri_mv(J.ct, trd, tr, tru) + J.cp * pr;
      qi = qi + tri_mv(J.ct, tid, ti, tiu) + J.cp * pi;
    }
    solve_sys3(2, xc, qr, qi, cplx);
    b[0] = xp; b[1] = xc; b[2] = xT;
    br[0] = pr; bi[0] = pi; br[1] = qr; bi[1] = qi; br[2] = tr; bi[2] = ti;
  }

  // common.py:63-65 over the plant's 3n unknowns
  WT_DEV vd rms3(vd a, vd b, vd cc_) const {
    vd s = wt_gsum(g, (a * a + b * b) + cc_ * cc_);
    return vsqrt(s) * g.inv_sqrtN;
  }

  // -------------------------------------------------------------------------------------
  // num_jac (common.py:311-382) at (y, f), restricted to the plants in `m`
  // -------------------------------------------------------------------------------------
  WT_DEV void num_jac(vb m) {
    const double REJECT = 2.0097183471152322e-14;  // EPS ** 0.875
    const double SMALL = 1.8189894035458565e-12;   // EPS ** 0.75
    const double BIG = 1.220703125e-4;             // EPS ** 0.25
    const double MINF = 1e3 * WT_EPS;
    lu->cadd(WTC_NJEV, seli(m, 1, 0));

    vd f[3], jfac[3];
    WT_UNROLL
    for (int v = 0; v < 3; ++v) {
      f[v] = pk(PK_F + v);
      jfac[v] = sel(fget(F_HAVE_JFAC), pk(PK_JFAC + v), 1.4901161193847656e-08);  // EPS ** 0.5 on the first call of a step
    }
    fset(F_HAVE_JFAC, m);

    // ---- base intermediates at y (identical bits to the evaluation that produced f)
    const vd pH = y[0], Cl = y[1], T = y[2];
    vd rho = wt_density(T);
    vd rho_up = shfl_down(rho, 1), rho_dn = shfl_up(rho, 1);
    vd s_up = wt_suppression(c, rho, rho_up);
    vd s_dn = shfl_up(s_up, 1);
    WtMix mx = wt_mix_row(g, c, s_dn, s_up);
    vd H = vexp10(-pH);
    vb bpos;
    vd bl = wt_rcp(wt_beta_ln10(c, H, bpos));  // 1 / (beta ln10)
    vd t12 = wt_dph_inlet(g, c, H, bl, bpos);
    vd kk = wt_arrhenius(T);
    vd kf = kk * wt_decay_factor(c, H);
    vd Hdn = wt_dnc(g, H), Hup = wt_upc(g, H);
    vd Cldn = wt_dnc(g, Cl), Clup = wt_upc(g, Cl);
    vd Tdn = wt_dnc(g, T), Tup = wt_upc(g, T);
    vd mixCl = wt_mix(mx, Cldn, Cl, Clup);
    // base f of the neighbours (rows of the column's stencil) and of row 0
    vd fdn[3], fup[3];
    WT_UNROLL
    for (int v = 0; v < 3; ++v) { fdn[v] = shfl_up(f[v], 1); fup[v] = shfl_down(f[v], 1); }
    vd f_row0 = vabs(shfl_idx(f[0], g.base));

    // ---- perturbation steps (common.py:330-345)
    vd ysc[3], hh[3];
    WT_UNROLL
    for (int v = 0; v < 3; ++v) {
      vd sgn = sel(f[v] >= 0.0, 1.0, -1.0);
      ysc[v] = sgn * vmax(vabs(y[v]), WT_ATOL);
      hh[v] = (y[v] + jfac[v] * ysc[v]) - y[v];
      vb zero = m & (hh[v] == 0.0);
      while (vany(zero)) {
        jfac[v] = sel(zero, jfac[v] * 10.0, jfac[v]);
        hh[v] = sel(zero, (y[v] + jfac[v] * ysc[v]) - y[v], hh[v]);
        zero = m & (hh[v] == 0.0);
      }
    }

    // results of the accepted evaluation of each column this lane's row sees
    vd d_pp[3], d_pt[3], d_cc[3], d_ct[3], d_tt[3], d_cp = vbroadcast(0.0);  // f_new - f
    WT_UNROLL
    for (int k = 0; k < 3; ++k) {
      d_pp[k] = vbroadcast(0.0); d_pt[k] = vbroadcast(0.0); d_cc[k] = vbroadcast(0.0);
      d_ct[k] = vbroadcast(0.0); d_tt[k] = vbroadcast(0.0);
    }
    vd maxd[3], scl[3];
    vb retry[3];
    WT_UNROLL
    for (int v = 0; v < 3; ++v) { retry[v] = vbroadcast_b(false); maxd[v] = vbroadcast(0.0); scl[v] = vbroadcast(0.0); }

    WT_NOUNROLL  // the retry pass is rare: one copy of the column evaluation (instruction-fetch bound kernel)
    for (int pass = 0; pass < 2; ++pass) {
      vd hc[3];
      if (pass == 1) {
        vb anyretry = m & (retry[0] | retry[1] | retry[2]);
        if (!vany(anyretry)) break;
        lu->cadd(WTC_JAC_RETRY, seli(anyretry, 1, 0));
      }
      WT_UNROLL
      for (int v = 0; v < 3; ++v)
        hc[v] = (pass == 0) ? hh[v] : sel(retry[v], (y[v] + (10.0 * jfac[v]) * ysc[v]) - y[v], hh[v]);

      // own perturbed quantities
      vd pHp = pH + hc[0];
      vd Hp = vexp10(-pHp);
      vb bposp;
      vd blp = wt_rcp(wt_beta_ln10(c, Hp, bposp));
      vd kfp_pH = kk * wt_decay_factor(c, Hp);
      vd Clp = Cl + hc[1];
      vd Tp = T + hc[2];
      fset(F_TRANGE, m & wt_gany(g, wt_t_out_of_range(Tp) | wt_t_out_of_range(T)));
      vd kfp_T = wt_arrhenius(Tp) * wt_decay_factor(c, H);
      vd rhop = wt_density(Tp);

      // ---- new f values of row z under each single-column perturbation
      vd n_pp[3], n_pt[3], n_cc[3], n_ct[3], n_tt[3], n_cp;
      // pH columns
      n_pp[1] = wt_dph(wt_dph_inlet(g, c, Hp, blp, bposp), wt_mix(mx, Hdn, Hp, Hup), blp, bposp);
      n_cp = wt_dcl(g, c, Cl, mixCl, kfp_pH);
      n_pp[0] = wt_dph(t12, wt_mix(mx, wt_dnc(g, Hp), H, Hup), bl, bpos);
      n_pp[2] = wt_dph(t12, wt_mix(mx, Hdn, H, wt_upc(g, Hp)), bl, bpos);
      // Cl columns
      n_cc[1] = wt_dcl(g, c, Clp, wt_mix(mx, Cldn, Clp, Clup), kf);
      n_cc[0] = wt_dcl(g, c, Cl, wt_mix(mx, wt_dnc(g, Clp), Cl, Clup), kf);
      n_cc[2] = wt_dcl(g, c, Cl, wt_mix(mx, Cldn, Cl, wt_upc(g, Clp)), kf);
      // T columns: the perturbed density can flip the Richardson switch of either interface
      {
        WtMix mo = wt_mix_row(g, c, wt_suppression(c, rho_dn, rhop), wt_suppression(c, rhop, rho_up));
        n_pt[1] = wt_dph(t12, wt_mix(mo, Hdn, H, Hup), bl, bpos);
        n_ct[1] = wt_dcl(g, c, Cl, wt_mix(mo, Cldn, Cl, Clup), kfp_T);
        n_tt[1] = wt_dt(g, c, Tp, wt_mix(mo, Tdn, Tp, Tup));
        WtMix ml = wt_mix_row(g, c, wt_suppression(c, shfl_up(rhop, 1), rho), s_up);
        n_pt[0] = wt_dph(t12, wt_mix(ml, Hdn, H, Hup), bl, bpos);
        n_ct[0] = wt_dcl(g, c, Cl, wt_mix(ml, Cldn, Cl, Clup), kf);
        n_tt[0] = wt_dt(g, c, T, wt_mix(ml, wt_dnc(g, Tp), T, Tup));
        WtMix mr = wt_mix_row(g, c, s_dn, wt_suppression(c, rho, shfl_down(rhop, 1)));
        n_pt[2] = wt_dph(t12, wt_mix(mr, Hdn, H, Hup), bl, bpos);
        n_ct[2] = wt_dcl(g, c, Cl, wt_mix(mr, Cldn, Cl, Clup), kf);
        n_tt[2] = wt_dt(g, c, T, wt_mix(mr, Tdn, T, wt_upc(g, Tp)));
      }

      // ---- column owner: arg-max row in species-major order (np.argmax: first maximum)
      vd best[3], bf[3], bn[3];
      WT_UNROLL
      for (int v = 0; v < 3; ++v) { best[v] = vbroadcast(0.0); bf[v] = vbroadcast(0.0); bn[v] = vbroadcast(0.0); }
      vb hasdn = !g.first, hasup = !g.last;
#define WT_CONSIDER(v, fbase, fnew, ok)                                   \
  {                                                                       \
    vd fb_ = (fbase), fn_ = (fnew);                                       \
    vd ad_ = vabs(fn_ - fb_);                                             \
    vb take_ = (ok) & (ad_ > best[v]);                                    \
    best[v] = sel(take_, ad_, best[v]);                                   \
    bf[v] = sel(take_, vabs(fb_), bf[v]);                                 \
    bn[v] = sel(take_, vabs(fn_), bn[v]);                                 \
  }
      vb yes = vbroadcast_b(true);
      // column pH_z: rows pH_{z-1}, pH_z, pH_{z+1}, Cl_z
      WT_CONSIDER(0, fdn[0], shfl_up(n_pp[2], 1), hasdn)
      WT_CONSIDER(0, f[0], n_pp[1], yes)
      WT_CONSIDER(0, fup[0], shfl_down(n_pp[0], 1), hasup)
      WT_CONSIDER(0, f[1], n_cp, yes)
      // column Cl_z: rows Cl_{z-1}, Cl_z, Cl_{z+1}
      WT_CONSIDER(1, fdn[1], shfl_up(n_cc[2], 1), hasdn)
      WT_CONSIDER(1, f[1], n_cc[1], yes)
      WT_CONSIDER(1, fup[1], shfl_down(n_cc[0], 1), hasup)
      // column T_z: rows pH, Cl, T of zones z-1, z, z+1
      WT_CONSIDER(2, fdn[0], shfl_up(n_pt[2], 1), hasdn)
      WT_CONSIDER(2, f[0], n_pt[1], yes)
      WT_CONSIDER(2, fup[0], shfl_down(n_pt[0], 1), hasup)
      WT_CONSIDER(2, fdn[1], shfl_up(n_ct[2], 1), hasdn)
      WT_CONSIDER(2, f[1], n_ct[1], yes)
      WT_CONSIDER(2, fup[1], shfl_down(n_ct[0], 1), hasup)
      WT_CONSIDER(2, fdn[2], shfl_up(n_tt[2], 1), hasdn)
      WT_CONSIDER(2, f[2], n_tt[1], yes)
      WT_CONSIDER(2, fup[2], shfl_down(n_tt[0], 1), hasup)
#undef WT_CONSIDER

      vb upd[3];
      WT_UNROLL
      for (int v = 0; v < 3; ++v) {
        // all differences zero -> argmax is row 0 of the system (pH of zone 0)
        vd sc = sel(best[v] > 0.0, vmax(bf[v], bn[v]), f_row0);
        if (pass == 0) {
          maxd[v] = best[v];
          scl[v] = sc;
          retry[v] = best[v] < REJECT * sc;
          upd[v] = vbroadcast_b(true);
        } else {
          upd[v] = retry[v] & (maxd[v] * sc < best[v] * scl[v]);
          jfac[v] = sel(m & upd[v], 10.0 * jfac[v], jfac[v]);
          hh[v] = sel(upd[v], hc[v], hh[v]);
          maxd[v] = sel(upd[v], best[v], maxd[v]);
          scl[v] = sel(upd[v], sc, scl[v]);
        }
      }
      // commit the evaluations of the columns that were (re)accepted: own column by this
      // lane's flag, neighbour columns by the owner's flag
      WT_UNROLL
      for (int v = 0; v < 3; ++v) {
        vb uo = upd[v], ud, uu;
        {
          vi ui = seli(upd[v], 1, 0);
          ud = shfl_up_i(ui, 1) != 0;
          uu = shfl_down_i(ui, 1) != 0;
        }
        if (v == 0) {
          d_pp[0] = sel(ud, n_pp[0] - f[0], d_pp[0]);
          d_pp[1] = sel(uo, n_pp[1] - f[0], d_pp[1]);
          d_pp[2] = sel(uu, n_pp[2] - f[0], d_pp[2]);
          d_cp = sel(uo, n_cp - f[1], d_cp);
        } else if (v == 1) {
          d_cc[0] = sel(ud, n_cc[0] - f[1], d_cc[0]);
          d_cc[1] = sel(uo, n_cc[1] - f[1], d_cc[1]);
          d_cc[2] = sel(uu, n_cc[2] - f[1], d_cc[2]);
        } else {
          d_pt[0] = sel(ud, n_pt[0] - f[0], d_pt[0]);
          d_pt[1] = sel(uo, n_pt[1] - f[0], d_pt[1]);
          d_pt[2] = sel(uu, n_pt[2] - f[0], d_pt[2]);
          d_ct[0] = sel(ud, n_ct[0] - f[1], d_ct[0]);
          d_ct[1] = sel(uo, n_ct[1] - f[1], d_ct[1]);
          d_ct[2] = sel(uu, n_ct[2] - f[1], d_ct[2]);
          d_tt[0] = sel(ud, n_tt[0] - f[2], d_tt[0]);
          d_tt[1] = sel(uo, n_tt[1] - f[2], d_tt[1]);
          d_tt[2] = sel(uu, n_tt[2] - f[2], d_tt[2]);
        }
      }
    }

    // ---- J = diff / h (column-wise h), edge columns do not exist -> 0
    vd hdn[3], hup[3];  // reciprocal steps of this and the neighbouring columns
    WT_UNROLL
    for (int v = 0; v < 3; ++v) { hh[v] = wt_rcp(hh[v]); hdn[v] = shfl_up(hh[v], 1); hup[v] = shfl_down(hh[v], 1); }
    vb hasdn = !g.first, hasup = !g.last;
#define WT_JSET(dst, val) dst = sel(m, (val), dst)
#define WT_JPARK(k, val) pkset(PK_JD + (k), (val), m)
    WT_JPARK(3, sel(hasdn, d_pp[0] * hdn[0], 0.0));
    WT_JPARK(4, d_pp[1] * hh[0]);
    WT_JPARK(5, sel(hasup, d_pp[2] * hup[0], 0.0));
    WT_JSET(J.cp, d_cp * hh[0]);
    WT_JPARK(6, sel(hasdn, d_cc[0] * hdn[1], 0.0));
    WT_JPARK(7, d_cc[1] * hh[1]);
    WT_JPARK(8, sel(hasup, d_cc[2] * hup[1], 0.0));
    WT_JSET(J.pt[0], sel(hasdn, d_pt[0] * hdn[2], 0.0));
    WT_JSET(J.pt[1], d_pt[1] * hh[2]);
    WT_JSET(J.pt[2], sel(hasup, d_pt[2] * hup[2], 0.0));
    WT_JSET(J.ct[0], sel(hasdn, d_ct[0] * hdn[2], 0.0));
    WT_JSET(J.ct[1], d_ct[1] * hh[2]);
    WT_JSET(J.ct[2], sel(hasup, d_ct[2] * hup[2], 0.0));
    WT_JPARK(0, sel(hasdn, d_tt[0] * hdn[2], 0.0));
    WT_JPARK(1, d_tt[1] * hh[2]);
    WT_JPARK(2, sel(hasup, d_tt[2] * hup[2], 0.0));
#undef WT_JSET
#undef WT_JPARK
    // ---- factor adaptation (common.py:377-380)
    WT_UNROLL
    for (int v = 0; v < 3; ++v) {
      vd fnew = jfac[v];
      fnew = sel(maxd[v] < SMALL * scl[v], fnew * 10.0, fnew);
      fnew = sel(maxd[v] > BIG * scl[v], fnew * 0.1, fnew);
      fnew = vmax(fnew, MINF);
      pkset(PK_JFAC + v, fnew, m);
    }
  }

  // radau.py:139-176
  WT_DEV vd predict_factor(vd h_abs, vd h_abs_old, vd err, vd err_old, vb have_old) const {
    vb noh = (!have_old) | (err == 0.0);
    // x ** 0.25 = sqrt(sqrt(x)) (differs from pow by rounding only)
    vd ie = wt_rcp(err);
    vd mult = sel(noh, 1.0, wt_div(h_abs, h_abs_old) * vsqrt(vsqrt(err_old * ie)));
    return vmin(mult, 1.0) * vsqrt(vsqrt(ie));
  }

  // Z_i[var] = sum_k T[i][k] W[k][var]   (radau.py:126)
  // (table-driven: with the stage loop rolled, an if-chain on i would be a branch inside the loop body)
  WT_DEV vd zrow(int i, int v) const {
    const int r = i < 2 ? 8 + 3 * i : 32;
    return (wt_rk[r] * W[0][v] + wt_rk[r + 1] * W[1][v]) + wt_rk[r + 2] * W[2][v];
  }

  // -------------------------------------------------------------------------------------
  // the whole step: solve_ivp(Radau) over [t0, t0+dt] from (y) -> y at t0+dt
  // -------------------------------------------------------------------------------------
  // The step is split in two so that the GPU can run it as two kernels (wt_kernels.cu):
  //   begin()  once-per-step work with the same control flow for every plant: solver set-up, f0, the initial step
  //            size and the first Jacobian (Radau.__init__, radau.py:295-347, 363-369);
  //   run()    the attempt loop (_step_impl, radau.py:405-545), data-dependent.
  // What crosses from one to the other: fl, J.pt / J.ct / J.cp, the parked f / jfac / J diagonal blocks and
  // PV_SELF_H (everything else begin() leaves behind is rebuilt by reset()).
  WT_DEV void integrate(vd t0, vd dt, vb plant_on, int max_attempts, double h_floor = 0.0) {
    begin(t0, dt, plant_on);
    if (h_floor > 0.0) run<true>(t0, max_attempts, h_floor);
    else run<false>(t0, max_attempts, 0.0);
  }

  // solver state at the start of a step (radau.py:295-347 minus f0 / h_abs / the Jacobian)
  WT_DEV void reset(vd t0, vd dt, vb plant_on) {
    pvset(PV_T_BOUND, t0 + dt);
    pvset(PV_MAX_STEP, vmin(dt, 10.0));
    // running = plant_on; current_jac and new_step start true (radau.py:363-369, :413)
    fl = seli(plant_on, (int)F_RUNNING, 0) | (int)(F_CURRENT_JAC | F_NEW_STEP);
    pvset(PV_SOL_TOLD, vbroadcast(0.0));
    pvset(PV_SOL_H, vbroadcast(1.0));
    const vb all = vbroadcast_b(true);
    WT_UNROLL
    for (int v = 0; v < 3; ++v) {
      pkset(PK_YOLD + v, y[v], all);
      WT_UNROLL
      for (int k = 0; k < 3; ++k) { W[k][v] = vbroadcast(0.0); pkset(PK_Q + 3 * v + k, vbroadcast(0.0), all); }
    }
    J.cp = vbroadcast(0.0);
    WT_UNROLL
    for (int k = 0; k < 3; ++k) { J.pt[k] = vbroadcast(0.0); J.ct[k] = vbroadcast(0.0); }
    pvset(PV_SELF_H_OLD, vbroadcast(0.0));
    pvset(PV_SELF_ERR_OLD, vbroadcast(0.0));
    pvset(PV_H_OLD, vbroadcast(0.0));
    pvset(PV_ERR_OLD, vbroadcast(0.0));
    pvset(PV_MIN_STEP, vbroadcast(0.0));
  }

  WT_DEV void begin(vd t0, vd dt, vb plant_on) {
    reset(t0, dt, plant_on);
    const vb all = vbroadcast_b(true);
    // ---- Radau.__init__: f0 and select_initial_step (radau.py:303-311, common.py:68-134)
    {
      vd h_abs;
      // f0 = f(y) and f1 = f(y + h0 f0) go through ONE rolled copy of the RHS (instruction-fetch bound kernel)
      vd sc[3], yy[3], fo[3], f[3];  // sc = 1 / scale
      vd d1 = vbroadcast(0.0), h0 = vbroadcast(0.0);
      const vd interval = vabs(pv(PV_T_BOUND) - t0);
      WT_UNROLL
      for (int v = 0; v < 3; ++v) { yy[v] = y[v]; sc[v] = wt_rcp(WT_ATOL + vabs(y[v]) * WT_RTOL); }
      WT_NOUNROLL
      for (int it = 0; it < 2; ++it) {
        vb bad;
        wt_rhs(g, c, yy[0], yy[1], yy[2], fo[0], fo[1], fo[2], bad);
        lu->cadd(WTC_NFEV, seli(fget(F_RUNNING), 1, 0));
        fset(F_TRANGE, fget(F_RUNNING) & wt_gany(g, bad));
        if (it == 0) {
          WT_UNROLL
          for (int v = 0; v < 3; ++v) f[v] = fo[v];
          vd d0 = rms3(y[0] * sc[0], y[1] * sc[1], y[2] * sc[2]);
          d1 = rms3(f[0] * sc[0], f[1] * sc[1], f[2] * sc[2]);
          h0 = sel((d0 < 1e-5) | (d1 < 1e-5), 1e-6, wt_div(0.01 * d0, d1));
          h0 = vmin(h0, interval);
          WT_UNROLL
          for (int v = 0; v < 3; ++v) yy[v] = y[v] + h0 * f[v];
        }
      }
      vd d2 = wt_div(rms3((fo[0] - f[0]) * sc[0], (fo[1] - f[1]) * sc[1], (fo[2] - f[2]) * sc[2]), h0);
      vd h1 = sel((d1 <= 1e-15) & (d2 <= 1e-15), vmax(h0 * 1e-3, 1e-6), vsqrt(vsqrt(wt_div(0.01, vmax(d1, d2)))));
      h_abs = vmin(vmin(100.0 * h0, h1), vmin(interval, pv(PV_MAX_STEP)));
      pvset(PV_SELF_H, h_abs);
      WT_UNROLL
      for (int v = 0; v < 3; ++v) pkset(PK_F + v, f[v], all);
    }
    fclr(F_RUNNING, trange());
    // first Jacobian (radau.py:363-369)
    {
      vb m = fget(F_RUNNING);
      if (vany(m)) {
        num_jac(m);
        fset(F_CURRENT_JAC, m);
        fclr(F_LU_VALID | F_NEED_JAC, m);
        fclr(F_RUNNING, trange());
      }
    }
  }

  // FLOOR (engine policy, not reference behaviour): the catch-up launches of plants that exhausted
  // their budget on the 8 C density discontinuity.  The step size of every attempt is kept >= h_floor, and AT the floor
  // (a) an error estimate above 1 no longer rejects, (b) a Newton iteration that stops unconverged with a current
  // Jacobian keeps its last iterate.  Such plant-steps report WTS_DEGRADED.  A compile-time variant: the ordinary
  // kernel carries none of it.
  template <bool FLOOR>
  WT_DEV void run(vd t0, int max_attempts, double h_floor) {
    if (max_attempts <= 0 || max_attempts > WT_HARD_MAX_ATTEMPTS) max_attempts = WT_HARD_MAX_ATTEMPTS;
    vd t = t0;
    vd h_abs = vbroadcast(0.0);  // set from PV_SELF_H on the first pass (F_NEW_STEP)
    vi attempts = vbroadcast_i(0);
    while (wt_cta_any(vany(fget(F_RUNNING)))) {
      // (1) Jacobian: first one, stale-J refresh (radau.py:467-473) or post-accept refresh (:519-521)
      {
        vb m = fget(F_RUNNING) & fget(F_NEED_JAC);
        if (vany(m)) {
          num_jac(m);
          fset(F_CURRENT_JAC, m);
          fclr(F_LU_VALID | F_NEED_JAC, m);
          fclr(F_RUNNING, trange());
        }
      }
      // base.py:204-208: finished once t reached t_bound
      fclr(F_RUNNING, t == pv(PV_T_BOUND));
#ifndef WT_CTA_LOCKSTEP
      if (!vany(fget(F_RUNNING))) break;
#endif

      // (2) _step_impl entry (radau.py:413-428)
      {
        vb m = fget(F_RUNNING) & fget(F_NEW_STEP);
        vd ms = 10.0 * vabs(vnextafter_up(t) - t);
        pvset(PV_MIN_STEP, ms, m);
        const vd self_h_abs = pv(PV_SELF_H), max_step = pv(PV_MAX_STEP);
        vb big = self_h_abs > max_step, small = self_h_abs < ms;
        vd hh_ = sel(big, max_step, sel(small, ms, self_h_abs));
        h_abs = sel(m, hh_, h_abs);
        if (FLOOR) h_abs = vmax(h_abs, h_floor);
        pvset(PV_H_OLD, pv(PV_SELF_H_OLD), m);
        pvset(PV_ERR_OLD, pv(PV_SELF_ERR_OLD), m);
        fclr(F_HAVE_OLD | F_REJECTED | F_NEW_STEP, m);
        fset(F_HAVE_OLD, m & fget(F_SELF_HAVE_OLD) & !(big | small));
      }
      // (3) attempt setup (radau.py:439-457)
      {
        vb f_ = fget(F_RUNNING) & (h_abs < pv(PV_MIN_STEP));
        fset(F_FAILED, f_);
        fclr(F_RUNNING, f_);
      }
      vd t_new = t + h_abs;
      {
        const vd t_bound = pv(PV_T_BOUND);
        t_new = sel(t_new - t_bound > 0.0, t_bound, t_new);
      }
      const vd h = t_new - t;
      h_abs = sel(fget(F_RUNNING), vabs(h), h_abs);
      vb at_floor = vbroadcast_b(false);
      // (h = fl(fl(t + h_floor) - t) can exceed h_floor by rounding: a relative slack far above that, far below a step ratio)
      if (FLOOR) at_floor = h_abs <= h_floor * 1.000001;
      vd scale[3];  // 1 / (atol + |y| rtol)
      WT_UNROLL
      for (int v = 0; v < 3; ++v) scale[v] = wt_rcp(WT_ATOL + vabs(y[v]) * WT_RTOL);
      const vd ih = wt_rcp(h);
      {
        const vd isolh = wt_rcp(pv(PV_SOL_H)), sol_told = pv(PV_SOL_TOLD);
        const vb hs = fget(F_HAVE_SOL);
        // Z0 = sol(t + h*C).T - y, W = TI.dot(Z0)   (radau.py:451-454, 555-578, :64)
        vd x0 = ((t + h * WT_C0) - sol_told) * isolh;
        vd x1 = ((t + h * WT_C1) - sol_told) * isolh;
        vd x2 = ((t + h * 1.0) - sol_told) * isolh;
        WT_UNROLL
        for (int v = 0; v < 3; ++v) {
          const vd q0 = pk(PK_Q + 3 * v), q1 = pk(PK_Q + 3 * v + 1), q2 = pk(PK_Q + 3 * v + 2), yo = pk(PK_YOLD + v);
          vd z0 = (((q0 * x0 + q1 * (x0 * x0)) + q2 * ((x0 * x0) * x0)) + yo) - y[v];
          vd z1 = (((q0 * x1 + q1 * (x1 * x1)) + q2 * ((x1 * x1) * x1)) + yo) - y[v];
          vd z2 = (((q0 * x2 + q1 * (x2 * x2)) + q2 * ((x2 * x2) * x2)) + yo) - y[v];
          z0 = sel(hs, z0, 0.0);
          z1 = sel(hs, z1, 0.0);
          z2 = sel(hs, z2, 0.0);
          W[0][v] = (WT_TI00 * z0 + WT_TI01 * z1) + WT_TI02 * z2;
          W[1][v] = (WT_TI10 * z0 + WT_TI11 * z1) + WT_TI12 * z2;
          W[2][v] = (WT_TI20 * z0 + WT_TI21 * z1) + WT_TI22 * z2;
        }
      }
      // (4) LU of (MU/h I - J), real and complex (radau.py:460-462)
      {
        vb m = fget(F_RUNNING) & !fget(F_LU_VALID);
        if (vany(m)) {
          factor(h, m, fget(F_RUNNING) & fget(F_LU_VALID));
          lu->cadd(WTC_NLU, seli(m, 2, 0));
          fset(F_LU_VALID, m);
        }
      }
      // Engine policy (straggler policy): budget of collocation solves per step.
      // The reference has none; plants sitting on the 8 C density discontinuity make it grind
      // through millions of micro-steps.  Exhausted budget == exception: state left untouched.
      {
        attempts = attempts + seli(fget(F_RUNNING), 1, 0);
        vb over = fget(F_RUNNING) & (attempts > max_attempts);
        fset(F_WORKLIMIT, over);
        fclr(F_RUNNING, over);
      }
      // (5)-(8) ONE pass loop for the simplified Newton iterations (radau.py:48-136), the error estimate
      // (radau.py:483-512) and the acceptance evaluation (radau.py:514-545).  The kernel is bound by instruction
      // fetch, so all of them go through ONE copy of "evaluate the RHS at three points -> solve -> scaled norm":
      //   Newton pass of a plant   points y + Z_i; real + complex right-hand sides; dW and the convergence logic
      //   closing pass 0 (cl0)     the pass after convergence: err = solve_real(f + ZE); the point y_new is evaluated
      //                            speculatively (scipy evaluates f(y_new) only after accepting; the evaluation, its
      //                            nfev count and its temperature check only take effect if the step is accepted)
      //   closing pass 1 (cl1)     only where a rejected step is about to be rejected again (radau.py:493-495):
      //                            err = solve_real(f(y + err) + ZE)
      // The plants of a warp are in different passes at the same time: a plant that converges early closes while its
      // neighbours still iterate.  In a closing pass the complex right-hand side is zero.
      const vd M_real = WT_MU_REAL * ih, Mc_re = WT_MU_CRE * ih, Mc_im = WT_MU_CIM * ih;
      vb converged = vbroadcast_b(false);
      vb active = fget(F_RUNNING), cl0 = vbroadcast_b(false), cl1 = vbroadcast_b(false);
      vd dW_norm_old = vbroadcast(0.0), rate = vbroadcast(0.0);
      vb have_norm_old = vbroadcast_b(false), have_rate = vbroadcast_b(false);
      vi n_iter = vbroadcast_i(0);
      vd y_new[3], ZE[3], escale[3], errv[3], f_new[3];
      vb bad_new = vbroadcast_b(false);
      vd err_norm = vbroadcast(0.0);
      WT_UNROLL
      for (int v = 0; v < 3; ++v) {
        y_new[v] = y[v]; ZE[v] = vbroadcast(0.0); escale[v] = vbroadcast(0.0); errv[v] = vbroadcast(0.0);
        f_new[v] = vbroadcast(0.0);
      }
      WT_NOUNROLL
      for (int k = 0; k < WT_NEWTON_MAXITER + 2; ++k) {
        const vb cl = cl0 | cl1;
        if (!vany(active | cl)) break;
        n_iter = seli(active, k + 1, n_iter);
        // A pass in which no plant iterates (all closing) needs ONE evaluation and no complex solve.
        const bool newton_pass = vany(active);
        vd fr[3], cr[3], ci[3], F[3], pc[3];
        WT_UNROLL
        for (int v = 0; v < 3; ++v) {
          fr[v] = vbroadcast(0.0); cr[v] = vbroadcast(0.0); ci[v] = vbroadcast(0.0); F[v] = vbroadcast(0.0);
          pc[v] = sel(cl1, y[v] + errv[v], y_new[v]);  // the point a closing plant evaluates
        }
        vb finite = vbroadcast_b(true), bad_any = vbroadcast_b(false), bad = vbroadcast_b(false);
        WT_STAGE_UNROLL
        for (int i = newton_pass ? 0 : 2; i < 3; ++i) {  // the three stages of a Newton pass; the LAST one is the closing evaluation
          vd pnt[3];
          WT_UNROLL
          for (int v = 0; v < 3; ++v) pnt[v] = sel(active, y[v] + zrow(i, v), pc[v]);
          wt_rhs(g, c, pnt[0], pnt[1], pnt[2], F[0], F[1], F[2], bad);
          bad_any = bad_any | (bad & active);
          finite = finite & visfinite(F[0]) & visfinite(F[1]) & visfinite(F[2]);
          const double tr = wt_rk[14 + i], t1 = wt_rk[17 + i], t2 = wt_rk[20 + i];
          WT_UNROLL
          for (int v = 0; v < 3; ++v) {
            fr[v] = fr[v] + F[v] * tr;
            cr[v] = cr[v] + F[v] * t1;
            ci[v] = ci[v] + F[v] * t2;
          }
        }
        // F, bad: the last evaluation = f(y_new) of a cl0 plant (speculative), f(y + err) of a cl1 plant (it counts)
        bad_any = bad_any | (bad & cl1);
        {
          vb tb = (active | cl1) & wt_gany(g, bad_any);
          fset(F_TRANGE, tb);
          fclr(F_RUNNING, tb);
          active = active & !tb;
          cl1 = cl1 & !tb;
        }
        active = active & !wt_gany(g, !finite);  // radau.py:91-92: break, not converged
        WT_UNROLL
        for (int v = 0; v < 3; ++v) {
          // Newton: TI F - M W (radau.py:96-99); closing: f + ZE resp. f(y + err) + ZE, no complex part
          const vd nr = fr[v] - M_real * W[0][v];
          const vd re = cr[v] - (Mc_re * W[1][v] - Mc_im * W[2][v]);
          const vd im = ci[v] - (Mc_re * W[2][v] + Mc_im * W[1][v]);
          const vd er = sel(cl1, F[v], pk(PK_F + v)) + ZE[v];
          fr[v] = sel(cl, er, nr);
          cr[v] = sel(cl, 0.0, re);
          ci[v] = sel(cl, 0.0, im);
        }
        solve_newton(fr, cr, ci, newton_pass);
        vd q = vbroadcast(0.0);
        WT_UNROLL
        for (int v = 0; v < 3; ++v) {
          const vd sr = sel(cl, escale[v], scale[v]);
          vd a = fr[v] * sr, b = cr[v] * scale[v], d = ci[v] * scale[v];
          q = q + ((a * a + b * b) + d * d);
        }
        // common.py:63-65 over the 3 n (closing) or 9 n (Newton) scaled unknowns
        const vd nrm = vsqrt(wt_gsum(g, q)) * sel(cl, g.inv_sqrtN, g.inv_sqrt3N);
        // ---- Newton pass: convergence control (radau.py:101-130).  (Not skipped in a pass without iterating plants:
        // the branch costs more in scheduling freedom than the ~100 masked instructions it would save -- measured.)
        vb cvn;
        {
          const vd dW_norm = nrm;
          vd new_rate = wt_div(dW_norm, dW_norm_old);
          rate = sel(active & have_norm_old, new_rate, rate);
          have_rate = have_rate | (active & have_norm_old);
          // rate ** (NEWTON_MAXITER - k)
          vd rp = rate;
          WT_NOUNROLL
          for (int e = 1; e < WT_NEWTON_MAXITER - k; ++e) rp = rp * rate;
          const vd i1r = wt_rcp(1.0 - rate);
          vb brk = active & have_rate & ((rate >= 1.0) | (rp * i1r * dW_norm > WT_NEWTON_TOL));
          vb forced = vbroadcast_b(false), upd = active & !brk;
          if (FLOOR) {  // at the floor a diverging iteration keeps its last iterate (no update) and closes
            forced = brk & at_floor & fget(F_CURRENT_JAC);
            brk = brk & !forced;
          }
          active = active & !brk;
          WT_UNROLL
          for (int v = 0; v < 3; ++v) {
            W[0][v] = sel(upd, W[0][v] + fr[v], W[0][v]);
            W[1][v] = sel(upd, W[1][v] + cr[v], W[1][v]);
            W[2][v] = sel(upd, W[2][v] + ci[v], W[2][v]);
          }
          cvn = active & ((dW_norm == 0.0) | (have_rate & (rate * i1r * dW_norm < WT_NEWTON_TOL)) | forced);
          if (FLOOR && k + 1 >= WT_NEWTON_MAXITER) {  // ... and so does one that has used up its iterations
            const vb f2 = active & !cvn & at_floor & fget(F_CURRENT_JAC);
            forced = forced | f2;
            cvn = cvn | f2;
          }
          if (FLOOR) fset(F_DEGRADED, forced);
          converged = converged | cvn;
          active = active & !cvn;
          dW_norm_old = sel(cl, dW_norm_old, dW_norm);
          have_norm_old = vbroadcast_b(true);
          if (k + 1 >= WT_NEWTON_MAXITER) active = vbroadcast_b(false);  // radau.py:87: at most NEWTON_MAXITER iterations
        }
        {
          // ---- closing pass: error norm, second estimate, accept / reject (radau.py:483-545)
          if (vany(cl)) {
            const vb clx = cl & fget(F_RUNNING);  // (not the plants whose evaluation at y + err just raised)
            err_norm = sel(clx, nrm, err_norm);
            const vb again = cl0 & fget(F_REJECTED) & (err_norm > 1.0);
            WT_UNROLL
            for (int v = 0; v < 3; ++v) {
              errv[v] = sel(cl0, fr[v], errv[v]);
              f_new[v] = sel(cl0, F[v], f_new[v]);
            }
            bad_new = selb(cl0, bad, bad_new);
            lu->cadd(WTC_NFEV, seli(again, 1, 0));  // the evaluation at y + err of the next pass
            const vb decide = clx & !again;
            vb rej = decide & (err_norm > 1.0);
            if (FLOOR) {  // at the floor the error test no longer rejects
              const vb frc = rej & at_floor;
              fset(F_DEGRADED, frc);
              rej = rej & !frc;
            }
            vb acc = decide & !rej;
            const vd safety = wt_div(vbroadcast(0.9 * (2 * WT_NEWTON_MAXITER + 1)), vfromint(n_iter + 2 * WT_NEWTON_MAXITER));
            const vd pf = predict_factor(h_abs, pv(PV_H_OLD), err_norm, pv(PV_ERR_OLD), fget(F_HAVE_OLD));
            h_abs = sel(rej, h_abs * vmax(safety * pf, 0.2), h_abs);
            if (FLOOR) h_abs = vmax(h_abs, h_floor);
            fclr(F_LU_VALID, rej);
            fset(F_REJECTED, rej);
            lu->cadd(WTC_NREJECT, seli(rej, 1, 0));
            // f(y_new): now it counts (radau.py:514-516); a temperature outside [0, 100] C raises there
            lu->cadd(WTC_NFEV, seli(acc, 1, 0));
            {
              const vb tb = acc & wt_gany(g, bad_new);
              fset(F_TRANGE, tb);
              fclr(F_RUNNING, tb);
              acc = acc & !tb;
            }
            const vb recompute = (n_iter > 2) & (rate > 1e-3);
            vd fct = vmin(safety * pf, 10.0);
            const vb keep = (!recompute) & (fct < 1.2);
            fct = sel(keep, 1.0, fct);
            fclr(F_LU_VALID, acc & !keep);
            fset(F_NEED_JAC, acc & recompute);
            fclr(F_CURRENT_JAC, acc);  // set again by num_jac when recomputed
            pvset(PV_SELF_H_OLD, pv(PV_SELF_H), acc);
            pvset(PV_SELF_ERR_OLD, err_norm, acc);
            pvset(PV_SELF_H, h_abs * fct, acc);
            WT_UNROLL
            for (int v = 0; v < 3; ++v) {
              const vd Z0 = zrow(0, v), Z1 = zrow(1, v), Z2 = zrow(2, v);  // W is final once a plant has converged
              pkset(PK_Q + 3 * v + 0, (Z0 * WT_P00 + Z1 * WT_P10) + Z2 * WT_P20, acc);
              pkset(PK_Q + 3 * v + 1, (Z0 * WT_P01 + Z1 * WT_P11) + Z2 * WT_P21, acc);
              pkset(PK_Q + 3 * v + 2, (Z0 * WT_P02 + Z1 * WT_P12) + Z2 * WT_P22, acc);
              pkset(PK_YOLD + v, y[v], acc);
              y[v] = sel(acc, y_new[v], y[v]);
              pkset(PK_F + v, f_new[v], acc);
            }
            pvset(PV_SOL_TOLD, t, acc);
            pvset(PV_SOL_H, t_new - t, acc);
            fset(F_SELF_HAVE_OLD | F_HAVE_SOL | F_NEW_STEP, acc);
            t = sel(acc, t_new, t);
            lu->cadd(WTC_NSTEPS, seli(acc, 1, 0));
            cl1 = again;
          }
          // ---- a plant that converged in this pass closes in the next one
          cl0 = cvn;
          if (vany(cvn)) {
            WT_UNROLL
            for (int v = 0; v < 3; ++v) {
              const vd Z0 = zrow(0, v), Z1 = zrow(1, v), Z2 = zrow(2, v);
              const vd yn = y[v] + Z2;
              y_new[v] = sel(cvn, yn, y_new[v]);
              ZE[v] = sel(cvn, ((Z0 * WT_E0 + Z1 * WT_E1) + Z2 * WT_E2) * ih, ZE[v]);
              escale[v] = sel(cvn, wt_rcp(WT_ATOL + vmax(vabs(y[v]), vabs(yn)) * WT_RTOL), escale[v]);
            }
          }
        }
      }
      // path counters of this attempt: n_iter Newton iterations with three RHS evaluations each (one update here
      // instead of two shared-memory read-modify-writes per iteration)
      lu->cadd(WTC_NNEWTON, n_iter);
      lu->cadd(WTC_NFEV, n_iter * 3);
      // (6) outcome of a collocation solve that did not converge (radau.py:464-481)
      {
        vb nc = fget(F_RUNNING) & !converged;
        lu->cadd(WTC_NNEWTON_FAIL, seli(nc, 1, 0));
        vb stale = nc & !fget(F_CURRENT_JAC);
        fset(F_NEED_JAC, stale);  // recompute J at (t, y, f), same h
        vb halve = nc & fget(F_CURRENT_JAC);
        h_abs = sel(halve, h_abs * 0.5, h_abs);
        if (FLOOR) h_abs = vmax(h_abs, h_floor);
        fclr(F_LU_VALID, halve);
      }
    }
  }
};

// ----------------------------------------------------------------------------------------
// reactor.py:490-541: post-processing of one step for this lane's zone.
// Returns the plant's status bits (replicated); writes derived[3] = {H, density, decay rate}.
// ----------------------------------------------------------------------------------------
template <class LuStore>
WT_DEV vi wt_finish_step(WtPlantStep<LuStore> &ps, const vd *y_in, vd *derived, vb &advance) {
  const WtGroup &g = ps.g;
  vi st = seli(ps.failed(), (int)WTS_SOLVER_FAILED, 0);
  st = st | seli(ps.trange(), (int)WTS_T_RANGE, 0) | seli(ps.worklimit(), (int)WTS_WORK_LIMIT, 0);
  st = st | seli(ps.degraded(), (int)WTS_DEGRADED, 0);
  advance = !(ps.trange() | ps.worklimit());  // exception inside solve_ivp: state untouched, time not advanced
  WT_UNROLL
  for (int v = 0; v < 3; ++v) ps.y[v] = sel(advance, ps.y[v], y_in[v]);
  vb nonfin = !(visfinite(ps.y[0]) & visfinite(ps.y[1]) & visfinite(ps.y[2]));
  st = st | seli(advance & wt_gany(g, nonfin), (int)WTS_NONFINITE, 0);
  // _update_derived_state (reactor.py:511-524); the decay rate raises when T is out of range
  derived[0] = vexp10(-ps.y[0]);
  derived[1] = wt_density(ps.y[2]);
  derived[2] = wt_arrhenius(ps.y[2]);
  vb tder = advance & wt_gany(g, wt_t_out_of_range(ps.y[2]));
  st = st | seli(tder, (int)WTS_T_RANGE_DERIVED, 0);
  // _enforce_physical_bounds (reactor.py:526-541), skipped when the line above raised
  vb clipok = advance & !tder;
  vb cp = clipok & wt_gany(g, (ps.y[0] < 0.0) | (ps.y[0] > 14.0));
  vb cc = clipok & wt_gany(g, ps.y[1] < 0.0);
  vb ct = clipok & wt_gany(g, (ps.y[2] < 0.0) | (ps.y[2] > 100.0));
  ps.y[0] = sel(cp, vmin(vmax(ps.y[0], 0.0), 14.0), ps.y[0]);
  ps.y[1] = sel(cc, vmax(ps.y[1], 0.0), ps.y[1]);
  ps.y[2] = sel(ct, vmin(vmax(ps.y[2], 0.0), 100.0), ps.y[2]);
  st = st | seli(cp, (int)WTS_CLIP_PH, 0) | seli(cc, (int)WTS_CLIP_CL, 0) | seli(ct, (int)WTS_CLIP_T, 0);
  return st;
}
