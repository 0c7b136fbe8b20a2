// wt_emu.cpp -- TEST BUILD of the kernel logic for the CPU-only build container.
//
// Compiles ics_wt_physicsengine_b200/csrc/wt_step_core.h with -DWT_EMU, i.e. the exact source
// the sm_100a kernel is built from, with every per-lane register widened to a 32-entry array
// (wt_simt.h).  Used by tests/test_emu_core.py to compare the warp-lockstep algorithm
// (structured Jacobian, PCR solves, per-plant masking) with the CPU oracle before spending
// GPU time.  Not a fallback: the Python package never loads this library.
#include <string.h>

#include "wt_step_core.h"

struct EmuLu {
  vd slot[128];
  void put(int s, const vd &x, const vb &m) { for (int l = 0; l < 32; ++l) if (m.v[l]) slot[s].v[l] = x.v[l]; }
  vd get(int s) const { return slot[s]; }
  vd cslot[128];  // the complex slot space (tensor memory on the GPU)
  void cx_put4(int s, const vd *x, const vb &m) { for (int i = 0; i < 4; ++i) for (int l = 0; l < 32; ++l) if (m.v[l]) cslot[s + i].v[l] = x[i].v[l]; }
  void cx_get4(int s, vd *x) const { for (int i = 0; i < 4; ++i) x[i] = cslot[s + i]; }
  void begin_factor(const vb &) {}
  void end_factor() {}
  vd cst[CK_N];
  void cput(int k, const vd &x) { cst[k] = x; }
  vd cget(int k) const { return cst[k]; }
  void csync() {}
  vi cnt[WTC_NCNT];
  vd pvs[16];
  vd pvget(int k) const { return pvs[k]; }
  void pvput(int k, const vd &x) { pvs[k] = x; }
  void czero() { for (int k = 0; k < WTC_NCNT; ++k) cnt[k] = vbroadcast_i(0); }
  void cadd(int k, const vi &inc) { cnt[k] = cnt[k] + inc; }
};

extern "C" {

// Same argument convention as wt_oracle_step_batch (AoS per plant, species-major y).
void wt_emu_step_batch(int P, int n, int nsteps, double dt, const double *par, const double *bnd,
                       int bnd_stride, double *t, double *y, double *flow_rate, uint32_t *status,
                       int32_t *counters, double *derived, int max_attempts, int floor_div) {
  const int gpw = 32 / n;
  for (int p0 = 0; p0 < P; p0 += gpw) {
    for (int s = 0; s < nsteps; ++s) {
      static EmuLu lu;
      WtPlantStep<EmuLu> ps;
      ps.g = wt_make_group(n);
      ps.lu = &lu;
      ps.pk0 = 64;  // real LU slots end at 30, the lane constants sit at 110
      vd vpar[WTP_NPAR], vbnd[WTB_NBND], t0, vdt = vbroadcast(dt);
      vb on;
      vd yin[3];
      for (int l = 0; l < 32; ++l) {
        int gi = l / n, z = l % n;
        int p = p0 + gi;
        bool ok = gi < gpw && p < P && !(status && (status[p] & WTS_HALT_MASK));
        on.v[l] = ok;
        int pp = ok ? p : p0;
        if (!ok) z = 0;
        for (int k = 0; k < WTP_NPAR; ++k) vpar[k].v[l] = par[(size_t)pp * WTP_NPAR + k];
        for (int k = 0; k < WTB_NBND; ++k) vbnd[k].v[l] = bnd[(size_t)pp * bnd_stride + k];
        t0.v[l] = t[pp];
        for (int v = 0; v < 3; ++v) yin[v].v[l] = y[(size_t)pp * 3 * n + v * n + z];
      }
      ps.c = wt_make_const(&lu, ps.g, 110, vpar, vbnd);
      lu.czero();
      for (int v = 0; v < 3; ++v) ps.y[v] = yin[v];
      ps.integrate(t0, vdt, on, max_attempts, floor_div > 0 ? dt / (double)floor_div : 0.0);
      vd der[3];
      vb adv;
      vi st = wt_finish_step(ps, yin, der, adv);
      for (int l = 0; l < 32; ++l) {
        if (!on.v[l]) continue;
        int gi = l / n, z = l % n, p = p0 + gi;
        for (int v = 0; v < 3; ++v) {
          y[(size_t)p * 3 * n + v * n + z] = ps.y[v].v[l];
          if (derived) derived[(size_t)p * 3 * n + v * n + z] = der[v].v[l];
        }
        if (z == 0) {
          if (adv.v[l]) {
            t[p] += dt;
            if (flow_rate)
              flow_rate[p] = bnd[(size_t)p * bnd_stride + WTB_INLET_FLOW] + bnd[(size_t)p * bnd_stride + WTB_ACID_FLOW] +
                             bnd[(size_t)p * bnd_stride + WTB_CL_FLOW];
          }
          if (status) status[p] = (uint32_t)st.v[l];
          if (counters)
            for (int k = 0; k < WTC_NCNT; ++k) counters[(size_t)p * WTC_NCNT + k] += lu.cnt[k].v[l];
        }
      }
    }
  }
}

// RHS of one plant through the lane-mapped code path (plant replicated in group 0).
void wt_emu_rhs(const double *par, const double *bnd, int n, const double *y, double *dy, int *bad_out) {
  WtGroup g = wt_make_group(n);
  vd vpar[WTP_NPAR], vbnd[WTB_NBND], yy[3], d[3];
  for (int l = 0; l < 32; ++l) {
    int z = l < n ? l : 0;
    for (int k = 0; k < WTP_NPAR; ++k) vpar[k].v[l] = par[k];
    for (int k = 0; k < WTB_NBND; ++k) vbnd[k].v[l] = bnd[k];
    for (int v = 0; v < 3; ++v) yy[v].v[l] = y[v * n + z];
  }
  static EmuLu st;
  WtConstT<EmuLu> c = wt_make_const(&st, g, 110, vpar, vbnd);
  vb bad;
  wt_rhs(g, c, yy[0], yy[1], yy[2], d[0], d[1], d[2], bad);
  int b = 0;
  for (int l = 0; l < n; ++l) {
    for (int v = 0; v < 3; ++v) dy[v * n + l] = d[v].v[l];
    b |= bad.v[l];
  }
  *bad_out = b;
}
}
