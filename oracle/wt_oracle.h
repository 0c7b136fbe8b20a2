/*
 * wt_oracle.h -- CPU restatement (TEST INFRASTRUCTURE, not product code) of the
 * reference hot path wt_simulator.core: IntegratedCSTR.step / derivatives and
 * AqueousChemistry.calculate_pH, plus the scipy Radau IIA integrator that
 * step() delegates to.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product path (the CUDA engine
 * behind include/wt_b200.h) never links or calls it.
 *
 * Parity pin: the reference ships no numeric test vectors for this path
 * (SURVEY.md section 4), so the oracle is pinned against outputs of the
 * reference itself, generated in the build container by oracle/gen_golden.py
 * (numpy 2.3.5 / scipy 1.18.1) and committed under tests/golden/.
 *
 * All citations are relative to the reference checkout (src/wt_simulator/...)
 * or to scipy 1.18.1 (scipy/integrate/_ivp/...).
 */
#ifndef WT_ORACLE_H
#define WT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-plant configuration vector (ReactorConfiguration, reactor.py:52-89) */
enum {
  WT_CFG_VOLUME = 0,
  WT_CFG_HEIGHT,
  WT_CFG_DIAMETER,
  WT_CFG_FLOW_RATE,
  WT_CFG_TURBULENT_INTENSITY,
  WT_CFG_RECIRCULATION_RATIO,
  WT_CFG_IMPELLER_SPEED,
  WT_CFG_IMPELLER_DIAMETER,
  WT_CFG_POWER_NUMBER,
  WT_CFG_INITIAL_PH,
  WT_CFG_ALKALINITY,
  WT_CFG_TOTAL_CARBONATE,
  WT_CFG_INITIAL_CHLORINE,
  WT_CFG_TEMPERATURE,
  WT_CFG_ENABLE_STRAT, /* 0.0 / 1.0 */
  WT_NCFG
};

/* ---- derived per-plant constants */
enum {
  WT_PAR_KW = 0,
  WT_PAR_KA1,
  WT_PAR_KA2,
  WT_PAR_KACL,
  WT_PAR_CT,    /* total_carbonate / 1000 [mol/L] */
  WT_PAR_KX,    /* K_exchange_per_s */
  WT_PAR_V,     /* superficial velocity [m/s] (from CONFIG flow) */
  WT_PAR_ZH,    /* zone height [m] */
  WT_PAR_VZL,   /* zone volume [L] */
  WT_PAR_VOLUME,/* tank volume [L] */
  WT_PAR_AT,    /* lateral + end area for heat loss [m^2] */
  WT_PAR_STRAT, /* 0.0 / 1.0 */
  WT_NPAR
};

/* ---- boundary vector (BoundaryConditions, reactor.py:150-186) */
enum {
  WT_BND_INLET_FLOW = 0,
  WT_BND_INLET_PH,
  WT_BND_INLET_CL,
  WT_BND_INLET_T,
  WT_BND_ACID_FLOW,
  WT_BND_ACID_CONC,
  WT_BND_CL_FLOW,
  WT_BND_CL_CONC,
  WT_BND_AMBIENT_T,
  WT_BND_HEAT_LOSS,
  WT_NBND
};

/* ---- per-plant status bits written by step */
enum {
  WT_ST_SOLVER_FAILED = 1u << 0,   /* Radau: step size below spacing (reactor.py:486-487 logs) */
  WT_ST_T_RANGE = 1u << 1,         /* ValueError from celsius_to_kelvin inside the solve; state unchanged */
  WT_ST_CLIP_PH = 1u << 2,         /* reactor.py:528-531 */
  WT_ST_CLIP_CL = 1u << 3,         /* reactor.py:533-536 */
  WT_ST_CLIP_T = 1u << 4,          /* reactor.py:538-541 */
  WT_ST_NONFINITE = 1u << 5,       /* a state value is NaN/inf after the step */
  WT_ST_T_RANGE_DERIVED = 1u << 6, /* ValueError in _update_derived_state (reactor.py:521-524): state assigned, not clipped */
  WT_ST_WORK_LIMIT = 1u << 7,      /* engine policy (not in the reference): attempt budget exhausted; state unchanged */
  WT_ST_DEFERRED = 1u << 8,        /* engine bookkeeping (never set by the oracle) */
  WT_ST_DEGRADED = 1u << 9         /* engine policy (not in the reference): floor mode forced an acceptance in this step */
};

/* ---- solver path counters */
enum {
  WT_CNT_NFEV = 0,  /* scipy nfev (excludes finite-difference Jacobian evaluations) */
  WT_CNT_NJEV,
  WT_CNT_NLU,
  WT_CNT_NSTEPS,    /* accepted internal Radau steps */
  WT_CNT_NNEWTON,   /* simplified-Newton iterations, all attempts */
  WT_CNT_NREJECT,   /* error-norm rejections */
  WT_CNT_NNEWTON_FAIL, /* collocation solves that did not converge */
  WT_CNT_RESERVED,
  WT_NCNT
};

#define WT_MAX_ZONES 32

/* reactor.py:203-270, chemistry.py:116-132, transport.py:202-254, 282-290.
 * Returns 0, or -1 if a construction-time check of the reference would raise. */
int wt_oracle_derive_params(const double *cfg, int n_zones, double *par);

/* reactor.py:272-448.  y = [pH(n), Cl(n), T(n)] species-major.
 * Returns 0, or 1 if the reference would raise ValueError (T outside [0,100]). */
int wt_oracle_rhs(const double *par, const double *bnd, int n, const double *y, double *dy);

/* common.py:260-382 at (y, f(y)); J row-major [3n x 3n]; factor_inout[3n]. */
void wt_oracle_num_jac(const double *par, const double *bnd, int n, const double *y, double *J,
                       double *factor_inout, int *have_factor);

/* reactor.py:450-541 with scipy Radau (radau.py, common.py, base.py).
 * y: in/out state (species-major), t: in/out plant time.
 * derived (optional, may be NULL): [H(n), density(n), decay_rate(n)].
 * Returns the status bitmask. */
uint32_t wt_oracle_step(const double *par, const double *bnd, int n, double dt,
                        double *t, double *y, double *flow_rate, double *derived,
                        int32_t *counters);

/* Engine policy knob mirrored for tests: budget of collocation solves per step (0 = unlimited). */
void wt_oracle_set_max_attempts(int m);
void wt_oracle_set_floor_div(int d); /* engine policy mirror: step-size floor dt / d with forced acceptance (0 = off) */

/* Batched driver: plant p uses cfg-derived par[p*WT_NPAR..], bnd[p*WT_NBND..] (or
 * bnd broadcast when bnd_stride == 0), y[p*3n..].  Runs `nsteps` steps of dt on
 * each plant with `nthreads` pthreads.  counters accumulate over steps. */
void wt_oracle_step_batch(int P, int n, int nsteps, double dt, const double *par,
                          const double *bnd, int bnd_stride, double *t, double *y,
                          double *flow_rate, uint32_t *status, int32_t *counters,
                          int nthreads);

/* chemistry.py:271-330 (with :193-269).  status: 0 ok, 1 derivative too small
 * (RuntimeError), 2 not converged after max_iter (RuntimeError).
 * temperature_c is the BufferSystem temperature (constants per chemistry.py:116-132). */
int wt_oracle_calc_ph(double alkalinity, double total_carbonate, double temperature_c,
                      double initial_guess, double tolerance, int max_iter,
                      double *ph_out, int32_t *iters_out);

/* test knob: perturb H by (1 +- eps) inside calculate_pH (stability probe) */
void wt_oracle_set_ph_h_eps(double eps);

void wt_oracle_calc_ph_batch(int P, const double *alk, const double *ct, const double *temp,
                             const double *guess, double *ph, int32_t *iters, int32_t *status,
                             int nthreads);

#ifdef __cplusplus
}
#endif
#endif
