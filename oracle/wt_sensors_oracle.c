/*
 * wt_sensors_oracle.c -- CPU restatement of the reference sensor suite.  TEST INFRASTRUCTURE ONLY
 * (same rules as wt_oracle.c: only tests/, smoke() and bench.py's CPU arms may use it).
 *
 * Follows the reference class by class, one plant at a time, plain structs:
 *   BaseSensor.read / calibrate / _check_for_faults / _apply_installation_effects   base_sensor.py:357-755
 *   SampleLine.transport_sample                                                     base_sensor.py:177-216
 *   pHSensor, ChlorineSensor (amperometric, DPD), TemperatureSensor (RTD PT100 / PT1000, thermocouple K / J),
 *   FlowSensor (magnetic, turbine), BaseSensor.reset
 *   create_realistic_sensor_suite + __main__.initialize_sensors                     sensors/__init__.py:41-120, __main__.py:84-118
 *
 * Parity pin: the reference seeds each sensor from secrets.randbits (base_sensor.py:331), so no
 * stream can be reproduced; this port is pinned IN DISTRIBUTION against 10,240 instances of the
 * reference suite (tests/golden/sensors_default_plant.npz, oracle/gen_golden_sensors.py) and the
 * CUDA kernel is then compared with this port value for value, because both draw from the same
 * counter-based Philox4x32-10 stream (counter = global plant id, read index, sensor*16 + block).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#define NSENS 7
#define RING 100

enum { SS_NORMAL = 0, SS_CALIBRATING, SS_WARMING_UP, SS_FAILED, SS_SATURATED, SS_DRIFT_WARNING, SS_CAL_EXPIRED,
       SS_OPEN_CIRCUIT, SS_SHORT_CIRCUIT, SS_OUT_OF_RANGE, SS_POWER_FAULT, SS_RATE_FAULT };
enum { F_NONE = 0, F_OPEN, F_SHORT, F_RANGE, F_RATE, F_POWER_LOW, F_POWER_HIGH };
enum { K_PH = 0, K_CL_AMP, K_CL_DPD, K_FLOW, K_TEMP };

typedef struct {
  /* BaseSensor attributes */
  int kind, zone_first, line;        /* zone_first: zone_index 0 (else -1) */
  double min_value, max_value, precision, drift_rate, warmup_time_s, max_rate_of_change, validity_hours;
  double current_value, supply_voltage, calibration_offset, last_calibration_time, power_on_time, cal_timestamp;
  int status, fault;
  double last_value;                 /* reading_history[-1].value */
  /* subclass attributes */
  double membrane_fouling, reference_contamination, days_since_cleaning; /* pH */
  double membrane_age_days;                                             /* amperometric */
  double reagent_potency, light_exposure_hours, reagent_age_days;       /* DPD */
  double electrode_fouling, full_scale;                                 /* flow (magnetic) */
  double bearing_wear_days;                                             /* flow (turbine), flow_sensor.py:90-92 */
  double cold_junction_drift;                                           /* temperature (thermocouple), temperature_sensor.py:96-99 */
  int variant;                       /* temperature: 0 RTD_PT100, 1 RTD_PT1000, 2 THERMOCOUPLE_K, 3 THERMOCOUPLE_J; flow: 0 MAGNETIC, 1 TURBINE */
  int hist_count;                    /* len(reading_history) */
  int has_cal_record;                /* bool(calibration_history): reset() clears the list, calibrate() appends */
  double slope_percentage;           /* pH: only refreshed by a read while a calibration record exists (ph_sensor.py:274-279) */
} sensor_t;

typedef struct {
  double ts[RING], val[RING];        /* deque(maxlen=100) as a ring: oldest at (head - count) */
  int head, count;
} line_t;

typedef struct {
  sensor_t s[NSENS];
  line_t line[2];
} suite_t;

int wt_oracle_suite_bytes(void) { return (int)sizeof(suite_t); }

/* ---- Philox4x32-10 + Box-Muller: the engine's RNG contract (see csrc/wt_sensors.cuh) ---- */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *o) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
typedef struct { uint32_t c0, c1, sens, k0, k1; } rng_t;
static void uniform2(const rng_t *g, int block, double *u0, double *u1) {
  uint32_t o[4];
  philox(g->c0, g->c1, g->sens * 16 + (uint32_t)block, 0x57544232u, g->k0, g->k1, o);
  *u0 = (double)((((uint64_t)(o[0] >> 5)) << 26) | (o[1] >> 6)) * (1.0 / 9007199254740992.0);
  *u1 = (double)((((uint64_t)(o[2] >> 5)) << 26) | (o[3] >> 6)) * (1.0 / 9007199254740992.0);
}
static void normal2(const rng_t *g, int block, double *z0, double *z1) {
  double u0, u1;
  uniform2(g, block, &u0, &u1);
  double r = sqrt(-2.0 * log(1.0 - u0));
  *z0 = r * cos(6.283185307179586 * u1);
  *z1 = r * sin(6.283185307179586 * u1);
}

/* ---- construction: create_realistic_sensor_suite + initialize_sensors ---- */
static void base_init(sensor_t *s, int kind, int zone_first, int line, double lo, double hi, double precision,
                      double drift_rate, double warmup, double max_rate, double validity_hours) {
  memset(s, 0, sizeof(*s));
  s->kind = kind; s->zone_first = zone_first; s->line = line;
  s->min_value = lo; s->max_value = hi; s->precision = precision; s->drift_rate = drift_rate;
  s->warmup_time_s = warmup; s->max_rate_of_change = max_rate; s->validity_hours = validity_hours;
  s->current_value = (lo + hi) / 2.0;     /* base_sensor.py:303 */
  s->status = SS_NORMAL; s->fault = F_NONE;
  s->supply_voltage = 24.0;               /* :309 */
  s->last_value = NAN;
  s->slope_percentage = 100.0;            /* ph_sensor.py:118 */
}
static void calibrate(sensor_t *s, double reference, double t) { /* base_sensor.py:701-755 */
  double offset = reference - s->current_value;
  s->calibration_offset = offset;
  s->last_calibration_time = t;
  s->status = SS_NORMAL;
  s->fault = F_NONE;
  s->power_on_time = t;
  s->cal_timestamp = t;
  s->has_cal_record = 1;
}
static void suite_init(suite_t *u, double t0, double cfg_flow, double cfg_cl, double cfg_T, int temp_kind, int flow_kind) {
  memset(u, 0, sizeof(*u));
  /* sensors/__init__.py:69-118 */
  base_init(&u->s[0], K_PH, 1, 0, 0.0, 14.0, 0.01, 0.01 / 24.0, 1800.0, 0.5, 24.0);   u->s[0].current_value = 7.0;
  base_init(&u->s[1], K_PH, 0, 1, 0.0, 14.0, 0.01, 0.01 / 24.0, 1800.0, 0.5, 24.0);   u->s[1].current_value = 7.0;
  base_init(&u->s[2], K_CL_AMP, 1, -1, 0.0, 10.0, 0.01, 0.02 / 24.0, 300.0, 1.0, 24.0); u->s[2].current_value = 0.0;
  base_init(&u->s[3], K_CL_DPD, 0, -1, 0.0, 10.0, 0.02, 0.02 / 24.0, 60.0, 1.0, 24.0);  u->s[3].current_value = 0.0;
  u->s[3].reagent_potency = 1.0;
  double FS = cfg_flow * 2.0;
  /* flow_sensor.py:66-69: turbine precision 1 % of full scale, magnetic 0.5 % */
  base_init(&u->s[4], K_FLOW, 1, -1, 0.0, FS, (flow_kind == 1 ? 0.01 : 0.005) * FS, 0.0, 10.0, FS, 8760.0); u->s[4].current_value = 0.0;
  u->s[4].full_scale = FS;
  u->s[4].variant = flow_kind;
  /* temperature_sensor.py:65-68: RTD precision 0.1 C, thermocouple 0.5 C */
  const double tprec = temp_kind >= 2 ? 0.5 : 0.1;
  base_init(&u->s[5], K_TEMP, 1, 0, -10.0, 110.0, tprec, 0.0, 30.0, 10.0, 8760.0);      u->s[5].current_value = 20.0;
  base_init(&u->s[6], K_TEMP, 0, 1, -10.0, 110.0, tprec, 0.0, 30.0, 10.0, 8760.0);      u->s[6].current_value = 20.0;
  u->s[5].variant = u->s[6].variant = temp_kind;
  /* __main__.py:96-105 */
  calibrate(&u->s[0], 7.0, t0);
  calibrate(&u->s[1], 7.0, t0);
  calibrate(&u->s[2], cfg_cl, t0);
  calibrate(&u->s[3], cfg_cl, t0);
  calibrate(&u->s[4], cfg_flow, t0);
  calibrate(&u->s[5], cfg_T, t0);
  calibrate(&u->s[6], cfg_T, t0);
}

/* ---- SampleLine.transport_sample, base_sensor.py:177-216 ---- */
static double transport_sample(line_t *L, double value, double timestamp, double delay) {
  L->ts[L->head] = timestamp;
  L->val[L->head] = value;
  L->head = (L->head + 1) % RING;
  if (L->count < RING) L->count++;
  double target = timestamp - delay;
  int i0 = (L->head - L->count + 2 * RING) % RING;
  int closest = i0;
  double min_diff = fabs(L->ts[i0] - target);
  for (int k = 0; k < L->count; ++k) {
    int i = (i0 + k) % RING;
    double d = fabs(L->ts[i] - target);
    if (d < min_diff) { min_diff = d; closest = i; }
  }
  return L->val[closest];
}

typedef struct { double value, raw_value, noise, drift, uncertainty; int status, fault; } reading_t;

/* BaseSensor.read, base_sensor.py:509-699 */
static reading_t base_read(sensor_t *s, suite_t *u, const rng_t *g, const double *inst, double true_value,
                           double t, double t_prev, int have_prev) {
  reading_t r;
  double v0 = s->supply_voltage;
  if (!(20.0 < v0 && v0 < 28.0)) {                       /* :556-577 */
    r.value = NAN; r.raw_value = NAN; r.noise = 0.0; r.drift = 0.0; r.uncertainty = 0.0;
    r.status = SS_POWER_FAULT; r.fault = v0 < 20.0 ? F_POWER_LOW : F_POWER_HIGH;
    s->last_value = NAN;
    return r;
  }
  double zv, zn;
  normal2(g, 0, &zv, &zn);
  s->supply_voltage = 24.0 + zv * 1.0;                    /* :579 */
  if (!(t - s->power_on_time >= s->warmup_time_s)) {      /* :582-595 */
    r.value = NAN; r.raw_value = NAN; r.noise = 0.0; r.drift = 0.0; r.uncertainty = 0.0;
    r.status = SS_WARMING_UP; r.fault = F_NONE;
    s->last_value = NAN;
    return r;
  }
  /* :598-600, _check_calibration_valid :432-436: no record -> not valid */
  int cal_expired = !s->has_cal_record || ((t - s->cal_timestamp) / 3600.0) > s->validity_hours;
  if (cal_expired) s->status = SS_CAL_EXPIRED;
  if (s->line >= 0) true_value = transport_sample(&u->line[s->line], true_value, t, inst[5]);   /* :603-614 */
  double drift_hours = (t - s->last_calibration_time) / 3600.0;
  double current_drift = s->drift_rate * drift_hours + s->calibration_offset;                    /* :617-620 */
  double noise = zn * s->precision;
  double alpha = 0.5;
  double raw_with_noise = true_value + noise + current_drift;
  s->current_value = alpha * raw_with_noise + (1 - alpha) * s->current_value;                    /* :626-630 */
  /* hysteresis: a no-op because the argument IS current_value (:438-462, :633) */
  /* _apply_installation_effects, :464-507 */
  if (inst[0] < 0.1 || inst[2] < 0.8 || inst[3] > 0.2 || inst[1] > 0.0) {
    double g0, g1, g2, g3, ub, ub2;
    normal2(g, 5, &g0, &g1);
    normal2(g, 6, &g2, &g3);
    uniform2(g, 4, &ub, &ub2);
    double v = s->current_value;
    if (inst[0] < 0.1) v += g0 * (s->precision * 2.0);
    if (inst[1] > 0.0 && ub < inst[1] / 60.0) v = NAN;
    else {
      if (inst[2] < 0.8) v += g1 * (s->precision * (2.0 - inst[2]));
      if (inst[3] > 0.2) v += g2 * (inst[3] * s->precision);
    }
    s->current_value = v;
  }
  double rate = 0.0;                                                                              /* :638-646 */
  if (have_prev) {
    double dt = t - t_prev;
    if (dt > 0 && isfinite(s->last_value)) rate = (s->current_value - s->last_value) / dt;
  }
  /* _check_for_faults, :357-409 */
  int fault = F_NONE;
  double span = s->max_value - s->min_value;
  if (!(20.0 < s->supply_voltage && s->supply_voltage < 28.0)) fault = s->supply_voltage < 20.0 ? F_POWER_LOW : F_POWER_HIGH;
  else if (s->current_value < s->min_value - 0.1 * span || s->current_value > s->max_value + 0.1 * span) fault = F_RANGE;
  else if (fabs(rate) > s->max_rate_of_change) fault = F_RATE;
  else {
    double u0, u1;
    uniform2(g, 1, &u0, &u1);
    if (u0 < 0.0001) fault = u1 < 0.5 ? F_OPEN : F_SHORT;
  }
  if (fault != F_NONE) {                                                                          /* :651-662 */
    s->fault = fault;
    if (fault == F_OPEN || fault == F_SHORT) { s->status = SS_FAILED; s->current_value = NAN; }
    else if (fault == F_RANGE) s->status = SS_OUT_OF_RANGE;
    else if (fault == F_POWER_LOW || fault == F_POWER_HIGH) s->status = SS_POWER_FAULT;
    else s->status = SS_RATE_FAULT;
  } else {                                                                                        /* :663-682 */
    s->fault = F_NONE;
    if (!isnan(s->current_value)) {
      double b = fmin(fmax(s->current_value, s->min_value), s->max_value);
      if (b != s->current_value) s->status = SS_SATURATED;
      else if (!cal_expired) s->status = SS_NORMAL;
      s->current_value = b;
    }
    if (fabs(current_drift) > 0.1 * span && s->status != SS_CAL_EXPIRED) s->status = SS_DRIFT_WARNING;
  }
  r.value = s->current_value; r.raw_value = true_value; r.noise = noise; r.drift = current_drift;
  r.status = s->status; r.uncertainty = s->precision * 2.0; r.fault = s->fault;
  s->last_value = r.value;
  return r;
}

static double clip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }

static reading_t sensor_read(sensor_t *s, suite_t *u, const rng_t *g, const double *inst, const double *pH,
                             const double *Cl, const double *T, int n, double flow_rate, double t, double t_prev,
                             int have_prev) {
  int z = s->zone_first ? 0 : n - 1;
  double true_value;
  switch (s->kind) {
    case K_PH: true_value = pH[z] + 0.003 * (T[z] - 25.0); break;                       /* ph_sensor.py:162-180 */
    case K_CL_AMP:
    case K_CL_DPD: {                                                                      /* chlorine_sensor.py:202-227 */
      double ratio = pow(10.0, 7.5 - pH[z]);
      true_value = Cl[z] * (0.5 + 0.5 * (ratio / (1 + ratio)));
      break;
    }
    case K_FLOW: true_value = flow_rate; break;                                           /* flow_sensor.py:98-102 */
    default: true_value = T[z]; break;                                                    /* temperature_sensor.py:103-108 */
  }
  reading_t r = base_read(s, u, g, inst, true_value, t, t_prev, have_prev);
  if (!isfinite(r.value)) return r;
  double dt = t - t_prev;
  double n0, n1;
  normal2(g, 2, &n0, &n1);
  double final_value;
  if (s->kind == K_PH) {                                                                  /* ph_sensor.py:216-336 */
    if (have_prev) {                                                                      /* _update_fouling :182-214 */
      double bio_rate = s->membrane_fouling > 0.05 ? 0.1 * exp(0.05 * (T[z] - 25)) : 0.001;
      double scaling_rate = inst[0] < 0.1 ? 100.0 * 0.0001 : 100.0 * 0.00001;
      s->membrane_fouling += (bio_rate + scaling_rate) * (dt / 86400.0);
      s->membrane_fouling = fmin(1.0, s->membrane_fouling);
      s->days_since_cleaning += dt / 86400.0;
    }
    double electrical_noise = n0 * (0.002 * (1.0 + 0.1 * fabs(r.value - 7.0)));
    double junction_noise = n1 * (0.005 * (1.0 + s->reference_contamination));
    double days = 0.0;                                                                    /* :274-279 */
    if (s->has_cal_record) {
      days = (t - s->cal_timestamp) / 86400.0;
      s->slope_percentage = fmax(90.0, 100.0 - 0.001 * days);
    }
    double slope_error = 0.0;
    if (!(4.0 < r.value && r.value < 7.0))
      slope_error = fmin(fabs(r.value - 4.0), fabs(r.value - 7.0)) * (100.0 - s->slope_percentage) / 100.0;
    double fouling_offset = s->membrane_fouling * 0.2;
    double f0, f1;
    normal2(g, 3, &f0, &f1);
    double fouling_noise = f0 * (s->membrane_fouling * 0.05);
    s->reference_contamination += 0.0001 * (days / 30.0);
    s->reference_contamination = fmin(0.5, s->reference_contamination);
    double reference_offset = s->reference_contamination * 0.1;
    final_value = r.value + electrical_noise + junction_noise + slope_error + fouling_offset + fouling_noise + reference_offset;
    final_value = clip(final_value, s->min_value, s->max_value);
    r.noise = r.noise + electrical_noise + junction_noise + fouling_noise;
    r.drift = r.drift + slope_error + fouling_offset + reference_offset;
    r.uncertainty = s->precision * 3.0;
  } else if (s->kind == K_CL_AMP) {                                                       /* chlorine_sensor.py:345-449 */
    if (have_prev) {
      s->membrane_fouling += (inst[0] < 0.1 ? 0.05 : 0.01) * (dt / 86400.0);
      s->membrane_fouling = fmin(1.0, s->membrane_fouling);
      s->membrane_age_days += dt / 86400.0;
    }
    double fouling_factor = 1.0 - 0.8 * s->membrane_fouling;
    double polarization_noise = n0 * (0.005 * (1.0 + s->membrane_age_days / 365.0));
    double diffusion_noise = n1 * 0.003;
    final_value = clip((r.value + 0.0) * fouling_factor + polarization_noise + diffusion_noise, s->min_value, s->max_value);
  } else if (s->kind == K_CL_DPD) {                                                       /* :280-317, 451-484 */
    if (have_prev) {
      double thermal_factor = exp((50000.0 / 8.314) * (1 / 293.15 - 1 / (20.0 + 273.15)));
      s->light_exposure_hours += dt / 3600.0;
      double photo_factor = 1.0 + 0.1 * (s->light_exposure_hours / 100.0);
      double degradation_rate = thermal_factor * photo_factor * 0.01;
      s->reagent_potency -= degradation_rate * (dt / 86400.0);
      s->reagent_potency = fmax(0.0, s->reagent_potency);
      s->reagent_age_days += dt / 86400.0;
    }
    final_value = clip(r.value * s->reagent_potency * 0.95 + n0 * 0.005, s->min_value, s->max_value);
  } else if (s->kind == K_FLOW) {                                                         /* flow_sensor.py:125-219 */
    if (s->variant == 1) {                                                                /* turbine :138-141, 180-199 */
      if (have_prev) s->bearing_wear_days += (dt / 86400.0) * (1.0 + inst[3] * 5.0);
      double friction_increase = 1.0 + 0.01 * (s->bearing_wear_days / 365.0);
      double friction_loss = (0.01 * friction_increase) * s->full_scale;
      double effective_value = r.value < friction_loss ? 0.0 : r.value - friction_loss;
      final_value = effective_value + n0 * (inst[3] * 0.01 * s->full_scale);
    } else {
      if (have_prev) s->electrode_fouling += 0.001 * (dt / 86400.0);
      double fouling_factor = fmax(0.9, 1.0 - 0.005 * s->electrode_fouling);
      final_value = r.value * fouling_factor * 1.0 + n0 * (0.001 * s->full_scale);
    }
    if (inst[1] > 0.0) {
      double ub, ub2;
      uniform2(g, 7, &ub, &ub2);
      if (ub < inst[1] / 60.0) final_value = 0.0;
    }
    if (final_value < 0.01 * s->full_scale) final_value = 0.0;
    final_value = clip(final_value, 0.0, s->max_value);
  } else {                                                                                /* temperature_sensor.py:110-171 */
    if (s->variant >= 2) {                                                                /* thermocouple :173-194 */
      double V_seebeck = 40.0 * (r.value - 25.0);
      s->cold_junction_drift += n0 * 0.01;
      double emf_noise = n1 * 0.5;
      double V_total = V_seebeck + emf_noise;
      final_value = (V_total / 40.0) + 25.0 + s->cold_junction_drift;
    } else {                                                                              /* RTD :150-171 */
      double R0 = s->variant == 1 ? 1000.0 : 100.0;
      double R_true = R0 * (1.0 + 0.00385 * r.value);
      double R_measured = R_true + 2.0 * 0.5;
      double I_A = 1.0 / 1000.0;
      double self_heating_error = 0.001 * ((I_A * I_A) * R_measured * 1000.0);
      double T_measured = (R_measured / R0 - 1.0) / 0.00385;
      final_value = T_measured + self_heating_error + n0 * 0.001;
    }
    double stem_error = 0.01 * (r.value - inst[4]);
    final_value += stem_error;
    final_value = clip(final_value, s->min_value, s->max_value);
    r.drift = r.drift + stem_error;
  }
  r.value = final_value;
  s->current_value = final_value;
  s->last_value = final_value;
  return r;
}

/* ---- batched drivers (AoS per plant): y [P][3n] species-major ---- */
void wt_oracle_sensors_init(int P, double t0, const double *cfg_flow, const double *cfg_cl, const double *cfg_T,
                            suite_t *st, int temp_kind, int flow_kind) {
  for (int p = 0; p < P; ++p) suite_init(&st[p], t0, cfg_flow[p], cfg_cl[p], cfg_T[p], temp_kind, flow_kind);
}

/* BaseSensor.reset (base_sensor.py:858-878) with the simulated time t in place of time.monotonic() */
void wt_oracle_sensors_reset(int P, int sensor, double t, suite_t *st) {
  for (int p = 0; p < P; ++p) {
    sensor_t *s = &st[p].s[sensor];
    s->current_value = (s->min_value + s->max_value) / 2.0;
    s->calibration_offset = 0.0;
    s->hist_count = 0;            /* reading_history.clear() */
    s->last_value = NAN;
    s->status = SS_NORMAL;
    s->fault = F_NONE;
    s->last_calibration_time = t;
    s->power_on_time = t;
    s->has_cal_record = 0;        /* calibration_history.clear(): every read reports CALIBRATION_EXPIRED until calibrate() */
    if (s->line >= 0) { st[p].line[s->line].head = 0; st[p].line[s->line].count = 0; }  /* delay_buffer.clear() */
  }
}

void wt_oracle_sensors_calibrate(int P, int sensor, double t, const double *ref, suite_t *st) {
  for (int p = 0; p < P; ++p) calibrate(&st[p].s[sensor], ref[p], t);
}

/* ---- maintenance operations -----------------------------------------------
 * op 0  pHSensor.calibrate_two_point(b1, b2, m1, m2, t)      ph_sensor.py:338-393   arg = {b1, b2, m1, m2}
 * op 1  pHSensor.clean_electrode(method, t)                  ph_sensor.py:395-434   arg[0] = 0 water_rinse, 1 acid_clean, 2 pepsin_clean
 * op 2  ChlorineSensor.replace_membrane(t)  (amperometric)   chlorine_sensor.py:486-509
 * op 3  ChlorineSensor.replace_reagent(t)   (DPD)            chlorine_sensor.py:511-537
 * slope_percentage / glass_etching are not carried: read() overwrites slope_percentage on every call
 * (ph_sensor.py:256-262) and nothing on the read path uses glass_etching.  Returns 0, or -1 where the
 * reference raises ValueError (wrong sensor kind, unknown cleaning method). */
static int maintain(sensor_t *s, int op, double t, const double *arg) {
  switch (op) {
    case 0:
      if (s->kind != K_PH) return -1;
      s->reference_contamination = 0.0;
      calibrate(s, (arg[0] + arg[1]) / 2.0, t);
      return 0;
    case 1: {
      if (s->kind != K_PH) return -1;
      const int m = (int)arg[0];
      if (m == 0) s->membrane_fouling *= 0.5;
      else if (m == 1) s->membrane_fouling *= 0.1;
      else if (m == 2) s->membrane_fouling *= 0.2;
      else return -1;
      s->days_since_cleaning = 0.0;
      s->power_on_time = t;
      return 0;
    }
    case 2:
      if (s->kind != K_CL_AMP) return -1;
      s->membrane_fouling = 0.0;
      s->membrane_age_days = 0.0;
      s->power_on_time = t;
      calibrate(s, 0.0, t);
      return 0;
    case 3:
      if (s->kind != K_CL_DPD) return -1;
      s->reagent_potency = 1.0;
      s->reagent_age_days = 0.0;
      s->light_exposure_hours = 0.0;
      calibrate(s, 0.0, t);
      return 0;
  }
  return -1;
}
int wt_oracle_sensors_maintain(int P, int sensor, int op, double t, const double *arg, suite_t *st) {
  int rc = 0;
  for (int p = 0; p < P; ++p) rc |= maintain(&st[p].s[sensor], op, t, arg);
  return rc;
}
/* state access for the tests: field 0 current_value 1 calibration_offset 2 last_calibration_time 3 power_on_time
 * 4 membrane_fouling 5 reference_contamination 6 days_since_cleaning 7 membrane_age_days 8 reagent_potency
 * 9 light_exposure_hours 10 reagent_age_days 11 status 12 fault */
static double *field_ptr(sensor_t *s, int f) {
  switch (f) {
    case 0: return &s->current_value; case 1: return &s->calibration_offset; case 2: return &s->last_calibration_time;
    case 3: return &s->power_on_time; case 4: return &s->membrane_fouling; case 5: return &s->reference_contamination;
    case 6: return &s->days_since_cleaning; case 7: return &s->membrane_age_days; case 8: return &s->reagent_potency;
    case 9: return &s->light_exposure_hours; case 10: return &s->reagent_age_days;
  }
  return 0;
}
void wt_oracle_sensors_poke(int P, int sensor, int field, const double *v, suite_t *st) {
  for (int p = 0; p < P; ++p) *field_ptr(&st[p].s[sensor], field) = v[p];
}
void wt_oracle_sensors_peek(int P, int sensor, double *out13, suite_t *st) {
  for (int p = 0; p < P; ++p) {
    for (int f = 0; f < 11; ++f) out13[p * 13 + f] = *field_ptr(&st[p].s[sensor], f);
    out13[p * 13 + 11] = st[p].s[sensor].status;
    out13[p * 13 + 12] = st[p].s[sensor].fault;
  }
}

typedef struct {
  int tid, nthreads, P, n;
  long long plant0;
  unsigned k;
  double t, t_prev;
  const double *y, *flow, *inst;
  suite_t *st;
  double *out;
  int32_t *status, *fault;
  uint64_t seed;
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (int p = j->tid; p < j->P; p += j->nthreads) {
    const double *yy = j->y + (size_t)p * 3 * j->n;
    unsigned long long gid = (unsigned long long)(j->plant0 + p);
    for (int s = 0; s < NSENS; ++s) {
      rng_t g = {(uint32_t)gid, j->k ^ ((uint32_t)(gid >> 32) << 24), (uint32_t)s, (uint32_t)j->seed, (uint32_t)(j->seed >> 32)};
      sensor_t *ss = &j->st[p].s[s];
      reading_t r = sensor_read(ss, &j->st[p], &g, j->inst, yy, yy + j->n, yy + 2 * j->n, j->n, j->flow[p],
                                j->t, j->t_prev, ss->hist_count > 0);
      ss->hist_count++;   /* every read appends to reading_history, also the warm-up / power-fault early returns */
      double *o = j->out + ((size_t)p * NSENS + s) * 5;
      o[0] = r.value; o[1] = r.raw_value; o[2] = r.noise; o[3] = r.drift; o[4] = r.uncertainty;
      j->status[(size_t)p * NSENS + s] = r.status;
      j->fault[(size_t)p * NSENS + s] = r.fault;
    }
  }
  return NULL;
}

void wt_oracle_sensors_read(int P, int n, long long plant0, unsigned read_index, double t, double t_prev, const double *y,
                            const double *flow, suite_t *st, double *out, int32_t *status, int32_t *fault,
                            const double *suite6, uint64_t seed, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  pthread_t th[64];
  job_t jobs[64];
  for (int i = 0; i < nthreads; ++i) {
    job_t j = {i, nthreads, P, n, plant0, read_index, t, t_prev, y, flow, suite6, st, out, status, fault, seed};
    jobs[i] = j;
    if (i > 0) pthread_create(&th[i], NULL, worker, &jobs[i]);
  }
  worker(&jobs[0]);
  for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
}
