// wt_sensors.cuh -- K3: the fused sensor-suite kernel.
//
// One thread per plant reads the 7 sensors of create_realistic_sensor_suite (reference
// src/wt_simulator/sensors/__init__.py:41-120) in the reference's dict order:
//   0 pH_inlet (zone 0, line A)   1 pH_outlet (zone n-1, line B)   2 chlorine_inlet (amperometric, zone 0)
//   3 chlorine_outlet (DPD, zone n-1)   4 flow_main (magnetic | turbine)   5 temp_inlet (line A)   6 temp_outlet (line B)
//   (temperature sensors: RTD PT100 as in the factory, or PT1000 / thermocouple K / J -- the suite configuration word)
// and runs, per sensor, BaseSensor.read (base_sensor.py:509-699) followed by the subclass
// post-processing (ph_sensor.py:216-336, chlorine_sensor.py:345-484, temperature_sensor.py:110-171,
// flow_sensor.py:125-219).  The reference's emergent behaviour is kept on purpose (SURVEY.md
// Appendix C / D): absorbing power and open/short-circuit faults, calibration offsets equal to the
// reference value, the pH and temperature sensors sharing one delay line, sensor-specific offsets
// fed back through the 0.5/0.5 lag filter, hysteresis being a no-op.
//
// Randomness: counter-based Philox4x32-10 keyed by the ensemble seed; counter = (GLOBAL plant id,
// read index, sensor*16 + block).  Results do not depend on how plants are sharded over GPUs and
// sensor state is resumable from (seed, read index).  The reference seeds numpy PCG64 from
// secrets.randbits (base_sensor.py:331), so only distributional parity with it is meaningful.
//
// State layout in HBM (fp64 / int32, plant index fastest):
//   sens[(field * 7 + sensor) * P + p]      field: WT_SF_*  (WT_NSF = 10 fields)
//   sens_i[(field * 7 + sensor) * P + p]    field: WT_SI_* (status, fault, len(reading_history), flags)
//   ring[((line * WT_RING + slot) * 2 + f) * P + p]   f: 0 timestamp, 1 value   (the stored sample
//       temperature of SampleLine.transport_sample is never used by read(): base_sensor.py:610-614)
//   ring_i[(line * 2 + f) * P + p]          f: 0 head (next write slot), 1 count
#pragma once

#include <math.h>
#include <stdint.h>

#define WT_NSENS 7
#define WT_RING 100       // SampleLine deque maxlen = max(100, int(delay) + 10), base_sensor.py:170-171
#define WT_SF_CUR 0       // current_value
#define WT_SF_VOLT 1      // supply_voltage
#define WT_SF_CALOFF 2    // calibration_offset
#define WT_SF_TCAL 3      // last_calibration_time == calibration record timestamp
#define WT_SF_LASTVAL 4   // reading_history[-1].value
#define WT_SF_AUX0 5      // pH: membrane_fouling | Cl amp: membrane_fouling | DPD: reagent_potency | flow: electrode_fouling (magnetic) / bearing_wear_days (turbine) | thermocouple: cold_junction_drift
#define WT_SF_AUX1 6      // pH: reference_contamination | Cl amp: membrane_age_days | DPD: light_exposure_hours
#define WT_SF_AUX2 7      // pH: days_since_cleaning | DPD: reagent_age_days
#define WT_SF_TPOWER 8   // power_on_time (differs from WT_SF_TCAL only after clean_electrode)
#define WT_SF_AUX3 9     // pH: slope_percentage (only refreshed while a calibration record exists, ph_sensor.py:274-279)
#define WT_NSF 10
#define WT_SI_STATUS 0
#define WT_SI_FAULT 1
#define WT_SI_HIST 2      // len(reading_history): reads since construction / reset()
#define WT_SI_FLAGS 3     // bit 0: calibration_history is empty (after reset(), until the next calibrate())
#define WT_NSI 4
#define WT_SO_VALUE 0     // outputs: out[(field * 7 + sensor) * P + p]
#define WT_SO_RAW 1
#define WT_SO_NOISE 2
#define WT_SO_DRIFT 3
#define WT_SO_UNC 4
#define WT_NSO 5

// SensorStatus / SensorFault enum order of base_sensor.py:49-75
enum { SS_NORMAL = 0, SS_CALIBRATING, SS_WARMING_UP, SS_FAILED, SS_SATURATED, SS_DRIFT_WARNING, SS_CAL_EXPIRED,
       SS_OPEN_CIRCUIT, SS_SHORT_CIRCUIT, SS_OUT_OF_RANGE, SS_POWER_FAULT, SS_RATE_FAULT };
enum { SFLT_NONE = 0, SFLT_OPEN, SFLT_SHORT, SFLT_RANGE, SFLT_RATE, SFLT_POWER_LOW, SFLT_POWER_HIGH };
enum { ST_PH = 0, ST_CL_AMP, ST_CL_DPD, ST_FLOW_MAG, ST_TEMP_RTD };

enum { WT_TEMP_RTD_PT100 = 0, WT_TEMP_RTD_PT1000 = 1, WT_TEMP_TC_K = 2, WT_TEMP_TC_J = 3 };  // TemperatureSensorType, temperature_sensor.py:29-35
enum { WT_FLOW_MAGNETIC = 0, WT_FLOW_TURBINE = 1 };                                           // FlowSensorType, flow_sensor.py:33-37
struct WtSuiteCfg {   // InstallationQuality of the suite (sensors/__init__.py:53-59) + line delay + sensor variants
  double flow_velocity, bubble_per_min, grounding, vibration_g, ambient_temp, line_delay_s;
  int temp_kind, flow_kind;
  uint32_t seed_lo, seed_hi;
};

struct SensorArgs {
  int P, n;
  long long plant0;      // global id of this shard's first plant (RNG counter)
  unsigned read_index;   // k: number of suite reads done before this one
  double t, t_prev;      // current_time of this read, of the previous read
  const double *clock;   // optional device clock {t, t_prev, read_index, t0, dt} overriding the three values above
                         // (CUDA-graph replays: the launch arguments are frozen, the clock is advanced by wt_clock_tick)
  const double *y;       // plant state [3][n][P]
  const double *flow;    // state.flow_rate [P]
  const double *cfg_flow, *cfg_cl, *cfg_T;  // per-plant configuration values (full scale, calibration references)
  double *sens; int *sens_i; double *ring; int *ring_i;
  double *out; int *out_status; int *out_fault;
  WtSuiteCfg s;
};

// ---- Philox4x32-10 (Salmon et al., SC'11) ------------------------------------------------------
__host__ __device__ inline void wt_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                          uint32_t *o) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
// two uniforms in [0,1) with 53 bits each
__host__ __device__ inline void wt_uniform2(const uint32_t *o, double *u0, double *u1) {
  *u0 = (double)((((uint64_t)(o[0] >> 5)) << 26) | (o[1] >> 6)) * (1.0 / 9007199254740992.0);
  *u1 = (double)((((uint64_t)(o[2] >> 5)) << 26) | (o[3] >> 6)) * (1.0 / 9007199254740992.0);
}

struct WtRng {
  uint32_t c0, c1, sens, k0, k1;
  __device__ void block(int b, uint32_t *o) const { wt_philox(c0, c1, (uint32_t)(sens * 16 + b), 0x57544232u, k0, k1, o); }
  __device__ void normal2(int b, double *z0, double *z1) const {  // Box-Muller
    uint32_t o[4];
    double u0, u1;
    block(b, o);
    wt_uniform2(o, &u0, &u1);
    const double r = sqrt(-2.0 * log(1.0 - u0));   // 1 - u0 in (0, 1]
    double s, c;
    sincos(6.283185307179586 * u1, &s, &c);
    *z0 = r * c;
    *z1 = r * s;
  }
  __device__ void uniform2(int b, double *u0, double *u1) const {
    uint32_t o[4];
    block(b, o);
    wt_uniform2(o, u0, u1);
  }
};

__host__ __device__ __forceinline__ int wt_sensor_type(int s) { return s < 2 ? ST_PH : (s == 2 ? ST_CL_AMP : (s == 3 ? ST_CL_DPD : (s == 4 ? ST_FLOW_MAG : ST_TEMP_RTD))); }

// SampleLine.transport_sample (base_sensor.py:177-216): append, then the buffered sample whose
// timestamp is nearest to t - delay; ties keep the FIRST (oldest) entry (strict '<').
__device__ double wt_transport_sample(const SensorArgs &a, int p, int line, double value, double t, double dt_hint) {
  const size_t P = (size_t)a.P;
  int *hd = a.ring_i + ((size_t)line * 2 + 0) * P + p, *ct = a.ring_i + ((size_t)line * 2 + 1) * P + p;
  int head = *hd, count = *ct;
  double *base = a.ring + (size_t)line * WT_RING * 2 * P + p;
  base[((size_t)head * 2 + 0) * P] = t;
  base[((size_t)head * 2 + 1) * P] = value;
  head = head + 1 == WT_RING ? 0 : head + 1;
  count = count < WT_RING ? count + 1 : WT_RING;
  *hd = head;
  *ct = count;
  const double target = t - a.s.line_delay_s;
  // The reference scans the deque from the oldest entry and keeps the FIRST minimum of |timestamp - target|.
  // Timestamps never decrease along the deque (read() rejects a decreasing current_time, base_sensor.py:543-549), so
  // the distance falls to its minimum and rises again: scanning from the NEWEST entry backwards, taking every tie
  // (-> the oldest of equal minima) and stopping at the first larger distance finds the same slot after reading
  // ~delay/dt + 2 timestamps instead of all 100 (the kernel is bound by this HBM traffic).  The scan goes in chunks of
  // 16 independent loads: one load per iteration with a data-dependent exit made the kernel latency-bound (measured:
  // 1.7x slower than the full scan it replaced).
  int slot = head == 0 ? WT_RING - 1 : head - 1;  // newest entry
  // Fast path: reads normally come at a steady interval, so the wanted sample sits ~delay / interval entries behind the
  // newest one.  ONE round of 8 independent loads around that guess finds it whenever the minimum of the (unimodal)
  // distance lies strictly inside the window -- or at a window edge that is also an end of the deque (checked against a
  // brute-force first-minimum on 400,000 random non-decreasing sequences with plateaus, tests/test_sensor_search.py); anything else
  // (irregular read times, a window on a slope, the first reads) falls through to the scan below.  j counts entries
  // back from the newest; the reference's FIRST minimum in deque order is the LARGEST such j.
  // Two guesses: two entries per read (both sample lines of the suite are shared by a pH and a temperature sensor,
  // sensors/__init__.py:62-67), then one (a line whose other sensor is dead -- absorbing power fault, open / short circuit
  // -- or still warming up gets one entry per read).  A wrong guess only costs the fall-through.
#pragma unroll 1
  for (int per_read = 2; per_read >= 1 && dt_hint > 0.0 && count > 0; --per_read) {
    const double q = (double)per_read * a.s.line_delay_s / dt_hint;
    int jg = q < (double)(count - 1) ? (int)(q + 0.5) : count - 1;
    int ja = jg - 3;
    ja = ja < 0 ? 0 : ja;
    int jb = ja + 7;
    jb = jb > count - 1 ? count - 1 : jb;
    double d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int sj = slot - (ja + k);
      sj = sj < 0 ? sj + WT_RING : sj;
      d[k] = (ja + k <= jb) ? fabs(base[((size_t)sj * 2) * P] - target) : INFINITY;
    }
    double dmin = d[0];
    int m = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (d[k] <= dmin) { dmin = d[k]; m = k; }   // ties: the larger j (the older entry)
    const int jm = ja + m;
    // newer edge: strictly larger than the minimum (a tie there -- the deque holds equal timestamps in pairs -- could be a
    // plateau that goes on falling beyond the window); older edge: the minimum (its LARGEST j) strictly inside
    if ((ja == 0 || d[0] > dmin) && (jm < jb || jb == count - 1) && dmin < INFINITY) {
      int sj = slot - jm;
      sj = sj < 0 ? sj + WT_RING : sj;
      return base[((size_t)sj * 2 + 1) * P];
    }
  }
  int best = slot, remaining = count;
  double best_d = INFINITY;
  bool done = false;
  while (remaining > 0 && !done) {
    const int m = remaining < 16 ? remaining : 16;
    double ts[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      int sj = slot - j;
      sj = sj < 0 ? sj + WT_RING : sj;
      ts[j] = j < m ? base[((size_t)sj * 2) * P] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < m && !done) {
        const double d = fabs(ts[j] - target);
        if (d > best_d) done = true;
        else {
          int sj = slot - j;
          best = sj < 0 ? sj + WT_RING : sj;
          best_d = d;
        }
      }
    }
    slot -= m;
    slot = slot < 0 ? slot + WT_RING : slot;
    remaining -= m;
  }
  return base[((size_t)best * 2 + 1) * P];
}

#ifndef WT_SENS_MINBLOCKS
#define WT_SENS_MINBLOCKS 8   // 64 registers, 32 warps per SM: the kernel is latency-bound (measured 0.28 / 0.26 / 0.24 ms at 4 / 6 / 8 blocks per SM)
#endif
__global__ void __launch_bounds__(128, WT_SENS_MINBLOCKS) wt_sensors_read_kernel(SensorArgs a) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.P) return;
  const size_t P = (size_t)a.P;
  const int n = a.n;
  const double t = a.clock ? a.clock[0] : a.t;
  const double dt_read = t - (a.clock ? a.clock[1] : a.t_prev);   // reading.timestamp - reading_history[-2].timestamp
  const unsigned read_index = a.clock ? (unsigned)a.clock[2] : a.read_index;
  const double FS = a.cfg_flow[p] * 2.0;          // sensors/__init__.py:104
  const unsigned long long gid = (unsigned long long)(a.plant0 + p);

  // blockIdx.y selects one of five independent work items of the plant -- {pH_inlet, temp_inlet} and {pH_outlet,
  // temp_outlet} (each pair shares a sample line and must be read in the reference's dict order), chlorine_inlet,
  // chlorine_outlet, flow_main -- instead of one thread walking all seven sensors: the kernel is bound by the latency
  // of its dependent HBM accesses (ncu: 80 % long-scoreboard stalls, 16 % DRAM utilisation with one thread per plant),
  // so five times the threads in flight is what speeds it up.  The Philox counter is (plant, read, sensor): no order.
  const int item = blockIdx.y;
  const int s_first = item, s_second = item < 2 ? item + 5 : -1;   // sensors 0..4, then 5 / 6 on the lines of 0 / 1
  for (int pass = 0; pass < 2; ++pass) {
    const int s = pass == 0 ? s_first : s_second;
    if (s < 0) break;
    const int type = wt_sensor_type(s);
    const int zone = (s == 0 || s == 2 || s == 5) ? 0 : n - 1;
    const int line = (s == 0 || s == 5) ? 0 : ((s == 1 || s == 6) ? 1 : -1);
    // static descriptor (ph_sensor.py:73-96, chlorine_sensor.py:84-117, temperature_sensor.py:57-85, flow_sensor.py:52-78)
    double vmin = 0.0, vmax, prec, drift_rate, warmup, max_rate, validity_h;
    const bool turbine = a.s.flow_kind == WT_FLOW_TURBINE, thermocouple = a.s.temp_kind >= WT_TEMP_TC_K;
    if (type == ST_PH) { vmax = 14.0; prec = 0.01; drift_rate = 0.01 / 24.0; warmup = 1800.0; max_rate = 0.5; validity_h = 24.0; }
    else if (type == ST_CL_AMP) { vmax = 10.0; prec = 0.01; drift_rate = 0.02 / 24.0; warmup = 300.0; max_rate = 1.0; validity_h = 24.0; }
    else if (type == ST_CL_DPD) { vmax = 10.0; prec = 0.02; drift_rate = 0.02 / 24.0; warmup = 60.0; max_rate = 1.0; validity_h = 24.0; }
    else if (type == ST_FLOW_MAG) { vmax = FS; prec = (turbine ? 0.01 : 0.005) * FS; drift_rate = 0.0; warmup = 10.0; max_rate = FS; validity_h = 8760.0; }
    else { vmin = -10.0; vmax = 110.0; prec = thermocouple ? 0.5 : 0.1; drift_rate = 0.0; warmup = 30.0; max_rate = 10.0; validity_h = 8760.0; }

    double *S = a.sens + (size_t)s * P + p;      // + field * 7 * P
#define SF(f) S[(size_t)(f) * WT_NSENS * P]
    int *SI = a.sens_i + (size_t)s * P + p;
    double *O = a.out + (size_t)s * P + p;
#define OUT(f) O[(size_t)(f) * WT_NSENS * P]
    WtRng rng;
    rng.c0 = (uint32_t)gid; rng.c1 = read_index ^ ((uint32_t)(gid >> 32) << 24); rng.sens = (uint32_t)s;
    rng.k0 = a.s.seed_lo; rng.k1 = a.s.seed_hi;

    double cur = SF(WT_SF_CUR);
    const double volt0 = SF(WT_SF_VOLT);
    int status = SI[0], fault = SI[(size_t)WT_SI_FAULT * WT_NSENS * P];
    // every read appends to reading_history, also the early returns below (base_sensor.py:570-595)
    const int hist = SI[(size_t)WT_SI_HIST * WT_NSENS * P];
    SI[(size_t)WT_SI_HIST * WT_NSENS * P] = hist + 1;
    const bool have_prev = hist > 0;               // len(reading_history) >= 1 before this read
    const bool has_cal = !(SI[(size_t)WT_SI_FLAGS * WT_NSENS * P] & 1);

    // base_sensor.py:556-577: the voltage drawn by the PREVIOUS read is checked first; a sensor that
    // ever drew V outside (20, 28) never redraws it -> absorbing
    if (!(20.0 < volt0 && volt0 < 28.0)) {
      OUT(WT_SO_VALUE) = nan(""); OUT(WT_SO_RAW) = nan(""); OUT(WT_SO_NOISE) = 0.0; OUT(WT_SO_DRIFT) = 0.0; OUT(WT_SO_UNC) = 0.0;
      a.out_status[(size_t)s * P + p] = SS_POWER_FAULT;
      a.out_fault[(size_t)s * P + p] = volt0 < 20.0 ? SFLT_POWER_LOW : SFLT_POWER_HIGH;
      SF(WT_SF_LASTVAL) = nan("");
      continue;
    }
    double z0, z1;
    rng.normal2(0, &z0, &z1);
    const double volt = 24.0 + z0 * 1.0;          // :579
    SF(WT_SF_VOLT) = volt;
    const double tcal = SF(WT_SF_TCAL);
    if (!(t - SF(WT_SF_TPOWER) >= warmup)) {      // :582-595
      OUT(WT_SO_VALUE) = nan(""); OUT(WT_SO_RAW) = nan(""); OUT(WT_SO_NOISE) = 0.0; OUT(WT_SO_DRIFT) = 0.0; OUT(WT_SO_UNC) = 0.0;
      a.out_status[(size_t)s * P + p] = SS_WARMING_UP;
      a.out_fault[(size_t)s * P + p] = SFLT_NONE;
      SF(WT_SF_LASTVAL) = nan("");
      continue;
    }
    // :598-600, _check_calibration_valid :432-436 (no record -> invalid), CalibrationRecord.is_expired
    const bool cal_expired = !has_cal || ((t - tcal) / 3600.0) > validity_h;
    if (cal_expired) status = SS_CAL_EXPIRED;

    // _get_true_value
    const double pHz = a.y[((size_t)0 * n + zone) * P + p], Clz = a.y[((size_t)1 * n + zone) * P + p],
                 Tz = a.y[((size_t)2 * n + zone) * P + p];
    double truev;
    if (type == ST_PH) truev = pHz + 0.003 * (Tz - 25.0);                               // ph_sensor.py:162-180
    else if (type == ST_CL_AMP || type == ST_CL_DPD) {                                    // chlorine_sensor.py:202-227
      const double ratio = exp10(7.5 - pHz);
      truev = Clz * (0.5 + 0.5 * (ratio / (1.0 + ratio)));
    } else if (type == ST_FLOW_MAG) truev = a.flow[p];
    else truev = Tz;
    if (line >= 0) truev = wt_transport_sample(a, p, line, truev, t, dt_read);            // :603-614

    const double drift = drift_rate * ((t - tcal) / 3600.0) + SF(WT_SF_CALOFF);          // :617-620
    const double noise = z1 * prec;                                                       // :623
    cur = 0.5 * (truev + noise + drift) + (1.0 - 0.5) * cur;                              // :626-630
    // _apply_hysteresis(self.current_value): direction = sign(value - self.current_value) = 0 -> no-op (:438-462)
    // _apply_installation_effects (:464-507)
    if (a.s.flow_velocity < 0.1 || a.s.grounding < 0.8 || a.s.vibration_g > 0.2 || a.s.bubble_per_min > 0.0) {
      double g0, g1, g2, g3, ub, ub2;
      rng.normal2(5, &g0, &g1);
      rng.normal2(6, &g2, &g3);
      rng.uniform2(4, &ub, &ub2);
      if (a.s.flow_velocity < 0.1) cur += g0 * (prec * 2.0);
      if (a.s.bubble_per_min > 0.0 && ub < a.s.bubble_per_min / 60.0) cur = nan("");
      else {
        if (a.s.grounding < 0.8) cur += g1 * (prec * (2.0 - a.s.grounding));
        if (a.s.vibration_g > 0.2) cur += g2 * (a.s.vibration_g * prec);
      }
    }
    const double lastv = SF(WT_SF_LASTVAL);
    double rate = 0.0;                                                                     // :638-646
    if (have_prev && dt_read > 0.0 && isfinite(lastv)) rate = (cur - lastv) / dt_read;

    // _check_for_faults (:357-409), first match wins
    int f = SFLT_NONE;
    const double span = vmax - vmin;
    if (!(20.0 < volt && volt < 28.0)) f = volt < 20.0 ? SFLT_POWER_LOW : SFLT_POWER_HIGH;
    else if (cur < vmin - 0.1 * span || cur > vmax + 0.1 * span) f = SFLT_RANGE;
    else if (fabs(rate) > max_rate) f = SFLT_RATE;
    else {
      double u0, u1;
      rng.uniform2(1, &u0, &u1);
      if (u0 < 0.0001) f = u1 < 0.5 ? SFLT_OPEN : SFLT_SHORT;
    }
    if (f != SFLT_NONE) {                                                                  // :651-662
      fault = f;
      if (f == SFLT_OPEN || f == SFLT_SHORT) { status = SS_FAILED; cur = nan(""); }
      else if (f == SFLT_RANGE) status = SS_OUT_OF_RANGE;
      else if (f == SFLT_POWER_LOW || f == SFLT_POWER_HIGH) status = SS_POWER_FAULT;
      else status = SS_RATE_FAULT;
    } else {                                                                               // :663-682
      fault = SFLT_NONE;
      if (!isnan(cur)) {
        const double b = fmin(fmax(cur, vmin), vmax);
        if (b != cur) status = SS_SATURATED;
        else if (!cal_expired) status = SS_NORMAL;
        cur = b;
      }
      if (fabs(drift) > 0.1 * span && status != SS_CAL_EXPIRED) status = SS_DRIFT_WARNING;
    }
    double value = cur, o_noise = noise, o_drift = drift, unc = prec * 2.0;

    // ---- subclass post-processing (only on a finite base reading) ----
    if (isfinite(value)) {
      double n0, n1;
      rng.normal2(2, &n0, &n1);
      if (type == ST_PH) {
        double foul = SF(WT_SF_AUX0), contam = SF(WT_SF_AUX1);
        if (have_prev) {                                                                   // ph_sensor.py:236-238, 182-214
          const double bio = foul > 0.05 ? 0.1 * exp(0.05 * (Tz - 25.0)) : 0.001;
          const double scal = 100.0 * (a.s.flow_velocity < 0.1 ? 0.0001 : 0.00001);
          foul = fmin(1.0, foul + (bio + scal) * (dt_read / 86400.0));
          SF(WT_SF_AUX2) += dt_read / 86400.0;
        }
        const double en = n0 * (0.002 * (1.0 + 0.1 * fabs(value - 7.0)));                  // :242-246
        const double jn = n1 * (0.005 * (1.0 + contam));                                   // :249-253
        double days = 0.0, slope_pct = SF(WT_SF_AUX3);                                      // :274-279
        if (has_cal) {
          days = (t - tcal) / 86400.0;
          slope_pct = fmax(90.0, 100.0 - 0.001 * days);
          SF(WT_SF_AUX3) = slope_pct;
        }
        double slope_err = 0.0;                                                             // :265-273
        if (!(4.0 < value && value < 7.0)) slope_err = fmin(fabs(value - 4.0), fabs(value - 7.0)) * (100.0 - slope_pct) / 100.0;
        const double foul_off = foul * 0.2;                                                 // :276-277
        double f0, f1;
        rng.normal2(3, &f0, &f1);
        const double fn = f0 * (foul * 0.05);
        contam = fmin(0.5, contam + 0.0001 * (days / 30.0));                                // :280-282
        const double ref_off = contam * 0.1;
        double fin = value + en + jn + slope_err + foul_off + fn + ref_off;                 // :285-293
        fin = fmin(fmax(fin, vmin), vmax);
        o_noise = noise + en + jn + fn;
        o_drift = drift + slope_err + foul_off + ref_off;
        unc = prec * 3.0;
        value = fin;
        SF(WT_SF_AUX0) = foul; SF(WT_SF_AUX1) = contam;
      } else if (type == ST_CL_AMP) {                                                       // chlorine_sensor.py:319-343, 405-449
        double foul = SF(WT_SF_AUX0), age = SF(WT_SF_AUX1);
        if (have_prev) {
          foul = fmin(1.0, foul + (a.s.flow_velocity < 0.1 ? 0.05 : 0.01) * (dt_read / 86400.0));
          age += dt_read / 86400.0;
        }
        const double pn = n0 * (0.005 * (1.0 + age / 365.0)), dn = n1 * 0.003;
        double fin = (value + 0.0) * (1.0 - 0.8 * foul) + pn + dn;
        value = fmin(fmax(fin, vmin), vmax);
        SF(WT_SF_AUX0) = foul; SF(WT_SF_AUX1) = age;
      } else if (type == ST_CL_DPD) {                                                       // :280-317, 451-484
        double pot = SF(WT_SF_AUX0), light = SF(WT_SF_AUX1);
        if (have_prev) {
          const double thermal = exp((50000.0 / 8.314) * (1.0 / 293.15 - 1.0 / (20.0 + 273.15)));
          light += dt_read / 3600.0;
          const double photo = 1.0 + 0.1 * (light / 100.0);
          pot = fmax(0.0, pot - thermal * photo * 0.01 * (dt_read / 86400.0));
          SF(WT_SF_AUX2) += dt_read / 86400.0;
        }
        double fin = value * pot * 0.95 + n0 * 0.005;
        value = fmin(fmax(fin, vmin), vmax);
        SF(WT_SF_AUX0) = pot; SF(WT_SF_AUX1) = light;
      } else if (type == ST_FLOW_MAG) {                                                     // flow_sensor.py:131-178, 201-219
        double foul = SF(WT_SF_AUX0), fin;   // electrode_fouling (magnetic) or bearing_wear_days (turbine)
        if (turbine) {                                                                     // :138-141, 180-199
          if (have_prev) foul += (dt_read / 86400.0) * (1.0 + a.s.vibration_g * 5.0);
          const double friction_loss = (0.01 * (1.0 + 0.01 * (foul / 365.0))) * FS;
          const double eff = value < friction_loss ? 0.0 : value - friction_loss;
          fin = eff + n0 * (a.s.vibration_g * 0.01 * FS);
        } else {
          if (have_prev) foul += 0.001 * (dt_read / 86400.0);
          fin = value * fmax(0.9, 1.0 - 0.005 * foul) * 1.0 + n0 * (0.001 * FS);
        }
        if (a.s.bubble_per_min > 0.0) {
          double ub, ub2;
          rng.uniform2(7, &ub, &ub2);
          if (ub < a.s.bubble_per_min / 60.0) fin = 0.0;
        }
        if (fin < 0.01 * FS) fin = 0.0;
        value = fmin(fmax(fin, 0.0), vmax);
        SF(WT_SF_AUX0) = foul;
      } else {                                                                              // temperature_sensor.py:116-171
        double fin;
        if (thermocouple) {                                                                 // :173-194
          const double v_seebeck = 40.0 * (value - 25.0);
          const double cjd = SF(WT_SF_AUX0) + n0 * 0.01;
          SF(WT_SF_AUX0) = cjd;
          const double v_total = v_seebeck + n1 * 0.5;
          fin = (v_total / 40.0) + 25.0 + cjd;
        } else {                                                                            // RTD :150-171
          const double R0 = a.s.temp_kind == WT_TEMP_RTD_PT1000 ? 1000.0 : 100.0;
          const double R_true = R0 * (1.0 + 0.00385 * value);
          const double R_meas = R_true + 2.0 * 0.5;
          const double I_A = 1.0 / 1000.0;
          const double she = 0.001 * (((I_A * I_A) * R_meas) * 1000.0);
          fin = (R_meas / R0 - 1.0) / 0.00385 + she + n0 * 0.001;
        }
        const double stem = 0.01 * (value - a.s.ambient_temp);
        fin += stem;
        o_drift = drift + stem;
        value = fmin(fmax(fin, vmin), vmax);
      }
      cur = value;   // self.current_value = final_value: fed back into the next read's lag filter
    }
    SF(WT_SF_CUR) = cur;
    SF(WT_SF_LASTVAL) = value;
    SI[0] = status;
    SI[(size_t)WT_SI_FAULT * WT_NSENS * P] = fault;
    OUT(WT_SO_VALUE) = value; OUT(WT_SO_RAW) = truev; OUT(WT_SO_NOISE) = o_noise; OUT(WT_SO_DRIFT) = o_drift; OUT(WT_SO_UNC) = unc;
    a.out_status[(size_t)s * P + p] = status;
    a.out_fault[(size_t)s * P + p] = fault;
#undef SF
#undef OUT
  }
}

// create_realistic_sensor_suite + __main__.initialize_sensors (calibrate at t0): initial state
__global__ void wt_sensors_init_kernel(int P, double t0, const double *cfg_flow, const double *cfg_cl, const double *cfg_T,
                                       double *sens, int *sens_i, int *ring_i) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const size_t Pz = (size_t)P;
  for (int s = 0; s < WT_NSENS; ++s) {
    const int type = wt_sensor_type(s);
    // constructor current_value (ph_sensor.py:110, chlorine_sensor.py:150, flow_sensor.py:89, temperature_sensor.py:101)
    const double cur0 = type == ST_PH ? 7.0 : (type == ST_TEMP_RTD ? 20.0 : 0.0);
    // calibrate(reference, t0): offset = reference - current_value (base_sensor.py:722-733)
    const double ref = type == ST_PH ? 7.0 : (type == ST_TEMP_RTD ? cfg_T[p] : (type == ST_FLOW_MAG ? cfg_flow[p] : cfg_cl[p]));
    double *S = sens + (size_t)s * Pz + p;
    S[(size_t)WT_SF_CUR * WT_NSENS * Pz] = cur0;
    S[(size_t)WT_SF_VOLT * WT_NSENS * Pz] = 24.0;
    S[(size_t)WT_SF_CALOFF * WT_NSENS * Pz] = ref - cur0;
    S[(size_t)WT_SF_TCAL * WT_NSENS * Pz] = t0;
    S[(size_t)WT_SF_TPOWER * WT_NSENS * Pz] = t0;
    S[(size_t)WT_SF_LASTVAL * WT_NSENS * Pz] = nan("");
    S[(size_t)WT_SF_AUX0 * WT_NSENS * Pz] = type == ST_CL_DPD ? 1.0 : 0.0;  // reagent_potency = 1
    S[(size_t)WT_SF_AUX1 * WT_NSENS * Pz] = 0.0;
    S[(size_t)WT_SF_AUX2 * WT_NSENS * Pz] = 0.0;
    S[(size_t)WT_SF_AUX3 * WT_NSENS * Pz] = 100.0;                          // slope_percentage, ph_sensor.py:137
    sens_i[(size_t)s * Pz + p] = SS_NORMAL;
    sens_i[((size_t)WT_SI_FAULT * WT_NSENS + s) * Pz + p] = SFLT_NONE;
    sens_i[((size_t)WT_SI_HIST * WT_NSENS + s) * Pz + p] = 0;
    sens_i[((size_t)WT_SI_FLAGS * WT_NSENS + s) * Pz + p] = 0;
  }
  for (int k = 0; k < 4; ++k) ring_i[(size_t)k * Pz + p] = 0;
}
