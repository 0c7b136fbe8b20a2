/*
 * wt_b200.h -- C ABI of the B200-native batched plant-stepping engine.
 *
 * Drop-in boundary for the data-parallel hot path of wt_simulator.core / wt_simulator.sensors
 * (reference: Guivernoir/ICS-WT-PhysicsEngine).  The reference has no FFI: its boundary is the
 * Python object API.  Each entry point below names the reference interface it stands in for;
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, every function returns 0 on success or a negative
 *     WT_ERR_* code (a positive value is a cudaError_t).  No exceptions cross the boundary.
 *   - "dev" pointers are device pointers of the current CUDA device; `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream).  Calls are stream-ordered and
 *     asynchronous; buffers are borrowed for the duration of the enqueued work only.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Data layout in HBM (fp64, structure of arrays, plant index fastest):
 *   state   y[(var * n_zones + zone) * P + p]     var: 0 pH, 1 chlorine [mg/L], 2 temperature [C]
 *   params  par[k * P + p]                        k: WT_PAR_*   (derived per-plant constants)
 *   bnd     bnd[k * bnd_stride + p]               k: WT_BND_*   bnd_stride = P, or 0 to broadcast
 *   time[p], flow_rate[p], status[p] (uint32 bit mask WT_ST_*), counters[k * P + p] (int32)
 */
#ifndef WT_B200_H
#define WT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WT_ABI_VERSION 4   /* 3: wt_step / wt_advance take a device workspace (wt_step_workspace_bytes); new entry points of round 2
                              4: wt_catch_up takes floor_div (floor mode), WT_ST_DEGRADED */
#define WT_MAX_ZONES 32

/* derived per-plant constants; computed on the host exactly as the reference constructors do
 * (reactor.py:228-270, chemistry.py:116-132, transport.py:202-254, 282-290) */
enum {
  WT_PAR_KW = 0, WT_PAR_KA1, WT_PAR_KA2, WT_PAR_KACL, WT_PAR_CT, WT_PAR_KX, WT_PAR_V, WT_PAR_ZH,
  WT_PAR_VZL, WT_PAR_VOLUME, WT_PAR_AT, WT_PAR_STRAT, WT_NPAR
};
/* BoundaryConditions fields in declaration order (reactor.py:168-186) */
enum {
  WT_BND_INLET_FLOW = 0, WT_BND_INLET_PH, WT_BND_INLET_CL, WT_BND_INLET_T, WT_BND_ACID_FLOW,
  WT_BND_ACID_CONC, WT_BND_CL_FLOW, WT_BND_CL_CONC, WT_BND_AMBIENT_T, WT_BND_HEAT_LOSS, WT_NBND
};
/* per-plant status bits written by wt_step (what the reference signals by logging / raising) */
enum {
  WT_ST_SOLVER_FAILED = 1u << 0,   /* solve_ivp status -1; reactor.py:486-487 only logs */
  WT_ST_T_RANGE = 1u << 1,         /* ValueError from celsius_to_kelvin (thermodynamics.py:146-157)
                                      inside the solve: state and time left untouched; plant HALTS */
  WT_ST_CLIP_PH = 1u << 2,         /* reactor.py:528-531 */
  WT_ST_CLIP_CL = 1u << 3,         /* reactor.py:533-536 */
  WT_ST_CLIP_T = 1u << 4,          /* reactor.py:538-541 */
  WT_ST_NONFINITE = 1u << 5,
  WT_ST_T_RANGE_DERIVED = 1u << 6, /* ValueError in _update_derived_state (reactor.py:521-524) */
  WT_ST_WORK_LIMIT = 1u << 7,      /* engine policy, not reference behaviour: attempt budget
                                      exhausted, state untouched; plant HALTS (unless deferred, below) */
  WT_ST_DEFERRED = 1u << 8,        /* the plant ran out of budget and is being caught up by wt_catch_up with a larger
                                      one: ordinary launches, statistics and sensors' consumers pass over it until
                                      wt_defer_rejoin */
  WT_ST_DEGRADED = 1u << 9         /* engine policy, not reference behaviour: this step was completed by a floor-mode
                                      catch-up (wt_catch_up with floor_div > 0) that forced an acceptance at the
                                      step-size floor; the state is a bounded-step continuation, not the reference's
                                      adaptive solution (which needs 1e5 .. 1e7 evaluations there) */
};
#define WT_ST_HALT_MASK (WT_ST_T_RANGE | WT_ST_WORK_LIMIT)
#define WT_ST_SKIP_MASK (WT_ST_HALT_MASK | WT_ST_DEFERRED)
/* solver path counters, accumulated per plant */
enum {
  WT_CNT_NFEV = 0, WT_CNT_NJEV, WT_CNT_NLU, WT_CNT_NSTEPS, WT_CNT_NNEWTON, WT_CNT_NREJECT,
  WT_CNT_NNEWTON_FAIL, WT_CNT_JAC_RETRY, WT_NCNT
};

enum {
  WT_OK = 0,
  WT_ERR_BAD_ARG = -1,
  WT_ERR_NO_DEVICE = -2,
  WT_ERR_ALLOC = -3
};

/* Library identification / device probe.  wt_device_count returns the number of CUDA devices
 * (0 when there is none -- the compute entry points then return WT_ERR_NO_DEVICE). */
int wt_abi_version(void);
int wt_device_count(void);
const char *wt_last_error(void);

/* ---------------------------------------------------------------------------------------
 * IntegratedCSTR.step(dt, boundary) for P plants          (reactor.py:450-541)
 *
 * Advances every non-halted plant by one step(dt): scipy-Radau solve over
 * [time[p], time[p]+dt], derived-state update, bound clipping.  In place on y/time/flow_rate.
 * A step is two launches on `stream`: the once-per-step solver set-up (f0, initial step size, first
 * finite-difference Jacobian) and the attempt loop run by persistent warps that draw groups of plants from a queue.
 *   workspace  device memory of wt_step_workspace_bytes(P, n_zones) bytes, owned by the caller and reusable by
 *              every later call with the same or a smaller P (the queue counter and the hand-off rows between the
 *              two launches); calls that may overlap on the device (different streams) need their own.
 *   derived    optional [3 * n_zones * P]: H_concentration, density, chlorine_decay_rate
 *              (ReactorState derived fields, reactor.py:136-147), may be NULL
 *   status     [P] in/out: plants whose word has a WT_ST_HALT_MASK bit are skipped
 *   counters   optional [WT_NCNT * P], accumulated, may be NULL
 *   max_attempts  budget of collocation solves per plant-step; 0 = no budget (reference behaviour), up to a hard
 *              stop at 2,000,000 collocation solves of one step (the plant then gets WT_ST_WORK_LIMIT as with any budget)
 * ------------------------------------------------------------------------------------- */
int wt_step(int P, int n_zones, double dt, const double *par_dev, const double *bnd_dev,
            int bnd_stride, double *time_dev, double *y_dev, double *flow_rate_dev,
            double *derived_dev, uint32_t *status_dev, int32_t *counters_dev, int max_attempts,
            void *workspace_dev, void *stream);
size_t wt_step_workspace_bytes(int P, int n_zones);

/* Same, `n_steps` consecutive step(dt) calls queued by one call (2 n_steps launches).
 * Equivalent to calling wt_step n_steps times with unchanged boundary conditions, except that the non-halting
 * status bits (clips, solver failure) of all n_steps steps are OR-ed into the final status word.
 *   order_dev  optional int32 [P]: permutation slot -> plant.  Results do not depend on it (plants
 *              are independent); it only decides which plants share a warp and which start first,
 *              e.g. plants sorted by the cost of their previous step (stragglers first, similar
 *              solver paths together).  NULL = identity.
 *   cost_dev   optional int32 [P]: work of this launch per plant (collocation solves + Newton
 *              iterations), the sort key for the next launch. */
int wt_advance(int P, int n_zones, int n_steps, double dt, const double *par_dev,
               const double *bnd_dev, int bnd_stride, double *time_dev, double *y_dev,
               double *flow_rate_dev, double *derived_dev, uint32_t *status_dev,
               int32_t *counters_dev, int max_attempts, const int32_t *order_dev,
               int32_t *cost_dev, void *workspace_dev, void *stream);

/* Deferral of budget-exhausted plants (the reference never drops a plant for work: solve_ivp runs to the end and a
 * failure only logs, reactor.py:476-490).
 *   wt_defer_collect  plants with WT_ST_WORK_LIMIT are appended to list_dev (capacity cap; count_dev must be 0 or hold
 *                     the number of entries already there) and re-marked WT_ST_DEFERRED: wt_step / wt_advance pass
 *                     over them, nothing else touches their state.  t_stop_dev (optional): *t_stop_dev += t_stop_inc first
 *                     (the end time of the block of steps that begins; no separate launch inside a captured graph).
 *   wt_catch_up       n_steps x step(dt) for the listed plants only, with their own (larger) budget, each plant only
 *                     until its time reaches *t_stop_dev; meant for a side stream while the ensemble moves on.  ONE
 *                     kernel launch for all n_steps (a warp stays with its plants).  Arrays are the ensemble's (row
 *                     stride ld = its plant count); workspace_dev is not used any more (may be NULL).
 *                     A plant that exhausts this budget too gets WT_ST_WORK_LIMIT and stays halted.
 *                     floor_div = 0: the reference's adaptive step control, as in wt_step.  floor_div > 0 (FLOOR MODE,
 *                     engine policy): the step size of every attempt is kept >= dt / floor_div, and at that floor an
 *                     error estimate above 1 no longer rejects and a Newton iteration that stops unconverged with a
 *                     current Jacobian keeps its last iterate; such plant-steps report WT_ST_DEGRADED.  This is how a
 *                     plant sitting on the 8 C density discontinuity (spatial.py:177-189) is continued at bounded cost
 *                     instead of being halted: <= ~4 floor_div collocation solves per step.
 *   wt_defer_rejoin   after the catch-up has finished (stream order): takes WT_ST_DEFERRED off the listed plants and
 *                     empties the list. */
int wt_defer_collect(int P, uint32_t *status_dev, int32_t *list_dev, int32_t *count_dev, int cap, double *t_stop_dev,
                     double t_stop_inc, void *stream);
int wt_catch_up(int cap, int ld, int n_zones, int n_steps, double dt, const double *par_dev, const double *bnd_dev,
                int bnd_stride, double *time_dev, double *y_dev, double *flow_rate_dev, double *derived_dev,
                uint32_t *status_dev, int32_t *counters_dev, int max_attempts, int floor_div, const int32_t *list_dev,
                const int32_t *count_dev, const double *t_stop_dev, void *workspace_dev, void *stream);
int wt_defer_rejoin(uint32_t *status_dev, const int32_t *list_dev, int32_t *count_dev, int cap, void *stream);

/* IntegratedCSTR.derivatives(t, y, boundary) for P plants (reactor.py:272-448).
 * dy has the layout of y; bad[p] != 0 where the reference would raise ValueError. */
int wt_derivatives(int P, int n_zones, const double *par_dev, const double *bnd_dev,
                   int bnd_stride, const double *y_dev, double *dy_dev, int32_t *bad_dev,
                   void *stream);

/* Host-buffer entry point (the call a host-only caller makes; used for end-to-end timing): copies
 * state / boundary (and the constants) in, runs wt_step, copies state / time / flow / status back,
 * and waits for completion.  All pointers are HOST pointers with the SoA layouts above;
 * bnd_stride is P or 0.  flags: WT_HOST_PARAMS_RESIDENT = the per-plant constants `par` are
 * unchanged since the previous call with the same (P, n_zones) and are not uploaded again.
 * Internally the call is pipelined over column slabs of the arrays: a copy-in stream, three compute streams and a
 * copy-out stream chained per slab by events (the upload of one slab overlaps the kernels of another and the download
 * of a third), with slab widths that ramp up and down (wt_step_host_plan); pinned host buffers are needed for the
 * overlap. */
#define WT_HOST_PARAMS_RESIDENT 1
/* The slab widths (plants) wt_step_host uses for P plants, in order: returns their number (<= 64) and writes up to `cap`
 * of them to sizes (may be NULL).  Host-only, no device needed.  Tuning: WT_B200_HOST_SLAB_MIN / _MAX / WT_B200_HOST_SLABS. */
int wt_step_host_plan(int P, int *sizes, int cap);
int wt_step_host(int P, int n_zones, double dt, const double *par, const double *bnd,
                 int bnd_stride, double *time, double *y, double *flow_rate, uint32_t *status,
                 int max_attempts, int flags);

/* ---------------------------------------------------------------------------------------
 * AqueousChemistry.calculate_pH(initial_guess) for P buffer systems   (chemistry.py:271-330)
 *   status: 0 converged, 1 |df/dpH| < 1e-15 (RuntimeError), 2 no convergence in 100 iterations
 *   (RuntimeError), 3 temperature outside [0,100] C (ValueError from the constructor)
 * ------------------------------------------------------------------------------------- */
int wt_calc_ph(int P, const double *alk_dev, const double *ct_dev, const double *temp_dev,
               const double *guess_dev, double *ph_dev, int32_t *iters_dev, int32_t *status_dev,
               void *stream);

/* ---------------------------------------------------------------------------------------
 * BaseSensor.get_statistics(window_seconds) for one sensor of every plant (base_sensor.py:757-856;
 * SURVEY.md section 8f rank 2) over a history of reading values kept on the device:
 *   hist_dev[(row * 7 + sensor) * P + p]  value of `sensor` of plant p at history row `row` (a ring the caller
 *                                         fills after each wt_sensors_read with the value rows of `out`)
 *   rows_dev[m]                            the ring rows inside the window (get_recent_readings: timestamps are the
 *                                          same for all plants, so the host selects them)
 *   out_dev[k * P + p], k = 0 mean, 1 std, 2 min, 3 max (over the finite values; NaN if there is none),
 *                       4 count (= m), 5 drift_rate (0.0: calculate_drift_rate takes the window newest-first, so
 *                       its "dt > 0" never holds, base_sensor.py:777-807), 6 fault_rate (non-finite / m).
 *   m == 0 (no readings yet) -> all zeros, as the reference.
 * ------------------------------------------------------------------------------------- */
int wt_sensor_window_stats(int P, int m, const double *hist_dev, const int32_t *rows_dev, int sensor,
                           double *out_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Wire-format tap: the Modbus input-register image of selected plants, built
 * on the device from the last read of the sensor suite, so that the untouched host Modbus layer can serve
 * any ensemble member (ModbusSlave.ir_block.setValues(0, row), di_block.setValues(0, bits)).  Replaces
 * update_modbus_inputs (__main__.py:166-224) + ModbusEncoder.float32_to_registers (modbus/protocols.py:34-58,
 * big-endian IEEE-754 single as (high word, low word)) + the |value| <= 1e9 check of
 * ModbusSlave.update_input_register (modbus/slave.py:139-164), with the addresses of ModbusRegisterMap
 * (modbus/register_map.py:119-244, 364-401):
 *   ir[k * WT_WIRE_NIR + a], a = register address 0..103: 0 pH_inlet, 4 pH_outlet, 6/8 chlorine in/out,
 *     10 flow_rate, 12/14 temperature in/out (two words each), 100 simulation_time, 102 system_status;
 *     pH_middle (2..3) is never written by the reference and stays 0.  NaN / inf readings are sent as 0.0.
 *   di[k * WT_WIRE_NDI + b]: b = 0 pH_inlet fault, 1 pH_outlet fault, 2 chlorine (inlet or outlet) fault
 *   ok[k] = 0 where the reference's update raises (a value outside +-1e9): that row is left all zero.
 *   sel_dev[K]: plant indices;  value_dev[s * P + p], fault_dev[s * P + p]: value row and fault codes of the
 *   last wt_sensors_read (s = sensor 0..6).
 * ------------------------------------------------------------------------------------- */
#define WT_WIRE_NIR 104
#define WT_WIRE_NDI 3
int wt_register_image(int K, const int32_t *sel_dev, int P, const double *value_dev, const int32_t *fault_dev,
                      double sim_time, uint16_t *ir_dev, uint8_t *di_dev, uint8_t *ok_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Per-plant diagnostics as one pass over the resident state.  Replaces,
 * for every plant of the ensemble at once:
 *   IntegratedCSTR.validate_conservation              reactor.py:570-611
 *   TransportModel.calculate_mixing_quality(chlorine) transport.py:338-384
 *   SpatialModel.calculate_spatial_gradients(x)       spatial.py:440-477    x = pH, chlorine, temperature
 *   SpatialModel.identify_thermocline                 spatial.py:353-379    (None -> NaN)
 *   SpatialModel.calculate_brunt_vaisala_frequency(i) spatial.py:322-351    on the density of the current T
 *   out[k * P + p], k = WT_DG_*;  n2_dev[i * P + p] (optional, may be NULL) = N^2 of interface i, i < n-1;
 *   h_dev (optional) = state.H_concentration rows (derived[0]); NULL -> 10^-pH;
 *   bad_dev[p] (optional) != 0 where the reference raises ValueError (T[0] outside [0, 100] C).
 * ------------------------------------------------------------------------------------- */
enum {
  WT_DG_TOTAL_CL_MG = 0, WT_DG_TOTAL_H_MOL, WT_DG_TOTAL_OH_MOL, WT_DG_CHARGE_BALANCE_MOL, WT_DG_THERMAL_ENERGY_KJ,
  WT_DG_CL_CV, WT_DG_CL_SEGREGATION,
  WT_DG_GRAD0,                       /* 3 variables x 8: mean, std, max, min, range, max_gradient, mean_gradient, gradient_location */
  WT_DG_THERMOCLINE_DEPTH = WT_DG_GRAD0 + 24, WT_DG_N2_MAX, WT_DG_N2_MIN, WT_NDIAG
};
int wt_diagnostics(int P, int n_zones, const double *par_dev, const double *y_dev, const double *h_dev,
                   double *out_dev, double *n2_dev, int32_t *bad_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Ensemble statistics: the payload of the ONE collective of the multi-GPU path.  The reference
 * has no counterpart (one plant, logging only: __main__.py:426-448, base_sensor.py:809-856);
 * BASELINE.json north_star asks for mean / variance / exceedance counts all-reduced over NCCL.
 *   shift_thr[7] = {shift_pH, shift_Cl, shift_T,  Cl_min, pH_lo, pH_hi, T_max}   (device)
 *   out[wt_stats_size(n)]:
 *     [0] live plants  [1] halted plants  [2] #outlet Cl < Cl_min  [3] #outlet pH outside [pH_lo,pH_hi]
 *     [4] #outlet T > T_max  [5] live plants whose last step was a floor-mode continuation (WT_ST_DEGRADED)
 *     [6] of the halted: plants over budget that are being caught up or wait for it (WT_ST_DEFERRED | WT_ST_WORK_LIMIT)
 *     [7] reserved
 *     [8 + 2*(var*n+zone)] = sum(x - shift[var]),  [9 + 2*(var*n+zone)] = sum (x - shift[var])^2   over live plants
 *   scratch: wt_stats_scratch_doubles(n) doubles of device workspace.  Deterministic (fixed
 *   summation order).  accumulate != 0 adds to `out` instead of overwriting it.
 * ------------------------------------------------------------------------------------- */
int wt_stats_size(int n_zones);
/* out[k] = sum over r < rows of in[r * n + k], in row order (bitwise reproducible): the statistics vectors of a rank's
 * sub-ensembles -> the rank's vector, before the all-reduce. */
int wt_sum_rows(int rows, int n, const double *in_dev, double *out_dev, void *stream);
int wt_stats_scratch_doubles(int n_zones);
int wt_stats(int P, int n_zones, const double *y_dev, const uint32_t *status_dev,
             const double *shift_thr_dev, double *out_dev, double *scratch_dev, int accumulate,
             void *stream);

/* ---------------------------------------------------------------------------------------
 * Sensor suite of create_realistic_sensor_suite (sensors/__init__.py:41-120): 7 sensors per plant in
 * the reference's dict order 0 pH_inlet, 1 pH_outlet, 2 chlorine_inlet (amperometric),
 * 3 chlorine_outlet (DPD), 4 flow_main (magnetic), 5 temp_inlet (RTD), 6 temp_outlet (RTD);
 * pH_x and temp_x share one SampleLine, as in the reference.
 *
 * Device buffers (plant index fastest; shapes in elements):
 *   sens     double [10][7][P]     per-sensor state (WT_SF_*: current_value, supply_voltage,
 *                                  calibration_offset, calibration time, last history value, 3 aux, power-on time,
 *                                  pH slope_percentage)
 *   sens_i   int32  [4][7][P]      sticky SensorStatus, SensorFault (enum order of base_sensor.py:49-75),
 *                                  len(reading_history), flags (bit 0: calibration_history is empty)
 *   ring     double [2][100][2][P] delay lines: (timestamp, value) per slot
 *   ring_i   int32  [2][2][P]      per line: head, count
 *   out      double [5][7][P]      SensorReading.value, raw_value, noise, drift, uncertainty
 *   out_status, out_fault  int32 [7][P]
 *
 * wt_sensors_init     = constructors + __main__.initialize_sensors (calibrate(reference, t0) with
 *                       reference 7.0 / initial_chlorine / temperature / flow_rate; __main__.py:96-105)
 * wt_sensors_calibrate = BaseSensor.calibrate(reference, t) for one sensor of every plant
 *                       (base_sensor.py:701-755); ref_dev NULL -> ref_scalar for all plants
 * wt_sensors_read     = <Sensor>.read(reactor_state, current_time) for the whole suite
 *                       (base_sensor.py:509-699 + ph_sensor.py:216-336, chlorine_sensor.py:345-484,
 *                       temperature_sensor.py:110-171, flow_sensor.py:125-219).  t must not decrease
 *                       (base_sensor.py:543-549 raises; the Python facade checks it).
 *     plant0      global id of the shard's first plant; read_index = number of earlier suite reads
 *                 (both only feed the counter-based Philox4x32-10 RNG: results are independent of sharding)
 *     suite8      HOST array {flow_velocity, air_bubble_frequency, grounding_quality, pipe_vibration_g,
 *                 ambient_temperature, sample-line transport delay [s], TemperatureSensorType of temp_inlet /
 *                 temp_outlet (0 rtd_pt100 = the factory's, 1 rtd_pt1000, 2 thermocouple_k, 3 thermocouple_j;
 *                 temperature_sensor.py:29-35, 150-194), FlowSensorType of flow_main (0 magnetic = the factory's,
 *                 1 turbine; flow_sensor.py:33-37, 180-219)}
 *     clock_dev   NULL, or a device array {t, t_prev, read_index, t0, dt} that overrides the t / t_prev / read_index
 *                 arguments: a captured (CUDA graph) step replays with frozen launch arguments and advances the
 *                 clock with wt_clock_tick after every read (t = t0 + read_index * dt)
 * wt_sensors_reset    = BaseSensor.reset() (base_sensor.py:858-878) for one sensor of every plant, with the
 *                       simulated time t where the reference stamps time.monotonic()
 * ------------------------------------------------------------------------------------- */
int wt_sensors_init(int P, double t0, const double *cfg_flow_dev, const double *cfg_chlorine_dev,
                    const double *cfg_temperature_dev, double *sens_dev, int32_t *sens_i_dev,
                    int32_t *ring_i_dev, void *stream);
int wt_sensors_calibrate(int P, int sensor, double t, const double *ref_dev, double ref_scalar,
                         double *sens_dev, int32_t *sens_i_dev, void *stream);

/* Maintenance operations of sensor `sensor` of every plant:
 *   op WT_MAINT_CAL2      pHSensor.calibrate_two_point(b1, b2, m1, m2, t)   ph_sensor.py:338-393     a0 = b1, a1 = b2
 *   op WT_MAINT_CLEAN     pHSensor.clean_electrode(method, t)               ph_sensor.py:395-434     a0 = 0 water_rinse, 1 acid_clean, 2 pepsin_clean
 *   op WT_MAINT_MEMBRANE  ChlorineSensor.replace_membrane(t)                chlorine_sensor.py:486-509
 *   op WT_MAINT_REAGENT   ChlorineSensor.replace_reagent(t)                 chlorine_sensor.py:511-537
 * Returns WT_ERR_BAD_ARG where the reference raises ValueError (wrong sensor kind, unknown cleaning method).
 * The measured buffer values m1, m2 only enter slope_percentage, which every read() overwrites
 * (ph_sensor.py:256-262): they do not change any later reading and are not part of the ABI. */
enum { WT_MAINT_CAL2 = 0, WT_MAINT_CLEAN = 1, WT_MAINT_MEMBRANE = 2, WT_MAINT_REAGENT = 3 };
int wt_sensors_maintain(int P, int sensor, int op, double t, double a0, double a1, double *sens_dev,
                        int32_t *sens_i_dev, void *stream);
int wt_sensors_read(int P, int n_zones, long long plant0, unsigned read_index, double t, double t_prev,
                    const double *y_dev, const double *flow_rate_dev, const double *cfg_flow_dev,
                    const double *cfg_chlorine_dev, const double *cfg_temperature_dev, double *sens_dev,
                    int32_t *sens_i_dev, double *ring_dev, int32_t *ring_i_dev, double *out_dev,
                    int32_t *out_status_dev, int32_t *out_fault_dev, const double *suite8,
                    uint64_t seed, const double *clock_dev, void *stream);
int wt_sensors_reset(int P, int sensor, double t, const double *cfg_flow_dev, double *sens_dev,
                     int32_t *sens_i_dev, int32_t *ring_i_dev, void *stream);
int wt_clock_tick(double *clock_dev, void *stream);

/* Per-sensor ensemble statistics of the last suite read: the sensor half of the all-reduce payload of SURVEY.md
 * section 8(e) / BASELINE configs[4] (reference analogue: BaseSensor.get_statistics, base_sensor.py:809-856,
 * taken across plants instead of across time).  Halted plants are skipped.
 *   stats[s * 22 + 0] readings with a finite value, [1] sum (value - shift7[s]), [2] sum (value - shift7[s])^2,
 *   [3 .. 14] SensorStatus histogram, [15 .. 21] SensorFault histogram        (s = sensor 0..6)
 *   scratch: wt_sensor_stats_scratch_doubles() doubles.  Deterministic (fixed summation order). */
int wt_sensor_stats_size(void);
int wt_sensor_stats_scratch_doubles(void);
int wt_sensor_stats(int P, const double *out_value_dev, const int32_t *out_status_dev,
                    const int32_t *out_fault_dev, const uint32_t *plant_status_dev, const double *shift7_dev,
                    double *stats_dev, double *scratch_dev, int accumulate, void *stream);

/* Scheduling helper: order_dev <- plants sorted by cost_dev (the per-plant work wt_advance reports), most expensive
 * first (counting sort; bins_dev: 1024 int32 of device scratch).  Results of wt_advance never depend on the order. */
int wt_cost_order(int P, const int32_t *cost_dev, int32_t *order_dev, int32_t *bins_dev, void *stream);

/* ---------------------------------------------------------------------------------------
 * Orchestrator (the caller of the path): actuator commands -> boundary conditions with the
 * reference's two layers of zero-trust clamps, in place on the boundary SoA [WT_NBND][P]:
 *   validate_flow_rate (__main__.py:57-63): NaN -> 0, clamp to [0, max]; read_modbus_commands (:227-252) clamps acid to
 *   2, chlorine to 1, inlet to 20 L/min; apply_boundary_conditions (:255-271) clamps again and updates the inlet flow
 *   only when the command exceeds 0.1 L/min.
 * wt_apply_commands     per-plant command arrays (a controller's output, device memory).
 * wt_scenario_commands  scenario scripting without per-step host traffic: S scripts of K piecewise-constant command
 *                       triplets cmd[S][K][3] = (acid, chlorine, inlet) with ascending breakpoints times[K]; plant p
 *                       follows script sid[p] (NULL: script 0).  The time is clock_dev[0] (the device clock advanced by
 *                       wt_clock_tick; NULL: the host value t).  Before times[0] the boundary is left untouched.
 * ------------------------------------------------------------------------------------- */
int wt_apply_commands(int P, const double *acid_dev, const double *chlorine_dev, const double *inlet_dev,
                      double *bnd_soa_dev, void *stream);
int wt_scenario_commands(int P, int K, int S, const double *times_dev, const double *cmd_dev,
                         const int32_t *sid_dev, const double *clock_dev, double t, double *bnd_soa_dev,
                         void *stream);

/* Measured-peak helper for the roofline denominator: runs a dependent-chain-free DFMA loop on
 * every SM and returns the sustained FP64 rate in TFLOP/s (2 flops per DFMA). */
int wt_measure_fp64_peak(double *tflops_out, int iters);

#ifdef __cplusplus
}
#endif
#endif
