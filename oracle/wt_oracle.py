"""ctypes binding of the CPU oracle (oracle/wt_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module.  The product package
(ics_wt_physicsengine_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwt_oracle.so")

# index constants, mirrored from wt_oracle.h
CFG_FIELDS = (
    "volume", "height", "diameter", "flow_rate", "turbulent_intensity", "recirculation_ratio",
    "impeller_speed", "impeller_diameter", "power_number", "initial_pH", "alkalinity",
    "total_carbonate", "initial_chlorine", "temperature", "enable_thermal_stratification",
)
NCFG = len(CFG_FIELDS)
NPAR = 12
BND_FIELDS = (
    "inlet_flow_rate", "inlet_pH", "inlet_chlorine", "inlet_temperature", "acid_flow_rate",
    "acid_concentration", "chlorine_flow_rate", "chlorine_concentration", "ambient_temperature",
    "heat_loss_coefficient",
)
NBND = len(BND_FIELDS)
NCNT = 8
CNT_NFEV, CNT_NJEV, CNT_NLU, CNT_NSTEPS, CNT_NNEWTON, CNT_NREJECT, CNT_NNEWTON_FAIL = range(7)

ST_SOLVER_FAILED = 1
ST_T_RANGE = 2
ST_CLIP_PH = 4
ST_CLIP_CL = 8
ST_CLIP_T = 16
ST_NONFINITE = 32
ST_T_RANGE_DERIVED = 64
ST_WORK_LIMIT = 128
ST_DEFERRED = 256
ST_DEGRADED = 512
ST_HALT_MASK = ST_T_RANGE | ST_WORK_LIMIT


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in ("wt_oracle.c", "wt_sensors_oracle.c", "wt_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libwt_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        up = C.POINTER(C.c_uint32)
        L.wt_oracle_derive_params.argtypes = [dp, C.c_int, dp]
        L.wt_oracle_derive_params.restype = C.c_int
        L.wt_oracle_rhs.argtypes = [dp, dp, C.c_int, dp, dp]
        L.wt_oracle_rhs.restype = C.c_int
        L.wt_oracle_step.argtypes = [dp, dp, C.c_int, C.c_double, dp, dp, dp, dp, ip]
        L.wt_oracle_step.restype = C.c_uint32
        L.wt_oracle_step_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, C.c_int,
                                           dp, dp, dp, up, ip, C.c_int]
        L.wt_oracle_step_batch.restype = None
        L.wt_oracle_set_max_attempts.argtypes = [C.c_int]
        L.wt_oracle_set_max_attempts.restype = None
        L.wt_oracle_set_floor_div.argtypes = [C.c_int]
        L.wt_oracle_set_floor_div.restype = None
        L.wt_oracle_set_ph_h_eps.argtypes = [C.c_double]
        L.wt_oracle_set_ph_h_eps.restype = None
        L.wt_oracle_calc_ph.argtypes = [C.c_double] * 5 + [C.c_int, dp, ip]
        L.wt_oracle_calc_ph.restype = C.c_int
        L.wt_oracle_calc_ph_batch.argtypes = [C.c_int, dp, dp, dp, dp, dp, ip, ip, C.c_int]
        L.wt_oracle_calc_ph_batch.restype = None
        L.wt_oracle_suite_bytes.restype = C.c_int
        L.wt_oracle_sensors_init.argtypes = [C.c_int, C.c_double, dp, dp, dp, C.c_void_p, C.c_int, C.c_int]
        L.wt_oracle_sensors_init.restype = None
        L.wt_oracle_sensors_reset.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.wt_oracle_sensors_reset.restype = None
        L.wt_oracle_sensors_calibrate.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.c_void_p]
        L.wt_oracle_sensors_calibrate.restype = None
        L.wt_oracle_sensors_read.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_uint, C.c_double, C.c_double, dp, dp,
                                             C.c_void_p, dp, ip, ip, dp, C.c_uint64, C.c_int]
        L.wt_oracle_sensors_read.restype = None
        L.wt_oracle_sensors_maintain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, dp, C.c_void_p]
        L.wt_oracle_sensors_maintain.restype = C.c_int
        L.wt_oracle_sensors_poke.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_void_p]
        L.wt_oracle_sensors_poke.restype = None
        L.wt_oracle_sensors_peek.argtypes = [C.c_int, C.c_int, dp, C.c_void_p]
        L.wt_oracle_sensors_peek.restype = None
        _lib = L
    return _lib


def set_max_attempts(m: int) -> None:
    lib().wt_oracle_set_max_attempts(int(m))


def set_floor_div(d: int) -> None:
    """Engine policy mirror (floor mode): step sizes >= dt / d with forced acceptance; 0 = off."""
    lib().wt_oracle_set_floor_div(int(d))


def set_ph_h_eps(eps: float) -> None:
    lib().wt_oracle_set_ph_h_eps(float(eps))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def derive_params(cfg: np.ndarray, n_zones: int) -> np.ndarray:
    """cfg: [P, NCFG] float64 -> par: [P, NPAR] float64."""
    cfg = np.ascontiguousarray(cfg, dtype=np.float64).reshape(-1, NCFG)
    par = np.empty((cfg.shape[0], NPAR), dtype=np.float64)
    for p in range(cfg.shape[0]):
        rc = lib().wt_oracle_derive_params(_dp(cfg[p]), n_zones, _dp(par[p]))
        if rc:
            raise ValueError(f"plant {p}: configuration rejected (rc={rc})")
    return par


def rhs(par: np.ndarray, bnd: np.ndarray, n: int, y: np.ndarray):
    par = np.ascontiguousarray(par, dtype=np.float64)
    bnd = np.ascontiguousarray(bnd, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    dy = np.empty(3 * n)
    rc = lib().wt_oracle_rhs(_dp(par), _dp(bnd), n, _dp(y), _dp(dy))
    return dy, rc


def step_batch(par, bnd, n, t, y, dt=1.0, nsteps=1, nthreads=1):
    """Advance P plants in place.

    par [P,NPAR]; bnd [P,NBND] or [NBND] (broadcast); t [P]; y [P,3n] species-major.
    Returns (status[P] uint32, counters[P,NCNT] int32, flow_rate[P]).
    """
    P = y.shape[0]
    assert par.flags.c_contiguous and y.flags.c_contiguous and t.flags.c_contiguous
    assert par.dtype == np.float64 and y.dtype == np.float64 and t.dtype == np.float64
    bnd = np.ascontiguousarray(bnd, dtype=np.float64)
    stride = 0 if bnd.ndim == 1 else NBND
    status = np.zeros(P, dtype=np.uint32)
    counters = np.zeros((P, NCNT), dtype=np.int32)
    flow = np.zeros(P)
    lib().wt_oracle_step_batch(P, n, nsteps, float(dt), _dp(par), _dp(bnd), stride, _dp(t), _dp(y),
                               _dp(flow), _up(status), _ip(counters), nthreads)
    return status, counters, flow


def calc_ph_batch(alk, ct, temp, guess, nthreads=1):
    alk = np.ascontiguousarray(alk, dtype=np.float64)
    ct = np.ascontiguousarray(ct, dtype=np.float64)
    temp = np.ascontiguousarray(temp, dtype=np.float64)
    guess = np.ascontiguousarray(guess, dtype=np.float64)
    P = alk.shape[0]
    ph = np.empty(P)
    iters = np.zeros(P, dtype=np.int32)
    status = np.zeros(P, dtype=np.int32)
    lib().wt_oracle_calc_ph_batch(P, _dp(alk), _dp(ct), _dp(temp), _dp(guess), _dp(ph), _ip(iters),
                                  _ip(status), nthreads)
    return ph, iters, status


SENSOR_NAMES = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")
# InstallationQuality of create_realistic_sensor_suite (sensors/__init__.py:53-59) + SampleLine delay (:62-67)
STANDARD_SUITE6 = np.array([0.5, 0.0, 0.9, 0.1, 30.0, (250 / 1000.0) / (500 / 1000.0 / 60.0)])


class SensorSuiteOracle:
    """P instances of the reference sensor suite (CPU port), calibrated at t0 as __main__.initialize_sensors does."""

    def __init__(self, cfg_flow, cfg_cl, cfg_T, t0, seed=0, plant0=0, suite6=None, nthreads=1, temp_kind=0, flow_kind=0):
        self.P = len(cfg_flow)
        self.t0, self.seed, self.plant0, self.nthreads = float(t0), int(seed), int(plant0), nthreads
        self.suite6 = np.ascontiguousarray(STANDARD_SUITE6 if suite6 is None else suite6, dtype=np.float64)
        self._buf = np.zeros(self.P * lib().wt_oracle_suite_bytes(), dtype=np.uint8)
        f, c, t = (np.ascontiguousarray(a, dtype=np.float64) for a in (cfg_flow, cfg_cl, cfg_T))
        lib().wt_oracle_sensors_init(self.P, self.t0, _dp(f), _dp(c), _dp(t), self._buf.ctypes.data, int(temp_kind),
                                     int(flow_kind))
        self.k = 0
        self.t_prev = self.t0

    def calibrate(self, sensor: int, reference, t: float):
        ref = np.ascontiguousarray(np.broadcast_to(np.asarray(reference, dtype=np.float64), (self.P,)))
        lib().wt_oracle_sensors_calibrate(self.P, sensor, float(t), _dp(ref), self._buf.ctypes.data)

    def reset(self, sensor: int, t: float):
        """BaseSensor.reset() with the simulated time t in place of time.monotonic()."""
        lib().wt_oracle_sensors_reset(self.P, sensor, float(t), self._buf.ctypes.data)

    FIELDS = ("current_value", "calibration_offset", "last_calibration_time", "power_on_time", "membrane_fouling",
              "reference_contamination", "days_since_cleaning", "membrane_age_days", "reagent_potency",
              "light_exposure_hours", "reagent_age_days", "status", "fault")

    def maintain(self, sensor: int, op: int, t: float, args=(0.0, 0.0, 0.0, 0.0)) -> int:
        """op 0 calibrate_two_point(b1, b2, m1, m2) | 1 clean_electrode(method code) | 2 replace_membrane | 3 replace_reagent."""
        a = np.ascontiguousarray(np.asarray(args, dtype=np.float64).reshape(4))
        return lib().wt_oracle_sensors_maintain(self.P, sensor, op, float(t), _dp(a), self._buf.ctypes.data)

    def poke(self, sensor: int, field: str, values):
        v = np.ascontiguousarray(np.broadcast_to(np.asarray(values, dtype=np.float64), (self.P,)))
        lib().wt_oracle_sensors_poke(self.P, sensor, self.FIELDS.index(field), _dp(v), self._buf.ctypes.data)

    def peek(self, sensor: int):
        out = np.zeros((self.P, 13))
        lib().wt_oracle_sensors_peek(self.P, sensor, _dp(out), self._buf.ctypes.data)
        return {k: out[:, i] for i, k in enumerate(self.FIELDS)}

    def read(self, y_pn, flow, t, n):
        """y_pn [P,3n] species-major, flow [P] -> (out [P,7,5] value/raw/noise/drift/uncertainty, status [P,7], fault [P,7])."""
        y = np.ascontiguousarray(y_pn, dtype=np.float64)
        fl = np.ascontiguousarray(flow, dtype=np.float64)
        out = np.zeros((self.P, 7, 5))
        st = np.zeros((self.P, 7), dtype=np.int32)
        ft = np.zeros((self.P, 7), dtype=np.int32)
        lib().wt_oracle_sensors_read(self.P, n, self.plant0, self.k, float(t), float(self.t_prev), _dp(y), _dp(fl),
                                     self._buf.ctypes.data, _dp(out), _ip(st), _ip(ft), _dp(self.suite6), self.seed,
                                     self.nthreads)
        self.k += 1
        self.t_prev = float(t)
        return out, st, ft
