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
ST_HALT_MASK = ST_T_RANGE | ST_WORK_LIMIT


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "wt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
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
        L.wt_oracle_set_ph_h_eps.argtypes = [C.c_double]
        L.wt_oracle_set_ph_h_eps.restype = None
        L.wt_oracle_calc_ph.argtypes = [C.c_double] * 5 + [C.c_int, dp, ip]
        L.wt_oracle_calc_ph.restype = C.c_int
        L.wt_oracle_calc_ph_batch.argtypes = [C.c_int, dp, dp, dp, dp, dp, ip, ip, C.c_int]
        L.wt_oracle_calc_ph_batch.restype = None
        _lib = L
    return _lib


def set_max_attempts(m: int) -> None:
    lib().wt_oracle_set_max_attempts(int(m))


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
