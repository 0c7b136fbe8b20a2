"""ctypes binding of the C ABI in include/wt_b200.h (csrc/libwt_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no
CPU fallback: if the shared library is missing, or no CUDA device is present when a compute
entry point is called, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# WT_B200_LIB selects an alternative build of the SAME library (A/B tuning builds); never a CPU path
LIB_PATH = os.environ.get("WT_B200_LIB") or os.path.join(_HERE, "csrc", "libwt_b200.so")

ABI_VERSION = 4   # WT_ABI_VERSION of include/wt_b200.h
NPAR = 12
NBND = 10
NCNT = 8

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
ST_SKIP_MASK = ST_HALT_MASK | ST_DEFERRED

CNT_NAMES = ("nfev", "njev", "nlu", "nsteps", "nnewton", "nreject", "nnewton_fail", "jac_retry")

EXPORTS = (
    "wt_abi_version", "wt_device_count", "wt_last_error", "wt_step", "wt_advance", "wt_derivatives",
    "wt_step_host", "wt_step_workspace_bytes", "wt_calc_ph", "wt_measure_fp64_peak", "wt_stats", "wt_stats_size", "wt_stats_scratch_doubles",
    "wt_sensors_init", "wt_sensors_calibrate", "wt_sensors_read", "wt_diagnostics", "wt_register_image",
    "wt_sensors_maintain", "wt_sensor_window_stats", "wt_sensors_reset", "wt_clock_tick", "wt_sensor_stats_size",
    "wt_sensor_stats_scratch_doubles", "wt_sensor_stats", "wt_cost_order", "wt_apply_commands", "wt_scenario_commands",
    "wt_defer_collect", "wt_catch_up", "wt_defer_rejoin", "wt_sum_rows", "wt_step_host_plan",
)


class EngineError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    L.wt_abi_version.restype = C.c_int
    if L.wt_abi_version() != ABI_VERSION:
        raise EngineError(f"{LIB_PATH} has ABI version {L.wt_abi_version()}, this package needs {ABI_VERSION}: rebuild "
                          "(python -c 'import __graft_entry__ as g; g.build()')")
    vp, dp, ip, up = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p  # raw device addresses
    L.wt_abi_version.restype = C.c_int
    L.wt_device_count.restype = C.c_int
    L.wt_last_error.restype = C.c_char_p
    L.wt_step.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, up, ip, C.c_int, vp, vp]
    L.wt_step.restype = C.c_int
    L.wt_step_workspace_bytes.argtypes = [C.c_int, C.c_int]
    L.wt_step_workspace_bytes.restype = C.c_size_t
    L.wt_advance.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, up, ip,
                             C.c_int, ip, ip, vp, vp]
    L.wt_advance.restype = C.c_int
    L.wt_derivatives.argtypes = [C.c_int, C.c_int, dp, dp, C.c_int, dp, dp, ip, vp]
    L.wt_derivatives.restype = C.c_int
    L.wt_step_host.argtypes = [C.c_int, C.c_int, C.c_double, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int]
    L.wt_step_host.restype = C.c_int
    L.wt_step_host_plan.argtypes = [C.c_int, ip, C.c_int]
    L.wt_step_host_plan.restype = C.c_int
    L.wt_calc_ph.argtypes = [C.c_int, dp, dp, dp, dp, dp, ip, ip, vp]
    L.wt_calc_ph.restype = C.c_int
    L.wt_stats_size.argtypes = [C.c_int]
    L.wt_stats_size.restype = C.c_int
    L.wt_stats_scratch_doubles.argtypes = [C.c_int]
    L.wt_stats_scratch_doubles.restype = C.c_int
    L.wt_stats.argtypes = [C.c_int, C.c_int, dp, up, dp, dp, dp, C.c_int, vp]
    L.wt_stats.restype = C.c_int
    L.wt_sensors_init.argtypes = [C.c_int, C.c_double, dp, dp, dp, dp, ip, ip, vp]
    L.wt_sensors_init.restype = C.c_int
    L.wt_sensors_calibrate.argtypes = [C.c_int, C.c_int, C.c_double, dp, C.c_double, dp, ip, vp]
    L.wt_sensors_calibrate.restype = C.c_int
    L.wt_sensors_read.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_uint, C.c_double, C.c_double, dp, dp, dp, dp, dp, dp, ip,
                                  dp, ip, dp, ip, ip, C.POINTER(C.c_double), C.c_uint64, dp, vp]
    L.wt_sensors_read.restype = C.c_int
    L.wt_sensors_reset.argtypes = [C.c_int, C.c_int, C.c_double, dp, dp, ip, ip, vp]
    L.wt_sensors_reset.restype = C.c_int
    L.wt_clock_tick.argtypes = [dp, vp]
    L.wt_clock_tick.restype = C.c_int
    L.wt_sensor_stats_size.restype = C.c_int
    L.wt_sensor_stats_scratch_doubles.restype = C.c_int
    L.wt_sensor_stats.argtypes = [C.c_int, dp, ip, ip, up, dp, dp, dp, C.c_int, vp]
    L.wt_sensor_stats.restype = C.c_int
    L.wt_cost_order.argtypes = [C.c_int, ip, ip, ip, vp]
    L.wt_cost_order.restype = C.c_int
    L.wt_defer_collect.argtypes = [C.c_int, up, ip, ip, C.c_int, dp, C.c_double, vp]
    L.wt_sum_rows.argtypes = [C.c_int, C.c_int, dp, dp, vp]
    L.wt_sum_rows.restype = C.c_int
    L.wt_defer_collect.restype = C.c_int
    L.wt_catch_up.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp, dp, C.c_int, dp, dp, dp, dp, up, ip, C.c_int,
                              C.c_int, ip, ip, dp, vp, vp]
    L.wt_catch_up.restype = C.c_int
    L.wt_defer_rejoin.argtypes = [up, ip, ip, C.c_int, vp]
    L.wt_defer_rejoin.restype = C.c_int
    L.wt_apply_commands.argtypes = [C.c_int, dp, dp, dp, dp, vp]
    L.wt_apply_commands.restype = C.c_int
    L.wt_scenario_commands.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, ip, dp, C.c_double, dp, vp]
    L.wt_scenario_commands.restype = C.c_int
    L.wt_diagnostics.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, ip, vp]
    L.wt_diagnostics.restype = C.c_int
    L.wt_register_image.argtypes = [C.c_int, ip, C.c_int, dp, ip, C.c_double, vp, vp, vp, vp]
    L.wt_register_image.restype = C.c_int
    L.wt_sensors_maintain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, dp, ip, vp]
    L.wt_sensors_maintain.restype = C.c_int
    L.wt_sensor_window_stats.argtypes = [C.c_int, C.c_int, dp, ip, C.c_int, dp, vp]
    L.wt_sensor_window_stats.restype = C.c_int
    L.wt_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.wt_measure_fp64_peak.restype = C.c_int
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().wt_last_error().decode(errors="replace")
        raise EngineError(f"{what} failed (rc={rc}): {msg}")


def require_device() -> None:
    if lib().wt_device_count() <= 0:
        raise EngineError("no CUDA device visible: the engine has no CPU fallback")


def measure_fp64_peak(iters: int = 200000) -> float:
    require_device()
    out = C.c_double(0.0)
    check(lib().wt_measure_fp64_peak(C.byref(out), iters), "wt_measure_fp64_peak")
    return out.value
