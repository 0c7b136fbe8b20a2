"""numpy restatement of the command -> boundary path of the reference's main loop.  TEST INFRASTRUCTURE ONLY
(tests/ only; the product path is the wt_apply_commands / wt_scenario_commands kernels).

Follows src/wt_simulator/__main__.py: validate_flow_rate (:57-63), read_modbus_commands (:239-246),
apply_boundary_conditions (:255-271).  Pinned against outputs of those functions themselves
(tests/golden/commands.npz, oracle/gen_golden_commands.py)."""
import numpy as np


def validate_flow_rate(value, max_value=20.0):
    """__main__.py:57-63: NaN -> 0.0, else max(0.0, min(value, max_value))."""
    v = np.asarray(value, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        return np.where(np.isnan(v), 0.0, np.maximum(0.0, np.minimum(v, max_value)))


def apply_commands(acid, chlorine, inlet, inlet_before):
    """read_modbus_commands' clamps (:239-246) then apply_boundary_conditions (:255-271).
    Returns (acid_flow_rate, chlorine_flow_rate, inlet_flow_rate) of the boundary afterwards."""
    a = validate_flow_rate(validate_flow_rate(acid, 2.0), 2.0)
    c = validate_flow_rate(validate_flow_rate(chlorine, 1.0), 1.0)
    i = validate_flow_rate(inlet, 20.0)
    return a, c, np.where(i > 0.1, validate_flow_rate(i, 20.0), np.asarray(inlet_before, dtype=np.float64))


def scenario_commands(times, cmd_sk3, sid, t):
    """Commands of the segment containing t for every plant (None before the first breakpoint)."""
    times = np.asarray(times, dtype=np.float64)
    k = int(np.searchsorted(times, t, side="right")) - 1
    if k < 0:
        return None
    sid = np.clip(np.asarray(sid), 0, cmd_sk3.shape[0] - 1)
    return cmd_sk3[sid, k, :]
