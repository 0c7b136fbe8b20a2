#!/usr/bin/env python
"""Golden vectors of the main loop's command -> boundary path, produced by RUNNING THE REFERENCE'S OWN FUNCTIONS.

wt_simulator/__main__.py imports pymodbus (absent here) at module level, so the three functions on this path --
validate_flow_rate (:57-63), read_modbus_commands (:227-252) and apply_boundary_conditions (:255-271) -- are taken out
of the file with ``ast`` (their source is compiled unchanged, never copied into the repo) and executed against the real
``BoundaryConditions`` of wt_simulator.core.reactor and a stand-in Modbus slave that only serves the three holding
registers.  Build container only.      python oracle/gen_golden_commands.py  ->  tests/golden/commands.npz
"""
import ast
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

WANT = ("validate_flow_rate", "read_modbus_commands", "apply_boundary_conditions")


def reference_functions():
    from typing import Optional, Tuple
    from wt_simulator.core.reactor import BoundaryConditions
    path = os.path.join(REF, "wt_simulator", "__main__.py")
    tree = ast.parse(open(path).read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    assert sorted(n.name for n in body) == sorted(WANT)
    ns = {"Optional": Optional, "Tuple": Tuple, "BoundaryConditions": BoundaryConditions, "ModbusSlave": object,
          "logger": logging.getLogger("golden")}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns, BoundaryConditions


class Slave:
    """What read_modbus_commands needs of ModbusSlave: is_running and the three holding registers."""
    is_running = True

    def __init__(self, acid, chlorine, inlet):
        self.regs = {"acid_flow_rate": acid, "chlorine_flow_rate": chlorine, "inlet_flow_rate": inlet}

    def read_holding_register(self, name):
        return self.regs[name]


def main():
    ns, BoundaryConditions = reference_functions()
    rng = np.random.default_rng(20260008)
    N = 4096
    special = [float("nan"), float("inf"), -float("inf"), -1.0, -0.0, 0.0, 0.05, 0.1, 0.1000001, 0.5, 1.0, 1.5, 2.0, 2.5, 19.9, 20.0,
               20.0000001, 1e9, -1e9]
    cmd = np.stack([rng.uniform(-1, 4, N), rng.uniform(-1, 2, N), rng.uniform(-2, 30, N)], axis=1)
    pick = rng.random((N, 3)) < 0.25
    cmd[pick] = rng.choice(special, size=int(pick.sum()))
    inlet_before = rng.uniform(0.0, 20.0, N)
    out = np.zeros((N, 3))
    for i in range(N):
        b = BoundaryConditions(inlet_flow_rate=float(inlet_before[i]))
        commands = ns["read_modbus_commands"](Slave(float(cmd[i, 0]), float(cmd[i, 1]), float(cmd[i, 2])))
        ns["apply_boundary_conditions"](b, commands)
        out[i] = (b.acid_flow_rate, b.chlorine_flow_rate, b.inlet_flow_rate)
    vfr_in = np.array(special + list(rng.uniform(-5, 30, 200)))
    vfr = np.array([[ns["validate_flow_rate"](float(v), mx) for v in vfr_in] for mx in (1.0, 2.0, 20.0)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "commands.npz"), commands=cmd, inlet_before=inlet_before,
                        boundary_after=out, vfr_in=vfr_in, vfr_max=np.array([1.0, 2.0, 20.0]), vfr_out=vfr,
                        source="wt_simulator/__main__.py:57-63, 227-252, 255-271 (functions extracted with ast and executed)")
    print("wrote", N, "cases;", int(np.isnan(cmd).any(axis=1).sum()), "with NaN commands")


if __name__ == "__main__":
    main()
