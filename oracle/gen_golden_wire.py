#!/usr/bin/env python
"""Golden vectors of the wire format, from the reference's own ModbusEncoder and ModbusRegisterMap
(modbus/protocols.py and modbus/register_map.py load standalone; the server itself needs pymodbus, which
is not installed and is out of scope).  Build-container only.

    python oracle/gen_golden_wire.py      # writes tests/golden/wire_image.npz
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, "wt_simulator", "modbus", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_" + name] = m
    spec.loader.exec_module(m)
    return m


def main():
    prot, rmap = _load("protocols"), _load("register_map")
    enc = prot.ModbusEncoder()
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.uniform(-20, 120, 400), rng.normal(0, 1e-3, 100), 10.0 ** rng.uniform(-40, 9, 200),
                           -(10.0 ** rng.uniform(-40, 9, 100)), [0.0, -0.0, 7.25, 1e9, -1e9, 3.4e38 * 0 + 16777217.0]])
    words = np.array([enc.float32_to_registers(float(v)) for v in vals], dtype=np.uint16)
    m = rmap.ModbusRegisterMap()
    addr = {r.name: (r.address, r.data_type) for r in m.input_registers}
    di = {r.name: r.address for r in m.discrete_inputs}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wire_image.npz"), values=vals, words=words,
                        ir_names=np.array(list(addr)), ir_addr=np.array([a for a, _ in addr.values()]),
                        ir_type=np.array([t for _, t in addr.values()]), di_names=np.array(list(di)),
                        di_addr=np.array(list(di.values())))
    print("wrote", len(vals), "encodings,", len(addr), "input registers,", len(di), "discrete inputs")


if __name__ == "__main__":
    main()
