"""CPU oracle of the Modbus input-register image -- TEST INFRASTRUCTURE.

Restates what the reference's main loop puts on the wire for ONE plant per call:
  update_modbus_inputs                 __main__.py:166-224   which readings go to which register, safe_value
                                                              (NaN / inf -> 0.0), system_status, the three fault bits
  ModbusEncoder.float32_to_registers   modbus/protocols.py:34-58   big-endian IEEE-754 single -> (high, low) words
  ModbusSlave.update_input_register    modbus/slave.py:139-164     the |value| <= 1e9 check
  ModbusRegisterMap                    modbus/register_map.py:119-244, 364-401   register addresses
Pinned against the reference's own encoder and register map by tests/golden/wire_image.npz
(oracle/gen_golden_wire.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import struct

import numpy as np

N_IR = 104      # input-register image: addresses 0..103
N_DI = 3
# sensor order of the batched suite -> input-register address (register_map.py:119-215, __main__.py:194-205)
SENSORS = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")
IR_ADDR = {"pH_inlet": 0, "pH_outlet": 4, "chlorine_inlet": 6, "chlorine_outlet": 8, "flow_main": 10,
           "temp_inlet": 12, "temp_outlet": 14}
IR_TIME, IR_STATUS = 100, 102


def float32_to_registers(value: float):
    high, low = struct.unpack(">HH", struct.pack(">f", value))
    return high, low


def register_image(values: np.ndarray, faults: np.ndarray, sim_time: float):
    """values [K, 7] sensor readings (may hold NaN / inf), faults [K, 7] SensorFault codes (0 = NONE).
    Returns (ir [K, N_IR] uint16, di [K, N_DI] uint8, ok [K] bool) -- ok is False where the reference's
    update raises (|value| > 1e9) and leaves the remaining registers untouched; such rows are all zero here."""
    K = values.shape[0]
    ir = np.zeros((K, N_IR), np.uint16)
    di = np.zeros((K, N_DI), np.uint8)
    ok = np.ones(K, bool)
    for k in range(K):
        safe = [0.0 if (v != v or v in (float("inf"), float("-inf"))) else float(v) for v in values[k]]
        if any(not (-1e9 <= v <= 1e9) for v in safe) or not (-1e9 <= sim_time <= 1e9):
            ok[k] = False
            continue
        for s, name in enumerate(SENSORS):
            ir[k, IR_ADDR[name]], ir[k, IR_ADDR[name] + 1] = float32_to_registers(safe[s])
        ir[k, IR_TIME], ir[k, IR_TIME + 1] = float32_to_registers(sim_time)
        ir[k, IR_STATUS] = 1 if np.any(faults[k] != 0) else 0
        di[k, 0] = faults[k, 0] != 0
        di[k, 1] = faults[k, 1] != 0
        di[k, 2] = (faults[k, 2] != 0) or (faults[k, 3] != 0)
    return ir, di, ok
