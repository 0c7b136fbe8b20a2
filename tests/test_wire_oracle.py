"""The wire-format oracle (oracle/wt_wire_oracle.py) against the reference's own ModbusEncoder outputs and
register map (tests/golden/wire_image.npz, oracle/gen_golden_wire.py)."""
import os

import numpy as np

from oracle import wt_wire_oracle as ww


def test_encoder_and_addresses_match_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "wire_image.npz"))
    got = np.array([ww.float32_to_registers(float(v)) for v in g["values"]], dtype=np.uint16)
    assert np.array_equal(got, g["words"])
    addr = dict(zip(g["ir_names"], g["ir_addr"]))
    # __main__.py:194-205 writes these registers from these readings
    ref_name = {"pH_inlet": "pH_inlet", "pH_outlet": "pH_outlet", "chlorine_inlet": "chlorine_inlet",
                "chlorine_outlet": "chlorine_outlet", "flow_main": "flow_rate", "temp_inlet": "temperature_inlet",
                "temp_outlet": "temperature_outlet"}
    for s, a in ww.IR_ADDR.items():
        assert addr[ref_name[s]] == a
    assert addr["simulation_time"] == ww.IR_TIME and addr["system_status"] == ww.IR_STATUS
    assert dict(zip(g["ir_names"], g["ir_type"]))["system_status"] == "uint16"
    assert max(addr.values()) + 1 < ww.N_IR + 1
    assert list(g["di_names"]) == ["sensor_fault_pH_inlet", "sensor_fault_pH_outlet", "sensor_fault_chlorine"]
    assert list(g["di_addr"]) == [0, 1, 2]


def test_image_semantics():
    v = np.array([[7.25, np.nan, 1.5, np.inf, 5.0, 20.0, -np.inf], [7.0, 7.1, 2e9, 0.1, 5.0, 20.0, 21.0]])
    f = np.array([[0, 0, 0, 3, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0]])
    ir, di, ok = ww.register_image(v, f, 12.0)
    assert list(ok) == [True, False] and not ir[1].any()
    assert (ir[0, 0], ir[0, 1]) == ww.float32_to_registers(7.25)
    assert (ir[0, 4], ir[0, 5]) == (0, 0) and (ir[0, 8], ir[0, 9]) == (0, 0)   # NaN / inf -> 0.0
    assert (ir[0, 2], ir[0, 3]) == (0, 0)                                      # pH_middle is never written
    assert ir[0, ww.IR_STATUS] == 1 and list(di[0]) == [0, 0, 1]
