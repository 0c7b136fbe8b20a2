#!/usr/bin/env python
"""Golden vectors of the sensor maintenance operations, produced by RUNNING THE UNMODIFIED REFERENCE.

For the suite of create_realistic_sensor_suite(ReactorConfiguration()) the relevant attributes of a sensor are
set to drawn values, the maintenance method is called as a user would, and the attributes are read back.
Build-container only.      python oracle/gen_golden_maint.py   ->  tests/golden/sensor_maintenance.npz
"""
import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

from wt_simulator.core.reactor import ReactorConfiguration  # noqa: E402
from wt_simulator.sensors import create_realistic_sensor_suite  # noqa: E402

NAMES = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")
ATTRS = ("current_value", "calibration_offset", "last_calibration_time", "power_on_time", "membrane_fouling",
         "reference_contamination", "days_since_cleaning", "membrane_age_days", "reagent_potency",
         "light_exposure_hours", "reagent_age_days")


def dump(s):
    row = [float(getattr(s, a, 0.0)) for a in ATTRS]
    enum_index = lambda e: list(type(e)).index(e)
    return row + [float(enum_index(s.status)), float(enum_index(s.fault))]


def main():
    rng = np.random.default_rng(424242)
    cases = []   # (sensor index, op, t, args[4], before[13], after[13], raised)
    for trial in range(60):
        suite = create_realistic_sensor_suite(ReactorConfiguration())
        for si, op in ((0, 0), (1, 0), (0, 1), (1, 1), (2, 2), (3, 3), (3, 2), (2, 3)):
            s = suite[NAMES[si]]
            s.current_value = float(rng.uniform(s.min_value, s.max_value))
            s.calibration_offset = float(rng.normal(0, 0.1))
            s.last_calibration_time = float(rng.uniform(0, 100))
            s.power_on_time = float(rng.uniform(0, 100))
            for a in ("membrane_fouling", "reference_contamination", "days_since_cleaning", "membrane_age_days",
                      "light_exposure_hours", "reagent_age_days"):
                if hasattr(s, a):
                    setattr(s, a, float(rng.uniform(0, 0.5)))
            if hasattr(s, "reagent_potency"):
                s.reagent_potency = float(rng.uniform(0.5, 1.0))
            t = float(rng.uniform(200, 5000))
            args = [0.0, 0.0, 0.0, 0.0]
            before = dump(s)
            raised = 0
            try:
                if op == 0:
                    args = [7.0, float(rng.choice([4.0, 10.0])), float(rng.normal(7.0, 0.05)), float(rng.normal(4.0, 0.05))]
                    s.calibrate_two_point(args[0], args[1], args[2], args[3], current_time=t)
                elif op == 1:
                    m = int(rng.integers(0, 3))
                    args[0] = float(m)
                    s.clean_electrode(("water_rinse", "acid_clean", "pepsin_clean")[m], current_time=t)
                elif op == 2:
                    s.replace_membrane(current_time=t)
                else:
                    s.replace_reagent(current_time=t)
            except ValueError:
                raised = 1
            cases.append([si, op, t, *args, *before, *dump(s), raised])
    out = np.array(cases)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sensor_maintenance.npz"), cases=out, attrs=np.array(ATTRS))
    print("wrote", len(cases), "cases; raised:", int(out[:, -1].sum()))


if __name__ == "__main__":
    main()
