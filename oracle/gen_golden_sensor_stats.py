#!/usr/bin/env python
"""Golden vectors of BaseSensor.get_statistics from the UNMODIFIED reference: a pH sensor's reading_history is
filled with SensorReading objects of chosen values / timestamps and get_statistics(window) is called.
Build-container only.      python oracle/gen_golden_sensor_stats.py -> tests/golden/sensor_statistics.npz"""
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
from wt_simulator.sensors.base_sensor import SensorFault, SensorReading, SensorStatus  # noqa: E402

from oracle.wt_sensor_stats_oracle import FIELDS  # noqa: E402


def main():
    rng = np.random.default_rng(99)
    K, P = 40, 64
    ts = np.cumsum(rng.choice([1.0, 1.0, 2.0, 5.0], size=K))
    vals = rng.normal(7.0, 0.05, size=(K, P))
    vals[rng.random((K, P)) < 0.15] = np.nan
    vals[:, 3] = np.nan                      # a plant whose sensor never produced a finite value
    drifts = rng.normal(0, 0.01, size=(K, P))
    windows = [0.5, 10.0, 60.0, 1e6]
    out = np.zeros((len(windows), 7, P))
    for p in range(P):
        s = create_realistic_sensor_suite(ReactorConfiguration())["pH_inlet"]
        for k in range(K):
            s.reading_history.append(SensorReading(timestamp=float(ts[k]), value=float(vals[k, p]), raw_value=float(vals[k, p]),
                                                   noise=0.0, drift=float(drifts[k, p]), status=SensorStatus.NORMAL,
                                                   uncertainty=0.02, fault=SensorFault.NONE))
        for w, win in enumerate(windows):
            st = s.get_statistics(win)
            out[w, :, p] = [st[k] for k in FIELDS]
    empty = create_realistic_sensor_suite(ReactorConfiguration())["pH_inlet"].get_statistics(60.0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sensor_statistics.npz"), timestamps=ts, values=vals,
                        windows=np.array(windows), out=out, empty=np.array([empty[k] for k in FIELDS]), fields=np.array(FIELDS))
    print("wrote", out.shape)


if __name__ == "__main__":
    main()
