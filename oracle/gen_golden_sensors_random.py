#!/usr/bin/env python
"""Golden DISTRIBUTION samples of the sensor suite on RANDOM config-5 plants, produced by RUNNING THE UNMODIFIED
REFERENCE (build container only).

N plants of ensembles.config5 (different full scales, calibration references, temperatures, zone profiles), each
with its own suite, calibrated as __main__.initialize_sensors does (__main__.py:96-105) at T0, read the plant's
(frozen) initial state at T0 + 1790 + k, k = 0..K: every sensor except the pH pair is warm from the first read,
the pH pair wakes up at k = 10 on the shared delay line.  Two variant sets:
  "standard"  the factory's suite (RTD PT100, magnetic flow meter)                    sensors/__init__.py:41-120
  "variants"  thermocouple K temperature sensors and a turbine flow meter built with the factory's arguments
              (temperature_sensor.py:173-194, flow_sensor.py:180-199); flow_main is reset() at k = 40 (warm again
              at k = 50, then CALIBRATION_EXPIRED because its calibration history is empty) and calibrated again at
              k = 60 (base_sensor.py:858-878) -- reset stamps time.monotonic(), so the two time attributes are
              overwritten with the simulated time right after the call.
The reference seeds every sensor from secrets.randbits (base_sensor.py:331): only distributions are comparable.

    python oracle/gen_golden_sensors_random.py [N]   # writes tests/golden/sensors_random_plants.npz
"""
import logging
import multiprocessing as mp
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

T0 = 500.0
T_FIRST = T0 + 1790.0
CHECKS = (0, 5, 12, 45, 55, 62, 130)   # read index k (time T_FIRST + k)
K = max(CHECKS)
K_RESET, K_RECAL = 40, 60
SEED = 20260004
NAMES = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")


class _State:
    __slots__ = ("pH", "chlorine", "temperature", "flow_rate")


def build_suite(cfg, variants):
    from wt_simulator.sensors import create_realistic_sensor_suite
    s = create_realistic_sensor_suite(cfg)
    if variants:
        from wt_simulator.sensors.flow_sensor import FlowSensor, FlowSensorType
        from wt_simulator.sensors.temperature_sensor import TemperatureSensor, TemperatureSensorType
        inst = s["pH_inlet"].installation
        s["flow_main"] = FlowSensor(name="flow_main", sensor_type=FlowSensorType.TURBINE, full_scale=cfg.flow_rate * 2.0,
                                    installation=inst)
        s["temp_inlet"] = TemperatureSensor(name="temp_inlet", zone_index=0, sensor_type=TemperatureSensorType.THERMOCOUPLE_K,
                                            sample_line=s["pH_inlet"].sample_line, installation=inst)
        s["temp_outlet"] = TemperatureSensor(name="temp_outlet", zone_index=-1, sensor_type=TemperatureSensorType.THERMOCOUPLE_K,
                                             sample_line=s["pH_outlet"].sample_line, installation=inst)
    return s


def worker(args):
    variants, cfg_rows, pH0, Cl0, T0s, flow = args
    from wt_simulator.core.reactor import ReactorConfiguration
    from wt_simulator.sensors import SensorFault, SensorStatus
    smap, fmap = {s: i for i, s in enumerate(SensorStatus)}, {f: i for i, f in enumerate(SensorFault)}
    n_inst = len(cfg_rows)
    vals = np.full((len(CHECKS), 7, n_inst), np.nan)
    raw = np.full((len(CHECKS), 7, n_inst), np.nan)
    stat = np.zeros((len(CHECKS), 7, n_inst), dtype=np.int8)
    flt = np.zeros((len(CHECKS), 7, n_inst), dtype=np.int8)
    for i in range(n_inst):
        cfg = ReactorConfiguration(flow_rate=float(cfg_rows[i][0]), initial_chlorine=float(cfg_rows[i][1]),
                                   temperature=float(cfg_rows[i][2]))
        sensors = build_suite(cfg, variants)
        for name, s in sensors.items():       # __main__.py:96-105
            if "pH" in name:
                s.calibrate(7.0, T0, "system_init")
            elif "chlorine" in name:
                s.calibrate(cfg.initial_chlorine, T0, "system_init")
            elif "temp" in name:
                s.calibrate(cfg.temperature, T0, "system_init")
            elif "flow" in name:
                s.calibrate(cfg.flow_rate, T0, "system_init")
        st = _State()
        st.pH, st.chlorine, st.temperature, st.flow_rate = pH0[i], Cl0[i], T0s[i], float(flow[i])
        ci = 0
        for k in range(K + 1):
            t = T_FIRST + k
            if variants and k == K_RESET:
                s = sensors["flow_main"]
                s.reset()
                s.last_calibration_time = t   # reset() stamps time.monotonic(); the ensemble runs on simulated time
                s.power_on_time = t
            if variants and k == K_RECAL:
                sensors["flow_main"].calibrate(cfg.flow_rate, t, "operator")
            rd = {name: s.read(st, t) for name, s in sensors.items()}
            if ci < len(CHECKS) and k == CHECKS[ci]:
                for j, name in enumerate(NAMES):
                    vals[ci, j, i] = rd[name].value
                    raw[ci, j, i] = rd[name].raw_value
                    stat[ci, j, i] = smap[rd[name].status]
                    flt[ci, j, i] = fmap[rd[name].fault]
                ci += 1
    return vals, raw, stat, flt


def main():
    from ics_wt_physicsengine_b200 import ensembles
    from ics_wt_physicsengine_b200.ensembles import CFG_FIELDS
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
    e = ensembles.config5(N, 10, seed=SEED)
    col = lambda k: e.cfg[:, CFG_FIELDS.index(k)]
    cfg_rows = np.stack([col("flow_rate"), col("initial_chlorine"), col("temperature")], axis=1)
    flow = col("flow_rate")
    nproc = os.cpu_count() or 1
    out = {}
    for tag, variants in (("standard", False), ("variants", True)):
        idx = np.array_split(np.arange(N), nproc * 4)
        jobs = [(variants, cfg_rows[i], e.pH0[i], e.Cl0[i], e.T0[i], flow[i]) for i in idx if len(i)]
        with np.errstate(all="ignore"), mp.Pool(nproc) as pool:
            res = pool.map(worker, jobs)
        out[tag] = [np.concatenate([r[j] for r in res], axis=2) for j in range(4)]
        v = out[tag][0]
        print(tag, v.shape)
        for ci, k in enumerate(CHECKS):
            print(k, [f"{np.nanmean(v[ci, j]):.3f}/{np.isnan(v[ci, j]).mean():.3f}" for j in range(7)])
    from wt_simulator.sensors import SensorFault, SensorStatus
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "sensors_random_plants.npz"),
        checks=np.array(CHECKS), t0=T0, t_first=T_FIRST, k_reset=K_RESET, k_recal=K_RECAL, seed=SEED, n=N, names=np.array(NAMES),
        std_values=out["standard"][0].astype(np.float32), std_raw=out["standard"][1].astype(np.float32),
        std_status=out["standard"][2], std_fault=out["standard"][3],
        var_values=out["variants"][0].astype(np.float32), var_raw=out["variants"][1].astype(np.float32),
        var_status=out["variants"][2], var_fault=out["variants"][3],
        status_names=np.array([s.value for s in SensorStatus]), fault_names=np.array([f.value for f in SensorFault]),
        numpy_version=np.__version__, scipy_version=scipy.__version__)


if __name__ == "__main__":
    main()
