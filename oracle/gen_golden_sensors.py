#!/usr/bin/env python
"""Golden DISTRIBUTION samples for the sensor suite, produced by RUNNING THE UNMODIFIED REFERENCE.

N independent instances of create_realistic_sensor_suite (sensors/__init__.py:41-120), calibrated
exactly as __main__.initialize_sensors does (__main__.py:96-105), read the trajectory of the
default plant (reactor.step(1.0) with default boundary) at t0 + k for k = 0..K.  The reference
seeds every sensor from secrets.randbits (base_sensor.py:331), so runs are irreproducible by
design: what is committed are the per-instance readings at a few check times, against which the
engine's (and the oracle port's) sensor outputs are compared IN DISTRIBUTION (moments + KS).

    python oracle/gen_golden_sensors.py [N]     # writes tests/golden/sensors_default_plant.npz
"""
import logging
import multiprocessing as mp
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

T0 = 1000.0
CHECKS = (5, 15, 40, 100, 400, 1805, 1840, 1900)   # read index k (time t0 + k)
K = max(CHECKS)
NAMES = ("pH_inlet", "pH_outlet", "chlorine_inlet", "chlorine_outlet", "flow_main", "temp_inlet", "temp_outlet")


class _State:
    __slots__ = ("pH", "chlorine", "temperature", "flow_rate")


def trajectory():
    from wt_simulator.core.reactor import BoundaryConditions, IntegratedCSTR, ReactorConfiguration
    cfg = ReactorConfiguration()
    r = IntegratedCSTR(cfg)
    b = BoundaryConditions()
    out = []
    for _ in range(K + 1):
        s = r.step(1.0, b)
        st = _State()
        st.pH, st.chlorine, st.temperature, st.flow_rate = s.pH.copy(), s.chlorine.copy(), s.temperature.copy(), float(s.flow_rate)
        out.append(st)
    return cfg, out


def status_codes():
    from wt_simulator.sensors import SensorFault, SensorStatus
    return {s: i for i, s in enumerate(SensorStatus)}, {f: i for i, f in enumerate(SensorFault)}


def worker(args):
    n_inst, traj_arrays = args
    from wt_simulator.core.reactor import ReactorConfiguration
    from wt_simulator.sensors import create_realistic_sensor_suite
    smap, fmap = status_codes()
    cfg = ReactorConfiguration()
    traj = []
    for pH, cl, T, fl in traj_arrays:
        st = _State()
        st.pH, st.chlorine, st.temperature, st.flow_rate = pH, cl, T, fl
        traj.append(st)
    vals = np.full((len(CHECKS), 7, n_inst), np.nan)
    stat = np.zeros((len(CHECKS), 7, n_inst), dtype=np.int8)
    flt = np.zeros((len(CHECKS), 7, n_inst), dtype=np.int8)
    for i in range(n_inst):
        sensors = create_realistic_sensor_suite(cfg)
        for name, s in sensors.items():       # __main__.py:96-105
            if "pH" in name:
                s.calibrate(7.0, T0, "system_init")
            elif "chlorine" in name:
                s.calibrate(cfg.initial_chlorine, T0, "system_init")
            elif "temp" in name:
                s.calibrate(cfg.temperature, T0, "system_init")
            elif "flow" in name:
                s.calibrate(cfg.flow_rate, T0, "system_init")
        ci = 0
        for k in range(K + 1):
            rd = {name: s.read(traj[k], T0 + k) for name, s in sensors.items()}
            if ci < len(CHECKS) and k == CHECKS[ci]:
                for j, name in enumerate(NAMES):
                    vals[ci, j, i] = rd[name].value
                    stat[ci, j, i] = smap[rd[name].status]
                    flt[ci, j, i] = fmap[rd[name].fault]
                ci += 1
    return vals, stat, flt


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 10240
    cfg, traj = trajectory()
    arrays = [(s.pH, s.chlorine, s.temperature, s.flow_rate) for s in traj]
    nproc = os.cpu_count() or 1
    chunks = [N // (nproc * 4)] * (nproc * 4)
    chunks[-1] += N - sum(chunks)
    with np.errstate(all="ignore"), mp.Pool(nproc) as pool:
        res = pool.map(worker, [(c, arrays) for c in chunks if c > 0])
    vals = np.concatenate([r[0] for r in res], axis=2)
    stat = np.concatenate([r[1] for r in res], axis=2)
    flt = np.concatenate([r[2] for r in res], axis=2)
    from wt_simulator.sensors import SensorFault, SensorStatus
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", "sensors_default_plant.npz"),
        values=vals, status=stat, fault=flt, checks=np.array(CHECKS), t0=T0, names=np.array(NAMES),
        status_names=np.array([s.value for s in SensorStatus]), fault_names=np.array([f.value for f in SensorFault]),
        traj_pH=np.array([s.pH for s in traj]), traj_Cl=np.array([s.chlorine for s in traj]),
        traj_T=np.array([s.temperature for s in traj]), traj_flow=np.array([s.flow_rate for s in traj]),
        numpy_version=np.__version__, scipy_version=scipy.__version__)
    print("wrote", vals.shape)
    for ci, k in enumerate(CHECKS):
        print(k, [f"{np.nanmean(vals[ci, j]):.4f}/{np.isnan(vals[ci, j]).mean():.3f}" for j in range(7)])


if __name__ == "__main__":
    main()
