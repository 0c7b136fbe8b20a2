#!/usr/bin/env python
"""SURVEY 8(d) "CPU baseline beside it": the UNMODIFIED REFERENCE stepping a slice of the bench ensemble in
multiprocessing.Pool(8) -- one IntegratedCSTR + its 7-sensor suite per plant, step(1.0) + one read of every sensor per
step (the work bench.py's `value` times), BLAS threads = 1, logging disabled.

Runs only in the build container (it imports /root/reference; the GPU box does not have it), so its result is a committed
measurement (profiles/r2_reference_pool8.json), not a bench.py leg: bench.py's CPU arm on the GPU box is the oracle C port.

The sample is bounded at 25 steps on purpose: SURVEY asks for >= 200, but from step 27 on plant 46 of this very slice
sits on the 8 C density discontinuity -- the reference needs 8.4 s for that one step() and did not finish a later one
within the 10 minutes it was given (the script prints such steps).  An aggregate over 200 steps would
measure that one plant.

    python oracle/ref_pool_baseline.py [plants=64] [steps=25]
"""
from __future__ import annotations

import json
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_ZONES, DT, WORKERS = 10, 1.0, 8


def work(args):
    lo, hi, steps, total = args
    import logging
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    logging.disable(logging.CRITICAL)
    from gen_golden import make_plant   # builds the reference's IntegratedCSTR / BoundaryConditions of an ensemble row
    from wt_simulator.sensors import create_realistic_sensor_suite

    from ics_wt_physicsengine_b200 import ensembles as ens
    e = ens.config5(total, N_ZONES)
    plants = []
    for p in range(lo, hi):
        r, b = make_plant(e, p)
        sensors = create_realistic_sensor_suite(r.config)
        for name, s in sensors.items():   # __main__.initialize_sensors (__main__.py:96-105)
            ref = 7.0 if "pH" in name else r.config.initial_chlorine if "chlorine" in name else \
                r.config.temperature if "temp" in name else r.config.flow_rate
            s.calibrate(ref, 0.0, "system_init")
        plants.append((r, b, sensors))
    t0 = time.perf_counter()
    done = 0
    for k in range(steps):
        for r, b, sensors in plants:
            try:
                ts = time.perf_counter()
                st = r.step(DT, b)
                if time.perf_counter() - ts > 5.0:   # a plant on the 8 C density discontinuity
                    print(f"plant {lo + plants.index((r, b, sensors))} step {k}: {time.perf_counter() - ts:.1f} s for one step()", flush=True)
                for s in sensors.values():
                    s.read(st, float(k + 1))
                done += 1
            except ValueError:
                pass
    return done, time.perf_counter() - t0


def main():
    import multiprocessing as mp
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    total = 1048576
    per = (P + WORKERS - 1) // WORKERS
    jobs = [(i * per, min(P, (i + 1) * per), steps, total) for i in range(WORKERS) if i * per < P]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(WORKERS) as pool:
        res = pool.map(work, jobs)
    wall = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    el = max(r[1] for r in res)   # the stepping loops run side by side; construction is not timed
    out = {"metric": "plant-zone-steps/sec", "value": done * N_ZONES / el, "unit": "plant-zone-steps/s", "kind": "reference",
           "cores": WORKERS, "sample": f"first {P} plants of config5({total}) x {steps} steps, IntegratedCSTR.step(1.0) + 7 sensor reads "
           f"per plant per step, multiprocessing.Pool({WORKERS}), BLAS threads 1", "plant_steps_completed": done,
           "seconds_stepping_max_over_workers": el, "seconds_wall_with_construction": wall, "host_cpus": os.cpu_count(),
           "numpy_version": np.__version__, "scipy_version": scipy.__version__, "where": "build container (no GPU)"}
    print(json.dumps(out))
    with open(os.path.join(ROOT, "profiles", "r2_reference_pool8.json"), "w") as f:
        f.write(json.dumps(out) + "\n")


if __name__ == "__main__":
    sys.path.insert(0, os.environ.get("WT_REFERENCE_SRC", "/root/reference/src"))
    os.environ["PYTHONPATH"] = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src") + os.pathsep + os.environ.get("PYTHONPATH", "")
    main()
