#!/usr/bin/env python
"""Tiny end-to-end run (step, advance, derivatives, sensors, stats, calc_pH) for compute-sanitizer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, calculate_pH_batch, ensembles  # noqa: E402
from ics_wt_physicsengine_b200.partition import EnsembleStatistics  # noqa: E402
from ics_wt_physicsengine_b200.sensors import create_realistic_sensor_suite  # noqa: E402

for n, P in ((10, 301), (20, 77), (5, 64), (32, 5), (2, 33)):
    e = ensembles.config2(P, n, seed=n)
    eng = PlantEnsemble(e, sort_every=1)
    suite = create_realistic_sensor_suite(eng, seed=3)
    suite.initialize(0.0)
    st = EnsembleStatistics(eng)
    for k in range(4):
        eng.step(1.0, e.bnd)
        suite.read(eng.state, float(k))
    eng.advance(2, 1.0, e.bnd)
    eng.derivatives(e.bnd)
    v = st.local()
    torch.cuda.synchronize()
    print(n, P, float(v[0]), float(eng.state.time.max()))
alk, ct, temp, guess = ensembles.config4(2000)
ph, it, s = calculate_pH_batch(alk, ct, temp, guess)
torch.cuda.synchronize()
print("ok", int(it.max()))
