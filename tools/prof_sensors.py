#!/usr/bin/env python
"""Driver for ncu captures of the sensor kernel in its steady state: P plants x 10 zones, suites calibrated at
t = -2000 s, 100 reads to fill the delay rings, then a few timed reads.   python tools/prof_sensors.py --plants 262144"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402
from ics_wt_physicsengine_b200.sensors import create_realistic_sensor_suite  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--plants", type=int, default=262144)
ap.add_argument("--reads", type=int, default=3)
a = ap.parse_args()
e = ensembles.config5(a.plants, 10)
eng = PlantEnsemble(e)
suite = create_realistic_sensor_suite(eng, seed=20260004)
suite.initialize(-2000.0)
for j in range(100):
    suite.read(None, float(j - 100))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reads + 1)]
ev[0].record()
for i in range(a.reads):
    suite.read(None, float(i))
    ev[i + 1].record()
torch.cuda.synchronize()
print("ms per read:", ["%.3f" % ev[i].elapsed_time(ev[i + 1]) for i in range(a.reads)])
