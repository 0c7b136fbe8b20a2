#!/usr/bin/env python
"""Small driver for ncu captures of the step kernel: P plants x n zones, a few step() launches.

    python tools/prof_step.py --plants 262144 --zones 10 --steps 3 --warmup 2
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--plants", type=int, default=262144)
ap.add_argument("--zones", type=int, default=10)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--config", default="config5")
ap.add_argument("--max-attempts", type=int, default=64)
ap.add_argument("--sort-every", type=int, default=0)
a = ap.parse_args()
e = getattr(ensembles, a.config)(a.plants, a.zones)
eng = PlantEnsemble(e, max_attempts=a.max_attempts, sort_every=a.sort_every)
bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
for _ in range(a.warmup):
    eng.step(1.0, bnd)
torch.cuda.synchronize()
eng.reset_counters()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
ev[0].record()
for i in range(a.steps):
    eng.step(1.0, bnd)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
c = eng.counters.sum(dim=1).cpu().numpy() / (a.plants * a.steps)
print("ms per step:", ["%.3f" % m for m in ms], "zone-steps/s: %.3e" % (a.plants * a.zones / (min(ms) * 1e-3)))
print("counters per plant-step:", dict(zip(("nfev", "njev", "nlu", "nsteps", "nnewton", "nreject", "nfail", "retry"), c.round(3))))
print("halted:", int(((eng.status & 130) != 0).sum()))
