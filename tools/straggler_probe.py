#!/usr/bin/env python
"""What would a deferral queue find?  Steps a config-5 ensemble with the default budget of 64 collocation solves per
plant-step, collects the plants the budget halts, and re-runs the very step that halted them (same state, same
boundary) with larger budgets: how many complete, and how many steps after that stay expensive.
    python tools/straggler_probe.py --plants 262144 --steps 60"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--plants", type=int, default=262144)
ap.add_argument("--steps", type=int, default=60)
a = ap.parse_args()
e = ensembles.config5(a.plants, 10)
eng = PlantEnsemble(e, max_attempts=64)
bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
halted_at = {}
for k in range(a.steps):
    eng.step(1.0, bnd)
    st = eng.status.cpu().numpy()
    for p in np.nonzero(st & 128)[0]:
        halted_at.setdefault(int(p), k)
idx = np.array(sorted(halted_at))
print(f"{a.plants} plants x {a.steps} steps, budget 64: {idx.size} plants halted by the budget "
      f"({idx.size / a.plants:.2e}; {idx.size / (a.plants * a.steps):.2e} of plant-steps)")
if idx.size == 0:
    sys.exit(0)
y = eng.state_numpy()[idx]
t = eng.state.time.cpu().numpy()[idx]
sub = e.slice(idx)
for budget in (256, 2048, 16384, 131072):
    s = PlantEnsemble(sub, max_attempts=budget)
    s.set_state(y[:, :10], y[:, 10:20], y[:, 20:], time=t)
    done_steps = np.zeros(idx.size, int)
    cost_first = None
    for k in range(5):
        s.reset_counters()
        s.step(1.0, sub.bnd)
        torch.cuda.synchronize()
        ok = (s.status.cpu().numpy() & 128) == 0
        done_steps += ok & (done_steps == k)
        if k == 0:
            c = s.counters.cpu().numpy()
            cost_first = (c[3] + c[5] + c[6])[ok]   # accepted + rejected + failed collocation solves of the completed ones
    comp = int((done_steps >= 1).sum())
    print(f"  budget {budget:6d}: {comp:4d} of {idx.size} complete the step that halted them"
          + (f" (median {int(np.median(cost_first))} attempts, max {int(cost_first.max())})" if comp else "")
          + f"; still running after 5 more steps with that budget: {int((done_steps >= 5).sum())}")
