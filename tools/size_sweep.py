#!/usr/bin/env python
"""Kernel time of one wt_step launch vs ensemble size (fixed per-launch cost = intercept), cost-sorted order."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402
full = ensembles.config5(262144, 10)
for cap in (64, 24):
    for P in (16384, 32768, 65536, 131072, 262144):
        e = full.slice(slice(0, P))
        eng = PlantEnsemble(e, max_attempts=cap, sort_every=1)
        bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
        for _ in range(6):
            eng.step(1.0, bnd)
        torch.cuda.synchronize()
        ms = []
        for _ in range(6):
            eng._order = torch.argsort(eng._cost, descending=True).to(torch.int32)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            so = eng.sort_every; eng.sort_every = 0
            eng.step(1.0, bnd)
            eng.sort_every = so
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        c = eng._cost.cpu().numpy()
        print(f"cap {cap} P {P:7d}: kernel {np.median(ms):.3f} ms  ({1e3*np.median(ms)/P*131072/1e3:.3f} per 131072)  max cost {c.max()} mean cost {c.mean():.2f}", flush=True)
