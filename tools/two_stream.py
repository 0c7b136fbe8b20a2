#!/usr/bin/env python
"""Does running two half-ensembles on two streams hide the per-launch tail?  (plants are independent)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402
P = int(os.environ.get("PLANTS", 131072))
full = ensembles.config5(P, 10)
def mk(sl):
    e = full.slice(sl)
    eng = PlantEnsemble(e, max_attempts=64, sort_every=int(os.environ.get("SORT", 1)))
    return eng, torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
for parts in (1, 2, 4):
    engs = [mk(slice(i * P // parts, (i + 1) * P // parts)) for i in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    def step_all():
        for (eng, bnd), s in zip(engs, streams):
            with torch.cuda.stream(s):
                eng.step(1.0, bnd)
    for _ in range(6):
        step_all()
    torch.cuda.synchronize()
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for _ in range(K):
        step_all()
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print(f"P {P} in {parts} stream(s): {e0.elapsed_time(e1)/K:.3f} ms per step", flush=True)
