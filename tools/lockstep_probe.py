#!/usr/bin/env python
"""How much work does warp lock-step add, for several plant orderings?  (scheduling study, GPU)

Per step the kernel emits per-plant path counters; a warp of floor(32/n) plants executes the MAX of its
plants' Newton iterations / factorizations / attempts / Jacobians.  Model cost per warp (SASS instruction
counts of the regions): 1774 nnewton + 1388 nlu/2 + 1512 attempts + 1267 njev + 1922.
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles  # noqa: E402

P, n = int(os.environ.get("PLANTS", 262144)), int(os.environ.get("ZONES", 10))
e = ensembles.config5(P, n)
eng = PlantEnsemble(e, max_attempts=64, sort_every=0)
bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
gpw = 32 // n

def cost_of(c):  # c: [8, P] int
    nfev, njev, nlu, nsteps, nnewton, nrej, nfail, _ = [c[i].double() for i in range(8)]
    att = nsteps + nrej + nfail
    return 1774 * nnewton + 1388 * nlu / 2 + 1512 * att + 1267 * njev + 1922, (nnewton, nlu / 2, att, njev)

def warp_cost(c, order):
    parts = cost_of(c[:, order])[1]
    Pw = (P // gpw) * gpw
    mx = [p[:Pw].reshape(-1, gpw).max(dim=1).values for p in parts]
    tot = 1774 * mx[0] + 1388 * mx[1] + 1512 * mx[2] + 1267 * mx[3] + 1922
    return tot.sum().item() * gpw  # per-plant-equivalent

prev = None
prev2 = None
for s in range(int(os.environ.get("STEPS", 8))):
    eng.reset_counters()
    eng.step(1.0, bnd)
    torch.cuda.synchronize()
    c = eng.counters.clone()
    ideal = cost_of(c)[0].sum().item()
    nat = torch.arange(P, device=c.device)
    out = {"natural": warp_cost(c, nat) / ideal}
    if prev is not None:
        pc, (pn, pl, pa, pj) = cost_of(prev)
        out["sum-key(prev)"] = warp_cost(c, torch.argsort((prev[3] + prev[5] + prev[6] + prev[4]), descending=True)) / ideal
        out["model-cost(prev)"] = warp_cost(c, torch.argsort(pc, descending=True)) / ideal
        sig = ((pa.long() * 64 + pn.long()) * 16 + pl.long()) * 4 + pj.long()
        out["signature(prev)"] = warp_cost(c, torch.argsort(sig, descending=True)) / ideal
        sig2 = ((pn.long() * 16 + pl.long()) * 64 + pa.long()) * 4 + pj.long()
        out["sig-newton-first(prev)"] = warp_cost(c, torch.argsort(sig2, descending=True)) / ideal
    if prev2 is not None:
        pc2 = cost_of(prev2)[0]
        out["max(prev,prev2)"] = warp_cost(c, torch.argsort(torch.maximum(pc, pc2), descending=True)) / ideal
        out["prev+0.5prev2"] = warp_cost(c, torch.argsort(pc + 0.5 * pc2, descending=True)) / ideal
        out["lexi(prev,prev2)"] = warp_cost(c, torch.argsort(pc * 1e6 + pc2, descending=True)) / ideal
    out["oracle(same step signature)"] = warp_cost(c, torch.argsort(cost_of(c)[0], descending=True)) / ideal
    print(s, " ".join(f"{k}={v:.3f}" for k, v in out.items()), flush=True)
    prev2 = prev
    prev = c
