#!/usr/bin/env python
"""Fixture of the plant-steps that exhaust the engine's attempt budget (tests/golden/overrun_plants.npz).

BASELINE configs[4] plants whose zone temperatures sit on the 8 C density discontinuity (spatial.py:177-189) feeding the
hard Richardson switch (spatial.py:266-275) make scipy's adaptive Radau grind through 1e3 .. 1e7 RHS evaluations for ONE
step(1.0).  This script finds such plant-steps with the CPU oracle (first 16,384 plants of config5(65536), 60 steps, budget 64),
stores their inputs, and records what the reference's own adaptive control does with them:

  * `y_exact`   the oracle with a budget of 200,000 collocation solves (status WORK_LIMIT where even that is not enough);
  * `y_ref`     the UNMODIFIED REFERENCE (IntegratedCSTR.step) on the few cheapest of them, run here (needs
                /root/reference; the GPU box does not have it) -- it shows the same thing as the oracle: the accepted
                trajectory leaves the range spanned by the zone, inlet and ambient temperatures by ~1 K within one
                second, which the (purely diffusive + relaxing) temperature equation cannot do;
  * nothing about floor mode: that is engine policy, the tests compute it.

    python oracle/gen_golden_overruns.py
"""
from __future__ import annotations

import logging
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

from ics_wt_physicsengine_b200 import ensembles as ens  # noqa: E402
from oracle import wt_oracle as wo  # noqa: E402

P, n, STEPS, BUDGET = 16384, 10, 60, 64
N_REF = 4   # plant-steps also run through the reference itself (the cheapest ones)


def main():
    e = ens.config5(65536, n).slice(slice(0, P))
    par = wo.derive_params(e.cfg, n)
    bnd = np.ascontiguousarray(e.bnd)
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1).copy()
    t = np.zeros(P)
    wo.set_floor_div(0)
    wo.set_max_attempts(BUDGET)
    halted = np.zeros(P, bool)
    found = []
    for s in range(STEPS):
        idx = np.nonzero(~halted)[0]
        ya, ta = y[idx].copy(), t[idx].copy()
        y0 = ya.copy()
        st, _, _ = wo.step_batch(par[idx].copy(), bnd[idx].copy(), n, ta, ya, dt=1.0, nthreads=os.cpu_count() or 1)
        y[idx], t[idx] = ya, ta
        bad = (st & wo.ST_WORK_LIMIT) != 0
        for k in np.nonzero(bad)[0]:
            found.append((int(idx[k]), s, y0[k].copy()))
        halted[idx[bad]] = True
    pl = np.array([f[0] for f in found])
    y0 = np.array([f[2] for f in found])
    t0 = np.array([float(f[1]) for f in found])
    print(f"{len(found)} plant-steps over the budget of {BUDGET}")
    # what the adaptive control of the reference does with them (oracle, large budget)
    wo.set_max_attempts(200000)
    ye, te = y0.copy(), t0.copy()
    st_e, cnt_e, _ = wo.step_batch(par[pl].copy(), bnd[pl].copy(), n, te, ye, dt=1.0, nthreads=os.cpu_count() or 1)
    wo.set_max_attempts(0)
    # ... and the reference itself on the cheapest ones
    from gen_golden import make_plant  # noqa: E402  (same helper as the other goldens)
    order = np.argsort(cnt_e[:, 0] + 10**9 * ((st_e & wo.ST_WORK_LIMIT) != 0))
    ref_rows = order[:N_REF]
    y_ref = np.full((len(ref_rows), 3 * n), np.nan)
    for j, k in enumerate(ref_rows):
        r, b = make_plant(e, int(pl[k]))
        r.state.pH, r.state.chlorine, r.state.temperature = y0[k, :n].copy(), y0[k, n:2 * n].copy(), y0[k, 2 * n:].copy()
        r.state.time = float(t0[k])
        r.step(1.0, b)
        y_ref[j] = np.concatenate([r.state.pH, r.state.chlorine, r.state.temperature])
        print(f"reference on plant {pl[k]}: max |T - T0| = {np.abs(y_ref[j, 2 * n:] - y0[k, 2 * n:]).max():.3f} K, "
              f"oracle vs reference rel {np.max(np.abs(y_ref[j] - ye[k]) / np.abs(y_ref[j])):.2e}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "overrun_plants.npz"),
                        plant=pl, t0=t0, y0=y0, par=par[pl], bnd=bnd[pl], n_zones=n, budget=BUDGET,
                        y_exact=ye, status_exact=st_e, counters_exact=cnt_e, ref_rows=ref_rows, y_ref=y_ref,
                        numpy_version=np.__version__, scipy_version=scipy.__version__)


if __name__ == "__main__":
    main()
