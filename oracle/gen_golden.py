#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference; the GPU box does not have it).
Nothing from the reference is copied: it is imported, driven through its public API
(IntegratedCSTR.step / derivatives, AqueousChemistry.calculate_pH) on the deterministic
ensembles of ics_wt_physicsengine_b200/ensembles.py, and its outputs are stored as small
.npz fixtures stamped with the numpy / scipy versions (the integrator lives in un-pinned
scipy, so a different scipy build must be detectable).

    python oracle/gen_golden.py            # writes tests/golden/*.npz
"""
from __future__ import annotations

import logging
import os
import sys
import warnings

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

import wt_simulator.core.reactor as ref_reactor  # noqa: E402
from wt_simulator.core.chemistry import AqueousChemistry, BufferSystem  # noqa: E402
from wt_simulator.core.reactor import BoundaryConditions, IntegratedCSTR, ReactorConfiguration  # noqa: E402

from ics_wt_physicsengine_b200 import ensembles as ens  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
STAMP = dict(numpy_version=np.__version__, scipy_version=scipy.__version__)

# ---- capture solve_ivp's path counters without touching the reference sources
_last = {}
_orig_solve_ivp = ref_reactor.solve_ivp


def _spy_solve_ivp(*a, **k):
    sol = _orig_solve_ivp(*a, **k)
    _last["nfev"], _last["njev"], _last["nlu"] = sol.nfev, sol.njev, sol.nlu
    _last["nsteps"] = len(sol.t) - 1
    _last["status"] = sol.status
    return sol


ref_reactor.solve_ivp = _spy_solve_ivp


def make_plant(e: ens.Ensemble, p: int):
    c = {k: float(e.cfg[p, i]) for i, k in enumerate(ens.CFG_FIELDS)}
    c["enable_thermal_stratification"] = bool(c["enable_thermal_stratification"])
    cfg = ReactorConfiguration(n_zones=e.n_zones, **c)
    r = IntegratedCSTR(cfg)
    r.state.pH = e.pH0[p].copy()
    r.state.chlorine = e.Cl0[p].copy()
    r.state.temperature = e.T0[p].copy()
    b = BoundaryConditions(**{k: float(e.bnd[p, i]) for i, k in enumerate(ens.BND_FIELDS)})
    return r, b


def ref_params(r: IntegratedCSTR) -> np.ndarray:
    """Per-plant constants in WT_PAR_* order, read off the reference objects."""
    cfg = r.config
    A_lat = np.pi * cfg.diameter * cfg.height
    A_ends = 2 * np.pi * (cfg.diameter / 2) ** 2
    return np.array([
        r.chemistry.Kw, r.chemistry.Ka1, r.chemistry.Ka2, r.chemistry.Ka_HOCl,
        r.buffer.total_carbonate / 1000.0,
        r.transport.K_matrix[0, 1],
        r.transport.superficial_velocity,
        r.spatial.zone_height,
        cfg.volume / cfg.n_zones,
        cfg.volume,
        A_lat + A_ends,
        1.0 if cfg.enable_thermal_stratification else 0.0,
    ], dtype=np.float64)


def run_trajectories(e: ens.Ensemble, nsteps: int, dt: float, record_every: int = 1):
    P, n = e.n_plants, e.n_zones
    nrec = nsteps // record_every
    Y = np.full((nrec, P, 3 * n), np.nan)
    cnt = np.zeros((nrec, P, 4), dtype=np.int32)   # nfev, njev, nlu, nsteps of the recorded step
    raised = np.full(P, -1, dtype=np.int32)        # step index at which step() raised ValueError
    failed = np.zeros(P, dtype=np.int32)
    par = np.zeros((P, 12))
    for p in range(P):
        r, b = make_plant(e, p)
        par[p] = ref_params(r)
        for s in range(nsteps):
            try:
                st = r.step(dt, b)
            except ValueError:
                raised[p] = s
                break
            if _last["status"] != 0:
                failed[p] += 1
            if (s + 1) % record_every == 0:
                k = (s + 1) // record_every - 1
                Y[k, p] = np.concatenate([st.pH, st.chlorine, st.temperature])
                cnt[k, p] = (_last["nfev"], _last["njev"], _last["nlu"], _last["nsteps"])
    return dict(Y=Y, counters=cnt, raised=raised, failed=failed, par=par, cfg=e.cfg, bnd=e.bnd,
                pH0=e.pH0, Cl0=e.Cl0, T0=e.T0, n_zones=n, dt=dt, nsteps=nsteps,
                record_every=record_every, **STAMP)


def gen_rhs(e: ens.Ensemble, seed: int):
    """derivatives() at perturbed (non-uniform) states."""
    rng = np.random.default_rng(seed)
    P, n = e.n_plants, e.n_zones
    Y = np.zeros((P, 3 * n))
    F = np.zeros((P, 3 * n))
    for p in range(P):
        r, b = make_plant(e, p)
        y = np.concatenate([e.pH0[p] + rng.normal(0, 0.3, n), np.abs(e.Cl0[p] + rng.normal(0, 0.2, n)),
                            np.clip(e.T0[p] + rng.normal(0, 0.5, n), 0.01, 99.9)])
        Y[p] = y
        F[p] = r.derivatives(0.0, y, b)
    return dict(Y=Y, F=F, cfg=e.cfg, bnd=e.bnd, n_zones=n, **STAMP)


def gen_calc_ph(P: int):
    alk, ct, temp, guess = ens.config4(P)
    N = alk.size
    ph = np.full(N, np.nan)
    iters = np.zeros(N, dtype=np.int32)
    status = np.zeros(N, dtype=np.int32)

    # iteration counts are not returned by the reference: wrap its own f / f' to count calls
    for i in range(N):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            chem = AqueousChemistry(BufferSystem(float(alk[i]), float(ct[i]), float(temp[i])))
        calls = [0]
        orig = chem.charge_balance_error

        def counted(pH, _o=orig, _c=calls):
            _c[0] += 1
            return _o(pH)

        chem.charge_balance_error = counted
        try:
            with np.errstate(all="ignore"):
                ph[i] = chem.calculate_pH(initial_guess=float(guess[i]))
            status[i] = 0
        except RuntimeError as ex:
            status[i] = 1 if "Derivative too small" in str(ex) else 2
        iters[i] = calls[0]
    return dict(alk=alk, ct=ct, temp=temp, guess=guess, ph=ph, iters=iters, status=status, **STAMP)


def main():
    os.makedirs(GOLD, exist_ok=True)
    with np.errstate(all="ignore"):
        # config 1: the default plant, one simulated hour (BASELINE configs[0])
        g1 = run_trajectories(ens.config1(5), 3600, 1.0, record_every=100)
        np.savez_compressed(os.path.join(GOLD, "config1_default_3600.npz"), **g1)
        g1b = run_trajectories(ens.config1(5), 20, 1.0)
        np.savez_compressed(os.path.join(GOLD, "config1_default_first20.npz"), **g1b)
        # config 2 slice: 64 random 10-zone plants, 25 steps each
        g2 = run_trajectories(ens.config2(64), 25, 1.0)
        np.savez_compressed(os.path.join(GOLD, "config2_64x10_25.npz"), **g2)
        # config 3 slice: 48 random 20-zone plants (T sweep, stratified), 12 steps each
        g3 = run_trajectories(ens.config3(48), 12, 1.0)
        np.savez_compressed(os.path.join(GOLD, "config3_48x20_12.npz"), **g3)
        # other dt values (step(dt) is an API argument)
        g5 = run_trajectories(ens.config2(16, 10, seed=7), 6, 10.0)
        np.savez_compressed(os.path.join(GOLD, "config2_16x10_dt10.npz"), **g5)
        g6 = run_trajectories(ens.config2(16, 5, seed=8), 10, 0.1)
        np.savez_compressed(os.path.join(GOLD, "config2_16x5_dt01.npz"), **g6)
        # RHS spot values
        np.savez_compressed(os.path.join(GOLD, "rhs_config2.npz"), **gen_rhs(ens.config2(64), 11))
        np.savez_compressed(os.path.join(GOLD, "rhs_config3.npz"), **gen_rhs(ens.config3(64), 12))
        # calculate_pH (BASELINE configs[3] slice)
        np.savez_compressed(os.path.join(GOLD, "calc_ph_4096.npz"), **gen_calc_ph(4096))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
