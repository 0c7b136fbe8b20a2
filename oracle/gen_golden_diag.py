#!/usr/bin/env python
"""Golden vectors of the diagnostics operators, produced by RUNNING THE UNMODIFIED REFERENCE.

Build-container only (needs /root/reference).  For a handful of random plants the reference objects are
built through their public constructors, the state arrays are overwritten with random zone profiles, and
validate_conservation / calculate_mixing_quality / calculate_spatial_gradients / identify_thermocline /
calculate_brunt_vaisala_frequency are called as a user would.

    python oracle/gen_golden_diag.py      # writes tests/golden/diagnostics_48.npz
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("WT_REFERENCE_SRC", "/root/reference/src")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
logging.disable(logging.CRITICAL)

from wt_simulator.core.reactor import IntegratedCSTR, ReactorConfiguration  # noqa: E402

from ics_wt_physicsengine_b200 import ensembles as ens  # noqa: E402
from oracle.wt_diag_oracle import FIELDS  # noqa: E402


def main():
    rows, n2s, ys, cfgs, ns = [], [], [], [], []
    rng = np.random.default_rng(20260777)
    for n, P, seed in ((5, 12, 1), (10, 24, 2), (20, 12, 3)):
        e = ens.config3(P, n, seed=seed) if n == 20 else ens.config2(P, n, seed=seed)
        for p in range(P):
            c = {k: float(e.cfg[p, i]) for i, k in enumerate(ens.CFG_FIELDS)}
            c["enable_thermal_stratification"] = bool(c["enable_thermal_stratification"])
            r = IntegratedCSTR(ReactorConfiguration(n_zones=n, **c))
            # random zone profiles: mixed, stratified, straddling 8 C
            r.state.pH = np.clip(e.pH0[p] + rng.normal(0, 0.3, n), 0.5, 13.5)
            r.state.chlorine = np.abs(e.Cl0[p] + rng.normal(0, 0.4, n)) * (0.0 if p % 11 == 5 else 1.0)
            r.state.temperature = np.clip(e.T0[p] + np.linspace(0, rng.uniform(-6, 6), n) + rng.normal(0, 0.2, n), 0.1, 99.0)
            r.state.update_derived()
            cons = r.validate_conservation()
            cv, seg = r.transport.calculate_mixing_quality(r.state.chlorine)
            out = {"total_chlorine_mg": cons["total_chlorine_mg"], "total_H_mol": cons["total_H_mol"],
                   "total_OH_mol": cons["total_OH_mol"], "charge_balance_mol": cons["charge_balance_mol"],
                   "thermal_energy_kJ": cons["thermal_energy_kJ"], "chlorine_cv": cv, "chlorine_segregation": seg}
            for name, x in (("pH", r.state.pH), ("chlorine", r.state.chlorine), ("temperature", r.state.temperature)):
                g = r.spatial.calculate_spatial_gradients(x, name)
                for k, v in g.items():
                    out[f"{name}_{k}"] = v
            r.spatial.update_density_profile(r.state.temperature)
            tc = r.spatial.identify_thermocline()
            out["thermocline_depth"] = np.nan if tc is None else tc
            n2 = np.array([r.spatial.calculate_brunt_vaisala_frequency(i) for i in range(n - 1)])
            out["brunt_vaisala_max"], out["brunt_vaisala_min"] = n2.max(), n2.min()
            rows.append([float(out[k]) for k in FIELDS])
            n2s.append(np.pad(n2, (0, 19 - len(n2)), constant_values=np.nan))
            ys.append(np.pad(np.concatenate([r.state.pH, r.state.chlorine, r.state.temperature]), (0, 60 - 3 * n),
                             constant_values=np.nan))
            hs = np.pad(r.state.H_concentration, (0, 20 - n), constant_values=np.nan)
            cfgs.append(np.concatenate([e.cfg[p], hs]))
            ns.append(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "diagnostics_48.npz"), out=np.array(rows), n2=np.array(n2s),
                        y=np.array(ys), cfg_h=np.array(cfgs), n_zones=np.array(ns), fields=np.array(FIELDS),
                        numpy_version=np.__version__)
    print("wrote", len(rows), "plants")


if __name__ == "__main__":
    main()
