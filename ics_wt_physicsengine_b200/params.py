"""Per-plant derived constants, computed on the host exactly as the reference constructors do.

Mirrors (vectorised over plants, same operation order):
  thermodynamics.py:219-226  Kw(T)            chemistry.py:123-132  Ka1, Ka2, Ka_HOCl
  thermodynamics.py:320-331  D_molecular(T)   transport.py:221-242, 282-290  velocity, K_exchange
  reactor.py:428-430         heat-loss area
All chemistry constants use the CONFIGURATION temperature (chemistry.py:116-132), never the zone
temperature -- a reference quirk the engine keeps.
"""
from __future__ import annotations

import numpy as np

from .ensembles import CFG_FIELDS, NCFG

_C = {k: i for i, k in enumerate(CFG_FIELDS)}
NPAR = 12
PAR_FIELDS = ("Kw", "Ka1", "Ka2", "Ka_HOCl", "C_T", "K_exchange_per_s", "superficial_velocity",
              "zone_height", "zone_volume_L", "volume", "heat_loss_area", "stratification")


def validate_cfg(cfg: np.ndarray, n_zones: int) -> None:
    """ReactorConfiguration.validate (reactor.py:91-110) + constructor-time checks, batched."""
    cfg = np.asarray(cfg, dtype=np.float64).reshape(-1, NCFG)
    if n_zones < 2:
        raise ValueError(f"Need at least 2 zones, got {n_zones}")  # transport.py:88-89
    if n_zones > 32:
        raise ValueError("the B200 engine maps zones to lanes of one warp: n_zones <= 32")
    vol, h, d = cfg[:, _C["volume"]], cfg[:, _C["height"]], cfg[:, _C["diameter"]]
    calc = np.pi * (d / 2) ** 2 * h * 1000
    err = np.abs(calc - vol) / vol
    if np.any(err > 0.01):
        p = int(np.argmax(err))
        raise ValueError(f"Volume mismatch: specified {vol[p]}L, calculated {calc[p]:.1f}L from geometry. "
                         f"Error: {err[p] * 100:.1f}% (plant {p})")
    assert np.all((0 < vol) & (vol < 1e6)), "Volume out of range"
    fr = cfg[:, _C["flow_rate"]]
    assert np.all((0 <= fr) & (fr < 1e5)), "Flow rate out of range (use 0 for batch mode)"
    if np.any(fr == 0):
        # reactor.py:226 formats residence_time=None -> TypeError in the reference constructor
        raise TypeError("flow_rate == 0 crashes the reference constructor (reactor.py:226); "
                        "use boundary.inlet_flow_rate = 0 for batch operation")
    ph = cfg[:, _C["initial_pH"]]
    assert np.all((0 <= ph) & (ph <= 14)), "pH out of range"
    cl = cfg[:, _C["initial_chlorine"]]
    assert np.all((0 <= cl) & (cl <= 10)), "Chlorine out of range"
    t = cfg[:, _C["temperature"]]
    assert np.all((0 <= t) & (t <= 40)), "Temperature out of typical range"
    if np.any(cfg[:, _C["alkalinity"]] < 0):
        raise ValueError("Alkalinity cannot be negative")  # chemistry.py:70-71
    if np.any(cfg[:, _C["total_carbonate"]] < 0):
        raise ValueError("Total carbonate cannot be negative")  # chemistry.py:72-75


def derive_params(cfg: np.ndarray, n_zones: int) -> np.ndarray:
    """cfg [P, NCFG] -> par [P, NPAR] (float64), index order = WT_PAR_* of include/wt_b200.h."""
    cfg = np.asarray(cfg, dtype=np.float64).reshape(-1, NCFG)
    P = cfg.shape[0]
    n = n_zones
    Tc = cfg[:, _C["temperature"]]
    TK = Tc + 273.15
    par = np.empty((P, NPAR), dtype=np.float64)
    par[:, 0] = 1.0e-14 * np.exp((55900.0 / 8.314) * (1.0 / 298.15 - 1.0 / TK))
    par[:, 1] = np.power(10.0, -(6.35 + (-0.008) * (Tc - 25.0)))
    par[:, 2] = np.power(10.0, -(10.33 + (-0.008) * (Tc - 25.0)))
    par[:, 3] = np.power(10.0, -(7.5 + 0.01 * (Tc - 25.0)))
    par[:, 4] = cfg[:, _C["total_carbonate"]] / 1000.0
    d = cfg[:, _C["diameter"]]
    area = np.pi * (d / 2) ** 2
    zh = cfg[:, _C["height"]] / n
    vz = cfg[:, _C["volume"]] / n
    q_m3_s = cfg[:, _C["flow_rate"]] / 60000.0
    par[:, 6] = q_m3_s / area
    n_rps = cfg[:, _C["impeller_speed"]] / 60.0
    d_turb = 0.1 * n_rps * cfg[:, _C["impeller_diameter"]] ** 2
    expo = 1800.0 * (1.0 / TK - 1.0 / 293.15)
    d_mol = 1.0e-9 * (TK / 293.15) * np.exp(-expo)
    d_eff = d_turb + d_mol
    k_exchange = d_eff * area / zh
    par[:, 5] = k_exchange / (vz / 1000.0)
    par[:, 7] = zh
    par[:, 8] = vz
    par[:, 9] = cfg[:, _C["volume"]]
    par[:, 10] = np.pi * d * cfg[:, _C["height"]] + 2 * np.pi * (d / 2) ** 2
    par[:, 11] = (cfg[:, _C["enable_thermal_stratification"]] != 0).astype(np.float64)
    return par
