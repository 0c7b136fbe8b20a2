"""CPU oracle of the ensemble diagnostics -- TEST INFRASTRUCTURE.

Plain-numpy restatement, vectorised over plants, of the reference's per-plant diagnostics:
  IntegratedCSTR.validate_conservation            reactor.py:570-611
  TransportModel.calculate_mixing_quality         transport.py:338-384
  SpatialModel.calculate_spatial_gradients        spatial.py:440-477
  SpatialModel.identify_thermocline               spatial.py:353-379
  SpatialModel.calculate_brunt_vaisala_frequency  spatial.py:322-351
  SpatialModel.calculate_water_density            spatial.py:142-197   (density of the CURRENT temperatures)
  TemperatureDependentKinetics.water_ionization_constant   thermodynamics.py:195-226
Pinned against the unmodified reference by tests/golden/diagnostics_*.npz (oracle/gen_golden_diag.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

import numpy as np

# field order of the per-plant output rows (== WT_DG_* in include/wt_b200.h)
FIELDS = (
    "total_chlorine_mg", "total_H_mol", "total_OH_mol", "charge_balance_mol", "thermal_energy_kJ",
    "chlorine_cv", "chlorine_segregation",
    *[f"{v}_{s}" for v in ("pH", "chlorine", "temperature")
      for s in ("mean_value", "std_value", "max_value", "min_value", "range", "max_gradient", "mean_gradient",
                "gradient_location")],
    "thermocline_depth", "brunt_vaisala_max", "brunt_vaisala_min",
)
NDIAG = len(FIELDS)
G_GRAVITY = 9.81


def density(T):
    """spatial.py:175-195 (salinity 0)."""
    cold = 999.97 - 0.008 * (T - 4.0) ** 2
    warm = 998.2 - 2.1e-4 * 998.2 * (T - 20.0)
    return np.where(T <= 8.0, cold, warm)


def diagnostics(par: np.ndarray, y: np.ndarray, n: int, H: np.ndarray | None = None):
    """par [P, 12] (WT_PAR_* order), y [P, 3n] species-major (pH, chlorine, temperature), optional
    H [P, n] (state.H_concentration; 10**-pH when omitted).  Returns (out [P, NDIAG], n2 [P, n-1], bad [P])
    where bad marks plants whose T[0] is outside [0, 100] C (the reference raises ValueError there)."""
    par = np.asarray(par, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    P = y.shape[0]
    pH, Cl, T = y[:, :n], y[:, n:2 * n], y[:, 2 * n:]
    if H is None:
        H = np.power(10.0, -pH)
    zone_volume = par[:, 9] / n                       # reactor.py:579
    zh = par[:, 7]
    height = zh * n
    out = np.full((P, NDIAG), np.nan)
    f = {k: i for i, k in enumerate(FIELDS)}
    # ---- validate_conservation (reactor.py:570-611)
    out[:, f["total_chlorine_mg"]] = np.sum(Cl, axis=1) * zone_volume
    total_H = np.sum(H, axis=1) * zone_volume / 1000
    T0 = T[:, 0]
    bad = (T0 < 0.0) | (T0 > 100.0)                   # thermodynamics.py:146-157
    with np.errstate(all="ignore"):
        Kw = 1.0e-14 * np.exp((55900.0 / 8.314) * (1.0 / 298.15 - 1.0 / (T0 + 273.15)))
        total_OH = np.sum(Kw[:, None] / H, axis=1) * zone_volume / 1000
    out[:, f["total_H_mol"]] = total_H
    out[:, f["total_OH_mol"]] = total_OH
    out[:, f["charge_balance_mol"]] = total_H - total_OH
    out[:, f["thermal_energy_kJ"]] = 998.2 * 4184 * (par[:, 9] / 1000) * np.mean(T - 20.0, axis=1) / 1000
    # ---- calculate_mixing_quality(chlorine) (transport.py:338-384)
    mean_c, std_c = np.mean(Cl, axis=1), np.std(Cl, axis=1)
    with np.errstate(all="ignore"):
        out[:, f["chlorine_cv"]] = np.where(mean_c > 0, std_c / mean_c, 0.0)
        seg = np.clip(std_c ** 2 / mean_c ** 2, 0.0, 1.0)
    out[:, f["chlorine_segregation"]] = np.where(mean_c ** 2 > 0.0, seg, 0.0)
    # ---- calculate_spatial_gradients (spatial.py:440-477)
    for name, x in (("pH", pH), ("chlorine", Cl), ("temperature", T)):
        g = np.abs(np.diff(x, axis=1) / zh[:, None])
        out[:, f[f"{name}_mean_value"]] = np.mean(x, axis=1)
        out[:, f[f"{name}_std_value"]] = np.std(x, axis=1)
        out[:, f[f"{name}_max_value"]] = np.max(x, axis=1)
        out[:, f[f"{name}_min_value"]] = np.min(x, axis=1)
        out[:, f[f"{name}_range"]] = np.max(x, axis=1) - np.min(x, axis=1)
        out[:, f[f"{name}_max_gradient"]] = np.max(g, axis=1)
        out[:, f[f"{name}_mean_gradient"]] = np.mean(g, axis=1)
        out[:, f[f"{name}_gradient_location"]] = np.argmax(g, axis=1)
    # ---- identify_thermocline (spatial.py:353-379): None -> NaN
    tg = np.abs(T[:, 1:] - T[:, :-1]) / zh[:, None]
    idx = np.argmax(tg, axis=1)
    mg = tg[np.arange(P), idx]
    depth = height - (idx + 0.5) * zh
    out[:, f["thermocline_depth"]] = np.where((par[:, 11] != 0) & (mg > 0.5), depth, np.nan)
    # ---- Brunt-Vaisala N^2 per interface (spatial.py:322-351) on the density of the current temperatures
    rho = density(T)
    drho_dz = (rho[:, 1:] - rho[:, :-1]) / zh[:, None]
    rho_avg = 0.5 * (rho[:, :-1] + rho[:, 1:])
    n2 = -(G_GRAVITY / rho_avg) * drho_dz
    out[:, f["brunt_vaisala_max"]] = np.max(n2, axis=1)
    out[:, f["brunt_vaisala_min"]] = np.min(n2, axis=1)
    return out, n2, bad
