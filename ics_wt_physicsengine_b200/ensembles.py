"""Deterministic synthetic plant ensembles for the BASELINE.json configurations.

Host-side numpy only.  The definitions follow SURVEY.md section 8(d); the same generators
feed the parity tests, the golden-vector script (oracle/gen_golden.py) and bench.py, so the
CUDA engine, the CPU oracle and the unmodified reference all see identical inputs.

Field order of the `cfg` and `bnd` matrices is the field order of the reference dataclasses
``ReactorConfiguration`` (reactor.py:52-89) and ``BoundaryConditions`` (reactor.py:150-186).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

CFG_FIELDS = (
    "volume", "height", "diameter", "flow_rate", "turbulent_intensity", "recirculation_ratio",
    "impeller_speed", "impeller_diameter", "power_number", "initial_pH", "alkalinity",
    "total_carbonate", "initial_chlorine", "temperature", "enable_thermal_stratification",
)
BND_FIELDS = (
    "inlet_flow_rate", "inlet_pH", "inlet_chlorine", "inlet_temperature", "acid_flow_rate",
    "acid_concentration", "chlorine_flow_rate", "chlorine_concentration", "ambient_temperature",
    "heat_loss_coefficient",
)
NCFG = len(CFG_FIELDS)
NBND = len(BND_FIELDS)
_C = {k: i for i, k in enumerate(CFG_FIELDS)}
_B = {k: i for i, k in enumerate(BND_FIELDS)}

SEED_CONFIG2 = 20260001
SEED_CONFIG3 = 20260002
SEED_CONFIG4 = 20260003
SEED_CONFIG5 = 20260004


@dataclass
class Ensemble:
    """P plants x n zones: configuration, boundary and initial state (all float64)."""

    n_zones: int
    cfg: np.ndarray  # [P, NCFG]
    bnd: np.ndarray  # [P, NBND]
    pH0: np.ndarray  # [P, n]
    Cl0: np.ndarray  # [P, n]
    T0: np.ndarray   # [P, n]

    @property
    def n_plants(self) -> int:
        return self.cfg.shape[0]

    def slice(self, sl) -> "Ensemble":
        return Ensemble(self.n_zones, self.cfg[sl].copy(), self.bnd[sl].copy(), self.pH0[sl].copy(),
                        self.Cl0[sl].copy(), self.T0[sl].copy())


def default_cfg_row() -> np.ndarray:
    """ReactorConfiguration() defaults (reactor.py:60-84)."""
    row = np.zeros(NCFG)
    row[_C["volume"]] = 1000.0
    row[_C["height"]] = 2.0
    row[_C["diameter"]] = 0.798
    row[_C["flow_rate"]] = 5.0
    row[_C["turbulent_intensity"]] = 0.15
    row[_C["recirculation_ratio"]] = 5.0
    row[_C["impeller_speed"]] = 60.0
    row[_C["impeller_diameter"]] = 0.3
    row[_C["power_number"]] = 5.0
    row[_C["initial_pH"]] = 7.0
    row[_C["alkalinity"]] = 100.0
    row[_C["total_carbonate"]] = 2.0
    row[_C["initial_chlorine"]] = 2.0
    row[_C["temperature"]] = 20.0
    row[_C["enable_thermal_stratification"]] = 1.0
    return row


def default_bnd_row() -> np.ndarray:
    """BoundaryConditions() defaults (reactor.py:168-186)."""
    row = np.zeros(NBND)
    row[_B["inlet_flow_rate"]] = 5.0
    row[_B["inlet_pH"]] = 7.5
    row[_B["inlet_chlorine"]] = 0.0
    row[_B["inlet_temperature"]] = 20.0
    row[_B["acid_flow_rate"]] = 0.0
    row[_B["acid_concentration"]] = 0.1
    row[_B["chlorine_flow_rate"]] = 0.0
    row[_B["chlorine_concentration"]] = 50.0
    row[_B["ambient_temperature"]] = 20.0
    row[_B["heat_loss_coefficient"]] = 0.0
    return row


def _uniform_state(cfg: np.ndarray, n: int):
    P = cfg.shape[0]
    pH0 = np.repeat(cfg[:, _C["initial_pH"]][:, None], n, axis=1)
    Cl0 = np.repeat(cfg[:, _C["initial_chlorine"]][:, None], n, axis=1)
    T0 = np.repeat(cfg[:, _C["temperature"]][:, None], n, axis=1)
    assert pH0.shape == (P, n)
    return pH0, Cl0, T0


def config1(n_zones: int = 5) -> Ensemble:
    """BASELINE config 1: the single default plant."""
    cfg = default_cfg_row()[None, :].copy()
    bnd = default_bnd_row()[None, :].copy()
    pH0, Cl0, T0 = _uniform_state(cfg, n_zones)
    return Ensemble(n_zones, cfg, bnd, pH0, Cl0, T0)


def _random_cfg_bnd(rng: np.random.Generator, P: int, t_lo: float, t_hi: float):
    cfg = np.repeat(default_cfg_row()[None, :], P, axis=0)
    bnd = np.repeat(default_bnd_row()[None, :], P, axis=0)
    cfg[:, _C["initial_pH"]] = rng.uniform(6.0, 9.0, P)
    cfg[:, _C["initial_chlorine"]] = rng.uniform(0.0, 5.0, P)
    cfg[:, _C["temperature"]] = rng.uniform(t_lo, t_hi, P)
    cfg[:, _C["flow_rate"]] = rng.uniform(1.0, 20.0, P)
    # diameter consistent with V = pi (D/2)^2 H * 1000 (ReactorConfiguration.validate, reactor.py:91-100)
    cfg[:, _C["diameter"]] = 2.0 * np.sqrt(cfg[:, _C["volume"]] / 1000.0 / (np.pi * cfg[:, _C["height"]]))
    cfg[:, _C["alkalinity"]] = rng.uniform(20.0, 300.0, P)
    cfg[:, _C["total_carbonate"]] = rng.uniform(0.5, 5.0, P)
    bnd[:, _B["inlet_flow_rate"]] = cfg[:, _C["flow_rate"]]
    bnd[:, _B["inlet_pH"]] = rng.uniform(6.5, 8.5, P)
    bnd[:, _B["inlet_chlorine"]] = rng.uniform(0.0, 2.0, P)
    return cfg, bnd


def config2(P: int = 4096, n_zones: int = 10, seed: int = SEED_CONFIG2) -> Ensemble:
    """BASELINE config 2 (and the physics inputs of config 5): noise-free random plants."""
    rng = np.random.default_rng(seed)
    cfg, bnd = _random_cfg_bnd(rng, P, 5.0, 35.0)
    bnd[:, _B["inlet_temperature"]] = cfg[:, _C["temperature"]] + rng.uniform(-5.0, 5.0, P)
    acid_on = rng.random(P) >= 0.5
    bnd[:, _B["acid_flow_rate"]] = np.where(acid_on, rng.uniform(0.0, 2.0, P), 0.0)
    cl_on = rng.random(P) >= 0.5
    bnd[:, _B["chlorine_flow_rate"]] = np.where(cl_on, rng.uniform(0.0, 1.0, P), 0.0)
    bnd[:, _B["heat_loss_coefficient"]] = 0.0
    pH0, Cl0, T0 = _uniform_state(cfg, n_zones)
    return Ensemble(n_zones, cfg, bnd, pH0, Cl0, T0)


def config3(P: int = 65536, n_zones: int = 20, seed: int = SEED_CONFIG3) -> Ensemble:
    """BASELINE config 3: temperature sweep 0-100 C with stratified / unstable profiles."""
    rng = np.random.default_rng(seed)
    cfg, bnd = _random_cfg_bnd(rng, P, 0.0, 40.0)
    H = cfg[:, _C["height"]]
    t_base = rng.uniform(0.5, 99.0, P)
    kind = rng.integers(0, 3, P)  # 0: uniform, 1: warm on top (stable above 4 C), 2: cold on top
    g = np.where(kind == 0, 0.0, np.where(kind == 1, 1.0, -1.0) * rng.uniform(0.0, 5.0, P))
    zc = (np.arange(n_zones)[None, :] + 0.5) * (H[:, None] / n_zones)
    T0 = np.clip(t_base[:, None] + g[:, None] * zc / H[:, None], 0.01, 99.9)
    bnd[:, _B["inlet_temperature"]] = rng.uniform(0.5, 99.0, P)
    cfg[:, _C["enable_thermal_stratification"]] = (rng.random(P) < 0.9).astype(np.float64)
    batch = rng.random(P) < 0.25
    bnd[:, _B["inlet_flow_rate"]] = np.where(batch, 0.0, cfg[:, _C["flow_rate"]])
    bnd[:, _B["heat_loss_coefficient"]] = np.where(rng.random(P) < 0.5, 0.0, 5.0)
    bnd[:, _B["ambient_temperature"]] = 20.0
    pH0, Cl0, _ = _uniform_state(cfg, n_zones)
    return Ensemble(n_zones, cfg, bnd, pH0, Cl0, T0)


def config4(P: int = 262144, seed: int = SEED_CONFIG4):
    """BASELINE config 4: calculate_pH stress inputs -> (alk, C_T, T, guess), each [P + 29]."""
    rng = np.random.default_rng(seed)
    alk = np.where(rng.random(P) < 0.2, 0.0, 10.0 ** rng.uniform(-3.0, 3.0, P))
    ct = np.where(rng.random(P) < 0.1, 0.0, 10.0 ** rng.uniform(-4.0, 1.5, P))
    temp = rng.uniform(0.0, 40.0, P)
    guess = rng.uniform(0.0, 14.0, P)
    grid = np.arange(0.0, 14.5, 0.5)  # deterministic grid on the default buffer (chemistry.py:546-550)
    alk = np.concatenate([alk, np.full(grid.size, 100.0)])
    ct = np.concatenate([ct, np.full(grid.size, 2.0)])
    temp = np.concatenate([temp, np.full(grid.size, 20.0)])
    guess = np.concatenate([guess, grid])
    return alk, ct, temp, guess


def config5(P: int = 1048576, n_zones: int = 10, seed: int = SEED_CONFIG5) -> Ensemble:
    """BASELINE config 5: the 1M-plant Monte-Carlo (physics inputs as config 2)."""
    return config2(P, n_zones, seed)
