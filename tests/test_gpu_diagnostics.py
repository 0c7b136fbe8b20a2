"""K5, the batched diagnostics kernel (wt_diagnostics), against the reference's own outputs (golden) and the
numpy oracle on large random ensembles.  Floating-point reductions: 1e-12 relative (summation order)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from ics_wt_physicsengine_b200 import IntegratedCSTR, PlantEnsemble, ReactorConfiguration, ensembles as ens  # noqa: E402
from ics_wt_physicsengine_b200.reactor import DIAG_FIELDS  # noqa: E402
from oracle import wt_diag_oracle as wd  # noqa: E402

RTOL = 1e-12


def _compare(got, want, n2_got, n2_want, zh):
    """Relative 1e-12 on well-conditioned fields.  Spread-type fields (std, range, gradients, CV, N^2, the
    energy relative to 20 C, the H+ - OH- balance) are differences of nearly equal numbers on well-mixed
    plants: there the error is bounded against the magnitude of the operands, not of the result."""
    assert DIAG_FIELDS == wd.FIELDS
    F = {k: i for i, k in enumerate(wd.FIELDS)}
    for i, k in enumerate(wd.FIELDS):
        a, b = got[k].cpu().numpy(), want[:, i]
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        ok = ~np.isnan(b)
        atol = np.full(b.shape, 1e-290)
        var = k.split("_")[0]
        if var in ("pH", "chlorine", "temperature") and k not in ("chlorine_cv", "chlorine_segregation"):
            scale = np.maximum(np.abs(want[:, F[f"{var}_max_value"]]), np.abs(want[:, F[f"{var}_min_value"]]))
            if k.endswith(("std_value", "range")):
                atol = 1e-13 * scale
            elif k.endswith(("max_gradient", "mean_gradient")):
                atol = 1e-13 * scale / zh
        elif k in ("chlorine_cv", "chlorine_segregation"):
            atol = np.full(b.shape, 1e-13)
        elif k == "charge_balance_mol":
            atol = RTOL * (np.abs(want[:, F["total_H_mol"]]) + np.abs(want[:, F["total_OH_mol"]]))
        elif k == "thermal_energy_kJ":
            atol = 1e-13 * np.abs(want[:, F["temperature_max_value"]]) * 998.2 * 4.184 * 1e3
        elif k.startswith("brunt_vaisala"):
            atol = 1e-12 * 9.81 / zh
        if k.endswith("gradient_location"):
            assert (a[ok] == b[ok]).mean() > 0.999, k   # an argmax may differ on ties of rounded gradients
        else:
            err = np.abs(a[ok] - b[ok])
            assert np.all(err <= RTOL * np.abs(b[ok]) + atol[ok]), (k, float(err.max()))
    assert np.all(np.abs(n2_got.cpu().numpy().T - n2_want) <= RTOL * np.abs(n2_want) + (1e-12 * 9.81 / zh)[:, None])


def test_reference_golden_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "diagnostics_48.npz"))
    for n in np.unique(g["n_zones"]):
        n = int(n)
        m = g["n_zones"] == n
        cfg = g["cfg_h"][m][:, :-20]
        y = g["y"][m][:, :3 * n]
        eng = PlantEnsemble(cfg, n_zones=n)
        eng.set_state(y[:, :n], y[:, n:2 * n], y[:, 2 * n:])
        d = eng.diagnostics(with_n2=True)
        assert int(d["bad"].sum()) == 0
        _compare(d, g["out"][m], d["n2"], g["n2"][m][:, :n - 1], eng.par_host[:, 7])


@pytest.mark.parametrize("cfg,P,n", [("config2", 20001, 10), ("config3", 65536, 20), ("config2", 777, 2), ("config2", 333, 32)])
def test_against_the_oracle_on_random_ensembles(cfg, P, n):
    e = getattr(ens, cfg)(P, n, seed=31 + n)
    eng = PlantEnsemble(e, max_attempts=64)
    for _ in range(2):
        eng.step(1.0, e.bnd)      # a stepped state: derived H comes from the kernel's own exp10
    d = eng.diagnostics(with_n2=True)
    y = eng.state_numpy()
    want, n2, bad = wd.diagnostics(eng.par_host, y, n, H=eng._derived[0].cpu().numpy().T)
    assert np.array_equal(d["bad"].cpu().numpy() != 0, bad)
    _compare(d, want, d["n2"], n2, eng.par_host[:, 7])


def test_single_plant_facade_matches_reference_keys():
    r = IntegratedCSTR(ReactorConfiguration())
    out = r.validate_conservation()
    assert set(out) == {"total_chlorine_mg", "total_H_mol", "total_OH_mol", "charge_balance_mol", "thermal_energy_kJ",
                        "zones", "timestamp"}
    # default plant: 5 zones x 200 L x 2 mg/L, pH 7 at 20 C (reactor.py:52-110)
    assert out["total_chlorine_mg"] == pytest.approx(2000.0, rel=1e-14)
    assert out["total_H_mol"] == pytest.approx(1e-7, rel=1e-12)  # sum(H) * zone_volume / 1000, as the reference writes it
    assert out["thermal_energy_kJ"] == 0.0 and out["zones"] == 5
