"""The diagnostics oracle (oracle/wt_diag_oracle.py) against outputs of the unmodified reference
(tests/golden/diagnostics_48.npz, produced by oracle/gen_golden_diag.py)."""
import os

import numpy as np

from ics_wt_physicsengine_b200.params import derive_params
from oracle import wt_diag_oracle as wd


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "diagnostics_48.npz"))
    assert tuple(g["fields"]) == wd.FIELDS
    for n in np.unique(g["n_zones"]):
        m = g["n_zones"] == n
        cfg = g["cfg_h"][m][:, :-20]
        H = g["cfg_h"][m][:, -20:][:, :n]
        y = g["y"][m][:, :3 * n]
        yield int(n), derive_params(cfg, int(n)), y, H, g["out"][m], g["n2"][m][:, :n - 1]


def test_oracle_matches_the_reference(golden_dir):
    f = {k: i for i, k in enumerate(wd.FIELDS)}
    seen_none = seen_depth = 0
    for n, par, y, H, want, want_n2 in _cases(golden_dir):
        out, n2, bad = wd.diagnostics(par, y, n, H)
        assert not bad.any()
        for k, i in f.items():
            a, b = out[:, i], want[:, i]
            assert np.array_equal(np.isnan(a), np.isnan(b)), k
            ok = ~np.isnan(b)
            # floating-point reductions: numpy's pairwise summation order is not restated, hence 1e-12
            assert np.allclose(a[ok], b[ok], rtol=1e-12, atol=1e-300), (k, n)
        assert np.allclose(n2, want_n2, rtol=1e-12, atol=1e-300)
        assert np.array_equal(out[:, f["temperature_gradient_location"]], want[:, f["temperature_gradient_location"]])
        seen_none += int(np.isnan(want[:, f["thermocline_depth"]]).sum())
        seen_depth += int((~np.isnan(want[:, f["thermocline_depth"]])).sum())
    assert seen_none > 0 and seen_depth > 0  # both branches of identify_thermocline are pinned


def test_temperature_out_of_range_is_flagged():
    par = derive_params(np.load(os.path.join(os.path.dirname(__file__), "golden", "diagnostics_48.npz"))["cfg_h"][:2, :-20], 5)
    y = np.tile(np.concatenate([np.full(5, 7.0), np.full(5, 2.0), np.full(5, 20.0)]), (2, 1))
    y[1, 10] = 120.0
    _, _, bad = wd.diagnostics(par, y, 5)
    assert list(bad) == [False, True]
