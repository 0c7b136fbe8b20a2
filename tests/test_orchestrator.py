"""Zero-trust command clamps of the reference's main loop (__main__.py:57-63, 227-271), batched."""
import math

import numpy as np
import torch

from ics_wt_physicsengine_b200.ensembles import BND_FIELDS, default_bnd_row
from ics_wt_physicsengine_b200.orchestrator import apply_boundary_conditions, validate_flow_rate


def ref_validate_flow_rate(value, max_value=20.0):   # restated from __main__.py:57-63
    if not isinstance(value, (int, float)):
        return 0.0
    if value != value:
        return 0.0
    return max(0.0, min(float(value), max_value))


def test_validate_flow_rate_matches_reference_semantics():
    vals = [-5.0, -0.0, 0.0, 0.05, 0.1, 0.1000001, 1.0, 2.0, 2.5, 19.9, 20.0, 25.0, float("nan"), float("inf"), -float("inf")]
    for mx in (1.0, 2.0, 20.0):
        got = validate_flow_rate(torch.tensor(vals, dtype=torch.float64), mx).numpy()
        want = np.array([ref_validate_flow_rate(v, mx) for v in vals])
        assert np.array_equal(got, want)


def test_apply_boundary_conditions_matches_reference_loop():
    rng = np.random.default_rng(0)
    P = 500
    acid = rng.uniform(-1, 4, P); chlor = rng.uniform(-1, 2, P); inlet = rng.uniform(-1, 30, P)
    acid[::17] = np.nan; inlet[::13] = np.nan; inlet[5] = 0.1; inlet[6] = 0.05
    bnd = torch.from_numpy(np.repeat(default_bnd_row()[:, None], P, axis=1).copy())
    apply_boundary_conditions(bnd, torch.from_numpy(acid), torch.from_numpy(chlor), torch.from_numpy(inlet))
    for p in range(P):
        a = ref_validate_flow_rate(ref_validate_flow_rate(float(acid[p]), 2.0), 2.0)
        c = ref_validate_flow_rate(ref_validate_flow_rate(float(chlor[p]), 1.0), 1.0)
        i_cmd = ref_validate_flow_rate(float(inlet[p]), 20.0)
        i = ref_validate_flow_rate(i_cmd, 20.0) if i_cmd > 0.1 else 5.0
        assert bnd[BND_FIELDS.index("acid_flow_rate"), p] == a
        assert bnd[BND_FIELDS.index("chlorine_flow_rate"), p] == c
        assert bnd[BND_FIELDS.index("inlet_flow_rate"), p] == i
    assert not math.isnan(float(bnd.sum()))
