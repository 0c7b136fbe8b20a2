"""Ensemble counterpart of the reference's main loop (src/wt_simulator/__main__.py:398-457).

Per plant: step -> read sensors -> controller commands -> zero-trust clamps -> boundary for the
next step, all on the device (no per-step host round trip).  The clamps are the reference's
``validate_flow_rate`` / ``read_modbus_commands`` / ``apply_boundary_conditions``
(__main__.py:57-63, 227-271): NaN -> 0, acid in [0, 2], chlorine in [0, 1], inlet in [0, 20] L/min
and only applied when the command exceeds 0.1 L/min.  These are elementwise tensor operations
(plumbing), so they run wherever the command tensors live (CUDA in production, CPU in the tests).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .ensembles import BND_FIELDS

_B = {k: i for i, k in enumerate(BND_FIELDS)}


def validate_flow_rate(value: torch.Tensor, max_value: float = 20.0) -> torch.Tensor:
    """__main__.py:57-63, batched: NaN -> 0, then clamp to [0, max_value]."""
    v = torch.where(torch.isnan(value), torch.zeros_like(value), value)
    return torch.clamp(v, min=0.0, max=max_value)


def apply_boundary_conditions(bnd_soa: torch.Tensor, acid_rate: torch.Tensor, chlorine_rate: torch.Tensor,
                              inlet_rate: torch.Tensor) -> None:
    """__main__.py:255-271 on a boundary batch ``bnd_soa [10, P]`` (WT_BND_* rows), in place.

    The commands first pass ``read_modbus_commands``' clamps (__main__.py:239-246), then the
    defence-in-depth clamps of ``apply_boundary_conditions``; the inlet flow is only updated where
    the (clamped) command is > 0.1 L/min."""
    acid = validate_flow_rate(validate_flow_rate(acid_rate, 2.0), 2.0)
    chlor = validate_flow_rate(validate_flow_rate(chlorine_rate, 1.0), 1.0)
    inlet = validate_flow_rate(inlet_rate, 20.0)
    bnd_soa[_B["acid_flow_rate"]].copy_(acid)
    bnd_soa[_B["chlorine_flow_rate"]].copy_(chlor)
    cur = bnd_soa[_B["inlet_flow_rate"]]
    cur.copy_(torch.where(inlet > 0.1, validate_flow_rate(inlet, 20.0), cur))


class EnsembleOrchestrator:
    """step -> sensors -> controller -> clamps -> next boundary, for every plant of an ensemble.

    ``controller(readings, state, k) -> (acid_rate[P], chlorine_rate[P], inlet_rate[P])`` is any
    device-side function (a PID bank, a scripted scenario ...); it replaces the SCADA client behind
    the reference's Modbus holding registers."""

    def __init__(self, ensemble, suite, boundary_soa: torch.Tensor, t0: float = 0.0):
        self.ens, self.suite, self.bnd = ensemble, suite, boundary_soa
        self.t0, self.k = float(t0), 0
        if suite is not None and not suite._initialized:
            suite.initialize(self.t0)

    def run(self, n_steps: int, dt: float, controller: Optional[Callable] = None) -> Dict:
        readings = None
        for _ in range(n_steps):
            state = self.ens.step(dt, self.bnd)                                        # __main__.py:403
            if self.suite is not None:
                readings = self.suite.read(state, self.t0 + self.k * dt)               # :408-410
            if controller is not None:
                acid, chlor, inlet = controller(readings, state, self.k)               # :422 (Modbus commands)
                apply_boundary_conditions(self.bnd, acid, chlor, inlet)                # :423
            self.k += 1
        return readings
