"""Ensemble counterpart of the reference's main loop (src/wt_simulator/__main__.py:398-457).

Per plant: step -> read sensors -> controller commands -> zero-trust clamps -> boundary for the next step, all on
the device.  The clamps are the reference's ``validate_flow_rate`` / ``read_modbus_commands`` /
``apply_boundary_conditions`` (__main__.py:57-63, 227-271): NaN -> 0, acid in [0, 2], chlorine in [0, 1], inlet in
[0, 20] L/min and only applied when the command exceeds 0.1 L/min.  They run in one kernel of the CUDA library
(``wt_apply_commands``); scripted scenarios (``ScenarioTable``: time-indexed command rows resident on the device,
``wt_scenario_commands``) need no per-step host-to-device traffic at all and replay from a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .ensembles import BND_FIELDS, NBND

_B = {k: i for i, k in enumerate(BND_FIELDS)}


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _check_bnd(bnd_soa: torch.Tensor) -> int:
    if bnd_soa.ndim != 2 or bnd_soa.shape[0] != NBND or bnd_soa.dtype != torch.float64 or not bnd_soa.is_contiguous() \
            or not bnd_soa.is_cuda:
        raise ValueError(f"boundary batch must be a contiguous float64 CUDA tensor [{NBND}, P] (rows {BND_FIELDS})")
    return int(bnd_soa.shape[1])


def apply_boundary_conditions(bnd_soa: torch.Tensor, acid_rate: torch.Tensor, chlorine_rate: torch.Tensor,
                              inlet_rate: torch.Tensor) -> None:
    """__main__.py:227-271 on a boundary batch ``bnd_soa [10, P]`` (WT_BND_* rows), in place, one kernel: the
    commands pass ``read_modbus_commands``' clamps (__main__.py:239-246), then the defence-in-depth clamps of
    ``apply_boundary_conditions``; the inlet flow is only updated where the (clamped) command is > 0.1 L/min."""
    _lib.require_device()
    P = _check_bnd(bnd_soa)
    cmd = [torch.as_tensor(x, dtype=torch.float64, device=bnd_soa.device).reshape(P).contiguous()
           for x in (acid_rate, chlorine_rate, inlet_rate)]
    with torch.cuda.device(bnd_soa.device):
        rc = _lib.lib().wt_apply_commands(P, _p(cmd[0]), _p(cmd[1]), _p(cmd[2]), _p(bnd_soa),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "wt_apply_commands")


class ScenarioTable:
    """S scripts of K piecewise-constant actuator command triplets (acid, chlorine, inlet flow [L/min]), resident on
    the device.  ``times`` are the ascending breakpoints [s]; plant p follows script ``script_of_plant[p]``.  The
    scripted values are commands: they go through the same clamps as operator commands."""

    def __init__(self, times: Sequence[float], commands, script_of_plant=None, device=None):
        _lib.require_device()
        t = np.asarray(times, dtype=np.float64).reshape(-1)
        c = np.asarray(commands, dtype=np.float64)
        if c.ndim == 2:
            c = c[None]
        if c.ndim != 3 or c.shape[1:] != (t.size, 3):
            raise ValueError("commands must be [S, K, 3] (or [K, 3]) with K = len(times)")
        if np.any(np.diff(t) < 0):
            raise ValueError("breakpoints must be ascending")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.K, self.S, self.device = int(t.size), int(c.shape[0]), dev
        self._times = torch.from_numpy(t).to(dev)
        self._cmd = torch.from_numpy(np.ascontiguousarray(c)).to(dev)
        self._sid = None if script_of_plant is None else torch.as_tensor(script_of_plant, dtype=torch.int32).to(dev).contiguous()

    def apply(self, bnd_soa: torch.Tensor, t: Optional[float] = None, clock: Optional[torch.Tensor] = None) -> None:
        """Commands of the segment that contains the time (``clock[0]`` on the device, else ``t``) -> boundary."""
        P = _check_bnd(bnd_soa)
        if self._sid is not None and self._sid.numel() != P:
            raise ValueError("script_of_plant must have one entry per plant")
        if clock is None and t is None:
            raise ValueError("a time is required: t (host) or clock (device)")
        with torch.cuda.device(bnd_soa.device):
            rc = _lib.lib().wt_scenario_commands(P, self.K, self.S, _p(self._times), _p(self._cmd), _p(self._sid), _p(clock),
                                                 float(t if t is not None else 0.0), _p(bnd_soa),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "wt_scenario_commands")


class EnsembleOrchestrator:
    """step -> sensors -> controller / scenario -> clamps -> next boundary, for every plant of an ensemble.

    ``controller(readings, state, k) -> (acid_rate[P], chlorine_rate[P], inlet_rate[P])`` is any device-side function
    (a PID bank ...); it replaces the SCADA client behind the reference's Modbus holding registers.  ``scenario`` is a
    ``ScenarioTable`` evaluated at the time of the step that follows."""

    def __init__(self, ensemble, suite, boundary_soa: torch.Tensor, t0: float = 0.0):
        self.ens, self.suite, self.bnd = ensemble, suite, boundary_soa
        self.t0, self.k = float(t0), 0
        if suite is not None and not suite._initialized:
            suite.initialize(self.t0)

    def run(self, n_steps: int, dt: float, controller: Optional[Callable] = None,
            scenario: Optional[ScenarioTable] = None) -> Optional[Dict]:
        readings = None
        for _ in range(n_steps):
            state = self.ens.step(dt, self.bnd)                                        # __main__.py:403
            if self.suite is not None:
                readings = self.suite.read(state, self.t0 + self.k * dt)               # :408-410
            if controller is not None:
                acid, chlor, inlet = controller(readings, state, self.k)               # :422 (Modbus commands)
                apply_boundary_conditions(self.bnd, acid, chlor, inlet)                # :423
            self.k += 1
            if scenario is not None:
                scenario.apply(self.bnd, t=self.t0 + self.k * dt)
        return readings
