"""Host-side mirror of ``wt_simulator.core.reactor`` for the B200 engine.

Same names, fields, argument meaning and error behaviour as the reference
(src/wt_simulator/core/reactor.py): ``ReactorConfiguration`` (:52-110), ``ReactorState``
(:113-147), ``BoundaryConditions`` (:150-186), ``IntegratedCSTR`` (:189-611) -- plus the
batched ``PlantEnsemble`` the engine exists for.  Everything numerical happens in the CUDA
library (csrc/libwt_b200.so) through the C ABI of include/wt_b200.h; this module only lays
out device memory (torch tensors) and mirrors the reference's interface.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field, fields
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .ensembles import BND_FIELDS, CFG_FIELDS, NBND, NCFG, Ensemble
from .params import NPAR, derive_params, validate_cfg

# rows of wt_diagnostics (WT_DG_* of include/wt_b200.h)
DIAG_FIELDS = (
    "total_chlorine_mg", "total_H_mol", "total_OH_mol", "charge_balance_mol", "thermal_energy_kJ",
    "chlorine_cv", "chlorine_segregation",
    *[f"{v}_{k}" for v in ("pH", "chlorine", "temperature")
      for k in ("mean_value", "std_value", "max_value", "min_value", "range", "max_gradient", "mean_gradient",
                "gradient_location")],
    "thermocline_depth", "brunt_vaisala_max", "brunt_vaisala_min",
)

logger = logging.getLogger(__name__)


# ----------------------------------------------------------------------------------------
# dataclasses with the reference's fields and defaults
# ----------------------------------------------------------------------------------------
@dataclass
class ReactorConfiguration:
    """reactor.py:52-110 (same fields, defaults and validate())."""

    volume: float = 1000.0
    height: float = 2.0
    diameter: float = 0.798
    n_zones: int = 5
    flow_rate: float = 5.0
    turbulent_intensity: float = 0.15
    recirculation_ratio: float = 5.0
    impeller_speed: float = 60.0
    impeller_diameter: float = 0.3
    power_number: float = 5.0
    initial_pH: float = 7.0
    alkalinity: float = 100.0
    total_carbonate: float = 2.0
    initial_chlorine: float = 2.0
    temperature: float = 20.0
    enable_thermal_stratification: bool = True
    inlet_pH: float = 7.5
    inlet_chlorine: float = 0.0
    inlet_temperature: float = 20.0

    def validate(self) -> None:
        validate_cfg(self.as_row()[None, :], self.n_zones)

    def as_row(self) -> np.ndarray:
        return np.array([float(getattr(self, k)) for k in CFG_FIELDS], dtype=np.float64)


@dataclass
class BoundaryConditions:
    """reactor.py:150-186 (same fields and defaults)."""

    inlet_flow_rate: float = 5.0
    inlet_pH: float = 7.5
    inlet_chlorine: float = 0.0
    inlet_temperature: float = 20.0
    acid_flow_rate: float = 0.0
    acid_concentration: float = 0.1
    chlorine_flow_rate: float = 0.0
    chlorine_concentration: float = 50.0
    ambient_temperature: float = 20.0
    heat_loss_coefficient: float = 0.0

    def as_row(self) -> np.ndarray:
        return np.array([float(getattr(self, k)) for k in BND_FIELDS], dtype=np.float64)


@dataclass
class ReactorState:
    """reactor.py:113-147: state of ONE plant (numpy arrays of length n_zones)."""

    time: float = 0.0
    pH: np.ndarray = field(default_factory=lambda: np.full(5, 7.0))
    chlorine: np.ndarray = field(default_factory=lambda: np.full(5, 2.0))
    temperature: np.ndarray = field(default_factory=lambda: np.full(5, 20.0))
    flow_rate: float = 5.0
    H_concentration: np.ndarray = field(init=False)
    density: np.ndarray = field(init=False)
    chlorine_decay_rate: np.ndarray = field(init=False)

    def __post_init__(self):
        self.update_derived()

    def update_derived(self):
        self.H_concentration = 10 ** (-self.pH)
        if not hasattr(self, "density"):
            self.density = np.full_like(self.pH, 998.2)
        if not hasattr(self, "chlorine_decay_rate"):
            self.chlorine_decay_rate = np.full_like(self.pH, 0.0001)


class EnsembleState:
    """State of P plants, resident in HBM.

    ``pH``, ``chlorine``, ``temperature`` (and the derived ``H_concentration``, ``density``,
    ``chlorine_decay_rate``) are ``[P, n_zones]`` views of the zone-major device storage
    ``[n_zones, P]``; ``time`` and ``flow_rate`` are ``[P]``.  ``state[p]`` gives a host
    ``ReactorState`` of one plant that the reference's sensors can read unchanged.
    """

    def __init__(self, y: torch.Tensor, derived: torch.Tensor, time: torch.Tensor, flow: torch.Tensor):
        self._y, self._derived, self.time, self.flow_rate = y, derived, time, flow

    @property
    def pH(self) -> torch.Tensor:
        return self._y[0].t()

    @property
    def chlorine(self) -> torch.Tensor:
        return self._y[1].t()

    @property
    def temperature(self) -> torch.Tensor:
        return self._y[2].t()

    @property
    def H_concentration(self) -> torch.Tensor:
        return self._derived[0].t()

    @property
    def density(self) -> torch.Tensor:
        return self._derived[1].t()

    @property
    def chlorine_decay_rate(self) -> torch.Tensor:
        return self._derived[2].t()

    def __len__(self) -> int:
        return self._y.shape[2]

    def __getitem__(self, p: int) -> ReactorState:
        y = self._y[:, :, p].cpu().numpy()
        d = self._derived[:, :, p].cpu().numpy()
        s = ReactorState(time=float(self.time[p]), pH=y[0].copy(), chlorine=y[1].copy(),
                         temperature=y[2].copy(), flow_rate=float(self.flow_rate[p]))
        s.H_concentration, s.density, s.chlorine_decay_rate = d[0].copy(), d[1].copy(), d[2].copy()
        return s


BoundaryLike = Union[BoundaryConditions, np.ndarray, torch.Tensor, Sequence[BoundaryConditions]]


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class PlantEnsemble:
    """P independent multi-zone CSTR plants advanced together on one B200.

    ``step(dt, boundary)`` has the call shape of ``IntegratedCSTR.step`` (reactor.py:450-509)
    applied to every plant: one scipy-Radau solve over ``[t, t+dt]`` per plant, derived state,
    bound clipping.  What the reference signals by raising / logging becomes bits of the
    per-plant ``status`` word (include/wt_b200.h, WT_ST_*); plants whose step raised in the
    reference (WT_ST_T_RANGE) HALT, as the reference's main loop stops on a physics exception
    (__main__.py:404-406).

    max_attempts: budget of collocation solves per plant-step (engine policy, see DESIGN.md
    "straggler policy"); 0 means no budget -- the reference's behaviour -- up to a hard stop at 2,000,000
    collocation solves of one step (WT_HARD_MAX_ATTEMPTS), after which the plant gets WT_ST_WORK_LIMIT like any
    other budget overrun.
    """

    DEFAULT_MAX_ATTEMPTS = 64

    def __init__(self, cfg: Union[np.ndarray, Sequence[ReactorConfiguration], Ensemble], n_zones: Optional[int] = None,
                 device: Union[str, torch.device, None] = None, max_attempts: int = DEFAULT_MAX_ATTEMPTS,
                 validate: bool = True, sort_every: int = 0, catch_up_attempts: int = 0, catch_up_floor_div: int = 0):
        _lib.require_device()
        init = None
        if isinstance(cfg, Ensemble):
            init, n_zones, cfg = cfg, cfg.n_zones, cfg.cfg
        elif not isinstance(cfg, np.ndarray):
            cfgs = list(cfg)
            nz = {c.n_zones for c in cfgs}
            if len(nz) != 1:
                raise ValueError("all plants of an ensemble must have the same n_zones")
            n_zones = nz.pop()
            cfg = np.stack([c.as_row() for c in cfgs])
        if n_zones is None:
            raise ValueError("n_zones is required with a configuration matrix")
        cfg = np.ascontiguousarray(cfg, dtype=np.float64).reshape(-1, NCFG)
        if validate:
            validate_cfg(cfg, n_zones)
        self.n_zones = int(n_zones)
        self.n_plants = int(cfg.shape[0])
        self.cfg = cfg
        self.max_attempts = int(max_attempts)
        # > 0: plants that exhaust max_attempts are not halted but DEFERRED: collect_deferred / catch_up / rejoin_deferred
        # continue them with this larger budget (partition.PipelinedShard runs the three per block of steps)
        self.catch_up_attempts = int(catch_up_attempts)
        # > 0: the catch-up runs in FLOOR MODE (engine policy, include/wt_b200.h wt_catch_up): step sizes >= dt / floor_div
        # with forced acceptance at the floor, WT_ST_DEGRADED on the plant-steps that needed it.  Bounded cost for the
        # plants that sit on the 8 C density discontinuity, where the reference's adaptive control needs 1e5 .. 1e7
        # evaluations per step.
        self.catch_up_floor_div = int(catch_up_floor_div)
        if self.catch_up_floor_div < 0:
            raise ValueError("catch_up_floor_div must be >= 0")
        # scheduling only (results do not depend on it): every `sort_every` launches the plants are
        # re-ordered by the work of their last step, most expensive first (0 = natural order)
        self.sort_every = int(sort_every)
        self._launches = 0
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        P, n = self.n_plants, self.n_zones
        par = derive_params(cfg, n)
        self.par_host = par
        with torch.cuda.device(self.device):
            self._par = torch.from_numpy(np.ascontiguousarray(par.T)).to(self.device)  # [NPAR, P]
            self._y = torch.empty((3, n, P), dtype=torch.float64, device=self.device)
            self._derived = torch.empty((3, n, P), dtype=torch.float64, device=self.device)
            self._time = torch.zeros(P, dtype=torch.float64, device=self.device)
            self._flow = torch.from_numpy(cfg[:, CFG_FIELDS.index("flow_rate")].copy()).to(self.device)
            self._status = torch.zeros(P, dtype=torch.int32, device=self.device)
            self._counters = torch.zeros((_lib.NCNT, P), dtype=torch.int32, device=self.device)
            self._bnd_bcast = torch.zeros(NBND, dtype=torch.float64, device=self.device)
            self._bnd_batch = None
            # cost order (scheduling only): a static buffer so that captured (CUDA graph) steps keep their pointers
            self._order = torch.arange(P, dtype=torch.int32, device=self.device) if self.sort_every > 0 else None
            self._cost = torch.zeros(P, dtype=torch.int32, device=self.device) if self.sort_every > 0 else None
            self._bins = torch.zeros(1024, dtype=torch.int32, device=self.device) if self.sort_every > 0 else None
            if self.catch_up_attempts > 0:
                cap = max(1024, P // 32)
                self._defer_cap = cap
                self._defer_list = torch.zeros(cap, dtype=torch.int32, device=self.device)
                self._defer_count = torch.zeros(1, dtype=torch.int32, device=self.device)
                self._t_stop = torch.zeros(1, dtype=torch.float64, device=self.device)
                self._ws_defer = torch.empty(int(_lib.lib().wt_step_workspace_bytes(cap, n)), dtype=torch.uint8, device=self.device)
            # device workspace of the step (work queue + hand-off rows between its two launches), reused by every step
            self._ws = torch.empty(int(_lib.lib().wt_step_workspace_bytes(P, n)), dtype=torch.uint8, device=self.device)
        self.state = EnsembleState(self._y, self._derived, self._time, self._flow)
        if init is not None:
            self.set_state(init.pH0, init.Cl0, init.T0)
        else:
            c = {k: cfg[:, i] for i, k in enumerate(CFG_FIELDS)}
            ones = np.ones((P, n))
            self.set_state(c["initial_pH"][:, None] * ones, c["initial_chlorine"][:, None] * ones,
                           c["temperature"][:, None] * ones)

    # ---- state access --------------------------------------------------------------------
    def set_state(self, pH, chlorine, temperature, time=None) -> None:
        """Overwrite the primary state ([P, n_zones] each), as assigning reactor.state.* does."""
        P, n = self.n_plants, self.n_zones
        for v, a in enumerate((pH, chlorine, temperature)):
            a = torch.as_tensor(np.array(a, dtype=np.float64) if not torch.is_tensor(a) else a,
                                dtype=torch.float64).reshape(P, n)
            self._y[v].copy_(a.t().to(self.device))
        if time is not None:
            self._time.copy_(torch.as_tensor(time, dtype=torch.float64).reshape(P).to(self.device))
        # placeholders of ReactorState.update_derived (reactor.py:136-147)
        self._derived[0].copy_(torch.pow(10.0, -self._y[0]))
        self._derived[1].fill_(998.2)
        self._derived[2].fill_(0.0001)

    def state_numpy(self) -> np.ndarray:
        """[P, 3*n_zones] species-major copy of the primary state (the reference's ODE vector)."""
        return self._y.permute(2, 0, 1).reshape(self.n_plants, 3 * self.n_zones).cpu().numpy()

    @property
    def status(self) -> torch.Tensor:
        return self._status

    @property
    def counters(self) -> torch.Tensor:
        """[WT_NCNT, P] accumulated solver path counters (nfev, njev, nlu, ...)."""
        return self._counters

    def reset_status(self) -> None:
        self._status.zero_()

    def reset_counters(self) -> None:
        self._counters.zero_()

    # ---- boundary --------------------------------------------------------------------------
    def _boundary(self, boundary: BoundaryLike):
        """-> (device tensor, stride).  One BoundaryConditions broadcasts to all plants."""
        if isinstance(boundary, BoundaryConditions):
            row = torch.from_numpy(boundary.as_row())
            self._bnd_bcast.copy_(row, non_blocking=True)
            return self._bnd_bcast, 0
        if torch.is_tensor(boundary):
            b = boundary
            if b.shape == (NBND, self.n_plants) and b.device == self.device and b.dtype == torch.float64 \
                    and b.is_contiguous():
                return b, self.n_plants  # already SoA on the device: borrowed, no copy
            b = b.to(torch.float64)
        elif isinstance(boundary, np.ndarray):
            b = torch.from_numpy(np.asarray(boundary, dtype=np.float64))
        else:
            b = torch.from_numpy(np.stack([x.as_row() for x in boundary]))
        if b.ndim == 1:
            self._bnd_bcast.copy_(b.reshape(NBND))
            return self._bnd_bcast, 0
        if b.shape != (self.n_plants, NBND):
            raise ValueError(f"boundary batch must be [P={self.n_plants}, {NBND}] (fields {BND_FIELDS})")
        if self._bnd_batch is None:
            self._bnd_batch = torch.empty((NBND, self.n_plants), dtype=torch.float64, device=self.device)
        self._bnd_batch.copy_(b.t())
        return self._bnd_batch, self.n_plants

    # ---- the hot path ----------------------------------------------------------------------
    def step(self, dt: float, boundary: BoundaryLike) -> EnsembleState:
        """Advance every plant by ``dt`` seconds (IntegratedCSTR.step, reactor.py:450-509)."""
        return self.advance(1, dt, boundary)

    def advance(self, n_steps: int, dt: float, boundary: BoundaryLike) -> EnsembleState:
        """``n_steps`` consecutive ``step(dt, boundary)`` calls fused into one kernel launch."""
        bnd, stride = self._boundary(boundary)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            if self.sort_every > 0 and self._launches > 0 and self._launches % self.sort_every == 0:
                _lib.check(_lib.lib().wt_cost_order(self.n_plants, _ptr(self._cost), _ptr(self._order), _ptr(self._bins),
                                                    C.c_void_p(stream)), "wt_cost_order")
            rc = _lib.lib().wt_advance(self.n_plants, self.n_zones, int(n_steps), float(dt), _ptr(self._par),
                                       _ptr(bnd), stride, _ptr(self._time), _ptr(self._y), _ptr(self._flow),
                                       _ptr(self._derived), _ptr(self._status), _ptr(self._counters),
                                       self.max_attempts, _ptr(self._order), _ptr(self._cost), _ptr(self._ws),
                                       C.c_void_p(stream))
        _lib.check(rc, "wt_advance")
        self._launches += 1
        return self.state

    # ---- deferral of budget-exhausted plants (include/wt_b200.h: wt_defer_collect / wt_catch_up / wt_defer_rejoin) ----
    def collect_deferred(self, t_stop_inc: float = 0.0) -> None:
        """Plants that ran out of ``max_attempts`` -> the deferred list (they keep their state, ordinary steps pass
        over them until ``rejoin_deferred``).  ``t_stop_inc`` is added to the stop time of the catch-up first."""
        with torch.cuda.device(self.device):
            rc = _lib.lib().wt_defer_collect(self.n_plants, _ptr(self._status), _ptr(self._defer_list), _ptr(self._defer_count),
                                             self._defer_cap, _ptr(self._t_stop), float(t_stop_inc),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "wt_defer_collect")

    def catch_up(self, n_steps: int, dt: float, boundary: BoundaryLike) -> None:
        """Up to ``n_steps`` x step(dt) for the deferred plants with the budget ``catch_up_attempts``, each only until its
        time reaches the device scalar ``_t_stop``.  Launch it on a side stream: it only touches the listed plants."""
        bnd, stride = self._boundary(boundary)
        with torch.cuda.device(self.device):
            rc = _lib.lib().wt_catch_up(self._defer_cap, self.n_plants, self.n_zones, int(n_steps), float(dt), _ptr(self._par),
                                        _ptr(bnd), stride, _ptr(self._time), _ptr(self._y), _ptr(self._flow), _ptr(self._derived),
                                        _ptr(self._status), _ptr(self._counters), self.catch_up_attempts, self.catch_up_floor_div,
                                        _ptr(self._defer_list),
                                        _ptr(self._defer_count), _ptr(self._t_stop), _ptr(self._ws_defer),
                                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "wt_catch_up")

    def rejoin_deferred(self) -> None:
        """After the catch-up (stream order): the listed plants take part in the ordinary steps again."""
        with torch.cuda.device(self.device):
            rc = _lib.lib().wt_defer_rejoin(_ptr(self._status), _ptr(self._defer_list), _ptr(self._defer_count), self._defer_cap,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "wt_defer_rejoin")

    def derivatives(self, boundary: BoundaryLike, y: Optional[torch.Tensor] = None):
        """Batched IntegratedCSTR.derivatives (reactor.py:272-448) -> (dy [3,n,P], bad [P])."""
        bnd, stride = self._boundary(boundary)
        y = self._y if y is None else y
        dy = torch.empty_like(self._y)
        bad = torch.zeros(self.n_plants, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.lib().wt_derivatives(self.n_plants, self.n_zones, _ptr(self._par), _ptr(bnd), stride,
                                           _ptr(y), _ptr(dy), _ptr(bad), C.c_void_p(stream))
        _lib.check(rc, "wt_derivatives")
        return dy, bad

    def diagnostics(self, with_n2: bool = False):
        """Per-plant diagnostics of the current state in one kernel: validate_conservation
        (reactor.py:570-611), calculate_mixing_quality of chlorine (transport.py:338-384),
        calculate_spatial_gradients of pH / chlorine / temperature (spatial.py:440-477), identify_thermocline
        (NaN where the reference returns None, spatial.py:353-379) and the Brunt-Vaisala N^2 extremes
        (spatial.py:322-351).  Returns {field: tensor[P]} (+ "n2": [n-1, P] and "bad": [P] with ``with_n2``)."""
        P, n = self.n_plants, self.n_zones
        out = torch.empty((len(DIAG_FIELDS), P), dtype=torch.float64, device=self.device)
        n2 = torch.empty((n - 1, P), dtype=torch.float64, device=self.device) if with_n2 else None
        bad = torch.zeros(P, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.lib().wt_diagnostics(P, n, _ptr(self._par), _ptr(self._y), _ptr(self._derived[0]), _ptr(out),
                                           _ptr(n2), _ptr(bad), C.c_void_p(stream))
        _lib.check(rc, "wt_diagnostics")
        res = {k: out[i] for i, k in enumerate(DIAG_FIELDS)}
        res["bad"] = bad
        if with_n2:
            res["n2"] = n2
        return res


class IntegratedCSTR:
    """Drop-in for ``wt_simulator.core.reactor.IntegratedCSTR`` (reactor.py:189-611), one plant.

    ``step(dt, boundary)`` mutates and returns ``self.state`` (a host ``ReactorState``) exactly as
    the reference does; the computation runs on the GPU through a 1-plant ``PlantEnsemble`` with
    an unlimited attempt budget (reference semantics).  Raises ``ValueError`` where the
    reference does (temperature outside [0, 100] C inside the solve).
    """

    def __init__(self, config: ReactorConfiguration, device=None):
        config.validate()
        self.config = config
        self._ens = PlantEnsemble([config], device=device, max_attempts=0)
        self.state = ReactorState(
            pH=np.full(config.n_zones, config.initial_pH),
            chlorine=np.full(config.n_zones, config.initial_chlorine),
            temperature=np.full(config.n_zones, config.temperature),
            flow_rate=config.flow_rate,
        )

    def step(self, dt: float, boundary: BoundaryConditions) -> ReactorState:
        e, s = self._ens, self.state
        e.set_state(s.pH[None, :], s.chlorine[None, :], s.temperature[None, :], time=[s.time])
        e.reset_status()
        e.step(dt, boundary)
        st = int(e.status[0])
        if st & _lib.ST_T_RANGE:
            raise ValueError("Temperature outside liquid water range [0.0, 100.0]°C inside the ODE solve "
                             "(thermodynamics.py:146-157)")
        if st & _lib.ST_WORK_LIMIT:
            # max_attempts = 0 means "no budget", but the kernel still stops after WT_HARD_MAX_ATTEMPTS (2,000,000)
            # collocation solves of ONE step (the reference would grind on: such steps take it hours).  The state was
            # left untouched; say so instead of returning an un-advanced state silently.
            raise RuntimeError("step(dt) gave up after 2,000,000 collocation solves (plant on the 8 C density "
                               "discontinuity, see DESIGN.md section 7); the state was not advanced")
        if st & _lib.ST_SOLVER_FAILED:
            logger.warning("ODE solver failed: Required step size is less than spacing between numbers.")
        new = e.state[0]
        s.pH, s.chlorine, s.temperature = new.pH, new.chlorine, new.temperature
        s.time, s.flow_rate = new.time, new.flow_rate
        s.H_concentration, s.density, s.chlorine_decay_rate = new.H_concentration, new.density, new.chlorine_decay_rate
        if st & _lib.ST_T_RANGE_DERIVED:
            raise ValueError("Temperature outside liquid water range [0.0, 100.0]°C in _update_derived_state")
        if st & _lib.ST_CLIP_PH:
            logger.error(f"pH out of bounds: clipped to {s.pH}")
        if st & _lib.ST_CLIP_CL:
            logger.warning(f"Negative chlorine detected: clipped to {s.chlorine}")
        if st & _lib.ST_CLIP_T:
            logger.error(f"Temperature out of bounds: clipped to {s.temperature}")
        return s

    def derivatives(self, t: float, y: np.ndarray, boundary: BoundaryConditions) -> np.ndarray:
        n = self.config.n_zones
        yy = torch.from_numpy(np.asarray(y, dtype=np.float64).reshape(3, n, 1)).to(self._ens.device)
        dy, bad = self._ens.derivatives(boundary, yy.contiguous())
        if int(bad[0]):
            raise ValueError("Temperature outside liquid water range [0.0, 100.0]°C")
        return dy.reshape(3 * n).cpu().numpy()

    def validate_conservation(self):
        """reactor.py:570-611 for this plant, computed by the batched diagnostics kernel."""
        e, s = self._ens, self.state
        e.set_state(s.pH[None, :], s.chlorine[None, :], s.temperature[None, :], time=[s.time])
        d = e.diagnostics()
        if int(d["bad"][0]):
            raise ValueError("Temperature outside liquid water range [0.0, 100.0]°C (thermodynamics.py:146-157)")
        out = {k: float(d[k][0]) for k in ("total_chlorine_mg", "total_H_mol", "total_OH_mol", "charge_balance_mol",
                                           "thermal_energy_kJ")}
        out["zones"] = self.config.n_zones
        out["timestamp"] = s.time
        return out

    def get_state_at_location(self, zone_idx: int, parameter: str) -> float:
        """reactor.py:543-568"""
        if zone_idx < 0 or zone_idx >= self.config.n_zones:
            raise ValueError(f"Invalid zone index: {zone_idx}")
        table = {"pH": self.state.pH, "chlorine": self.state.chlorine, "temperature": self.state.temperature,
                 "density": self.state.density}
        if parameter not in table:
            raise ValueError(f"Unknown parameter: {parameter}")
        return table[parameter][zone_idx]
