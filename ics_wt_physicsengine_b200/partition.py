"""Ensemble partitioner and the one collective of the multi-GPU path.

Plants are independent, so an ensemble shards into contiguous blocks of ceil(P/G) plants, one
process per GPU, with NO exchange while stepping.  The only collective is a sum all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests) of a small fp64 statistics vector produced by
the ``wt_stats`` kernel: live/halted counts, exceedance counts and shifted first/second moments
per (variable, zone).  Counts are carried as fp64 (exact below 2**53) so one all-reduce moves
everything.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .ensembles import Ensemble

STATS_HDR = 8
VARS = ("pH", "chlorine", "temperature")


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of plants owned by `rank` (ceil(P/G) per rank, last may be short)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per = (total + world - 1) // world
    lo = min(rank * per, total)
    return lo, min(lo + per, total)


def shard_ensemble(e: Ensemble, rank: int, world: int) -> Ensemble:
    lo, hi = shard_bounds(e.n_plants, rank, world)
    return e.slice(slice(lo, hi))


def stats_size(n_zones: int) -> int:
    return STATS_HDR + 6 * n_zones


@dataclass
class StatsSpec:
    """Shifts (for variance stability) and exceedance thresholds of the statistics vector."""

    shift_pH: float = 7.0
    shift_chlorine: float = 2.0
    shift_temperature: float = 20.0
    chlorine_min: float = 0.2     # outlet residual below the disinfection minimum [mg/L]
    pH_low: float = 6.5
    pH_high: float = 8.5
    temperature_max: float = 30.0

    def as_row(self) -> np.ndarray:
        return np.array([self.shift_pH, self.shift_chlorine, self.shift_temperature, self.chlorine_min,
                         self.pH_low, self.pH_high, self.temperature_max], dtype=np.float64)


def finalize_stats(vec: np.ndarray, n_zones: int, spec: StatsSpec) -> Dict[str, np.ndarray]:
    """Turn a (summed) statistics vector into means / variances / exceedance fractions."""
    vec = np.asarray(vec, dtype=np.float64)
    live = vec[0]
    out: Dict[str, np.ndarray] = {"live": live, "halted": vec[1]}
    denom = live if live > 0 else np.nan
    out["frac_outlet_chlorine_low"] = vec[2] / denom
    out["frac_outlet_pH_out_of_band"] = vec[3] / denom
    out["frac_outlet_temperature_high"] = vec[4] / denom
    shifts = (spec.shift_pH, spec.shift_chlorine, spec.shift_temperature)
    body = vec[STATS_HDR:STATS_HDR + 6 * n_zones].reshape(3, n_zones, 2)
    for v, name in enumerate(VARS):
        m1 = body[v, :, 0] / denom
        out[f"mean_{name}"] = shifts[v] + m1
        out[f"var_{name}"] = np.maximum(body[v, :, 1] / denom - m1 * m1, 0.0)  # population variance
    return out


class EnsembleStatistics:
    """Device-side accumulation + all-reduce of the ensemble statistics of one PlantEnsemble shard."""

    def __init__(self, ensemble, spec: Optional[StatsSpec] = None):
        self.ens = ensemble
        self.spec = spec or StatsSpec()
        L = _lib.lib()
        n = ensemble.n_zones
        self.size = L.wt_stats_size(n)
        dev = ensemble.device
        self._spec_dev = torch.from_numpy(self.spec.as_row()).to(dev)
        self._out = torch.zeros(self.size, dtype=torch.float64, device=dev)
        self._scratch = torch.empty(L.wt_stats_scratch_doubles(n), dtype=torch.float64, device=dev)

    def local(self) -> torch.Tensor:
        """Statistics vector of this shard (device tensor, overwritten on every call)."""
        e = self.ens
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(e.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.lib().wt_stats(e.n_plants, e.n_zones, p(e._y), p(e._status), p(self._spec_dev), p(self._out),
                                     p(self._scratch), 0, C.c_void_p(stream))
        _lib.check(rc, "wt_stats")
        return self._out

    def allreduce(self, group=None) -> torch.Tensor:
        """Shard statistics summed over all ranks (NCCL all-reduce; a no-op without a process group)."""
        v = self.local()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        return v

    def result(self, group=None) -> Dict[str, np.ndarray]:
        return finalize_stats(self.allreduce(group).cpu().numpy(), self.ens.n_zones, self.spec)
