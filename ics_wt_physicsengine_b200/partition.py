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


class PipelinedShard:
    """One rank's shard run as ``parts`` independent sub-ensembles, each on its own CUDA stream.

    Plants never interact, so the sub-ensembles need no ordering among themselves: their streams are
    joined only when somebody reads results (``synchronize`` / ``stats``).  What this buys is the drain of
    every launch: a step kernel cannot end before its slowest warp, and while its last blocks finish the
    SMs run half empty -- about 0.2 ms per launch on B200, 8 % of a step at 131,072 plants (the per-GPU
    shard of the 8-GPU run).  With two or more streams the next launch of another sub-ensemble fills that
    drain (measured: tools/two_stream.py).  Results are identical to one big ensemble: the sensor noise is
    keyed by the GLOBAL plant id and the statistics vector is additive.
    """

    def __init__(self, ensemble: Ensemble, parts: int = 2, device=None, plant0: int = 0, sensor_seed: Optional[int] = None,
                 spec: Optional[StatsSpec] = None, **engine_kw):
        from .reactor import PlantEnsemble
        P = ensemble.n_plants
        parts = max(1, min(int(parts), P))
        self.n_plants, self.n_zones = P, ensemble.n_zones
        self.bounds = [shard_bounds(P, i, parts) for i in range(parts)]
        self.bounds = [b for b in self.bounds if b[1] > b[0]]
        self.engines = [PlantEnsemble(ensemble.slice(slice(lo, hi)), device=device, **engine_kw) for lo, hi in self.bounds]
        self.device = self.engines[0].device
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        self.bnd = [torch.from_numpy(np.ascontiguousarray(ensemble.bnd[lo:hi].T)).to(self.device) for lo, hi in self.bounds]
        self.suites = None
        if sensor_seed is not None:
            from .sensors import create_realistic_sensor_suite
            self.suites = [create_realistic_sensor_suite(e, seed=sensor_seed, plant0=plant0 + lo)
                           for e, (lo, _) in zip(self.engines, self.bounds)]
        self.stats_parts = [EnsembleStatistics(e, spec) for e in self.engines]
        self._sum = torch.zeros(self.stats_parts[0].size, dtype=torch.float64, device=self.device)
        self.fork()

    def fork(self) -> None:
        """Make every sub-ensemble stream wait for the work already queued on the caller's current stream."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def synchronize(self) -> None:
        """Join: the caller's current stream waits for every sub-ensemble stream (stream order, no host sync)."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def initialize_sensors(self, t0: float) -> None:
        for suite, s in zip(self.suites, self.streams):
            with torch.cuda.stream(s):
                suite.initialize(t0)

    def step(self, dt: float, read_time: Optional[float] = None) -> None:
        """One ``step(dt)`` of every plant with the ensemble's own boundary rows, then (if the shard has
        sensor suites and ``read_time`` is given) one read of the 7-sensor suite at ``read_time``."""
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            with torch.cuda.stream(s):
                e.step(dt, self.bnd[i])
                if self.suites is not None and read_time is not None:
                    self.suites[i].read(e.state, read_time)

    def advance(self, n_steps: int, dt: float) -> None:
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            with torch.cuda.stream(s):
                e.advance(n_steps, dt, self.bnd[i])

    def stats(self, group=None) -> torch.Tensor:
        """Statistics vector of the whole shard, summed over the ranks of ``group`` (one all-reduce)."""
        for sp, s in zip(self.stats_parts, self.streams):
            with torch.cuda.stream(s):
                sp.local()
        self.synchronize()
        torch.sum(torch.stack([sp._out for sp in self.stats_parts]), dim=0, out=self._sum)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self._sum, op=dist.ReduceOp.SUM, group=group)
        self.fork()  # the statistics buffers are reused by the next call
        return self._sum

    # aggregate views (they join the streams first)
    def time_sum(self) -> torch.Tensor:
        self.synchronize()
        return torch.stack([e.state.time.sum() for e in self.engines]).sum()

    def counters_sum(self) -> torch.Tensor:
        self.synchronize()
        return torch.stack([e.counters.sum(dim=1) for e in self.engines]).sum(dim=0)

    def reset_counters(self) -> None:
        self.synchronize()
        for e in self.engines:
            e.reset_counters()
        self.fork()

    def halted(self) -> int:
        self.synchronize()
        return int(sum(int(((e.status & _lib.ST_HALT_MASK) != 0).sum()) for e in self.engines))
