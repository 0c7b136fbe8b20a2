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
SENSOR_STATS = 22   # per sensor: valid count, shifted sum, shifted sum of squares, 12 status bins, 7 fault bins
N_SENSORS = 7


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


def stats_size(n_zones: int, sensors: bool = False) -> int:
    return STATS_HDR + 6 * n_zones + (N_SENSORS * SENSOR_STATS if sensors else 0)


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
    # shifts of the per-sensor moments (pH, pH, chlorine, chlorine, flow, temperature, temperature)
    sensor_shifts: Tuple[float, ...] = (7.0, 7.0, 2.0, 2.0, 10.0, 20.0, 20.0)

    def as_row(self) -> np.ndarray:
        return np.array([self.shift_pH, self.shift_chlorine, self.shift_temperature, self.chlorine_min,
                         self.pH_low, self.pH_high, self.temperature_max], dtype=np.float64)


def finalize_stats(vec: np.ndarray, n_zones: int, spec: StatsSpec) -> Dict[str, np.ndarray]:
    """Turn a (summed) statistics vector into means / variances / exceedance fractions."""
    vec = np.asarray(vec, dtype=np.float64)
    live = vec[0]
    out: Dict[str, np.ndarray] = {"live": live, "halted": vec[1], "degraded": vec[5], "pending_catch_up": vec[6]}
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
    off = STATS_HDR + 6 * n_zones
    if vec.size >= off + N_SENSORS * SENSOR_STATS:
        # per-sensor statistics of the last suite read over the live plants (reference analogue across time:
        # BaseSensor.get_statistics, base_sensor.py:809-856): valid fraction, mean / variance of the finite readings,
        # SensorStatus and SensorFault histograms (fractions of the live plants)
        sb = vec[off:off + N_SENSORS * SENSOR_STATS].reshape(N_SENSORS, SENSOR_STATS)
        nv = sb[:, 0]
        dv = np.where(nv > 0, nv, np.nan)
        m1 = sb[:, 1] / dv
        out["sensor_valid_count"] = nv
        out["sensor_valid_fraction"] = nv / denom
        out["sensor_mean"] = np.asarray(spec.sensor_shifts) + m1
        out["sensor_var"] = np.maximum(sb[:, 2] / dv - m1 * m1, 0.0)
        out["sensor_status_hist"] = sb[:, 3:15] / denom
        out["sensor_fault_hist"] = sb[:, 15:22] / denom
    return out


class EnsembleStatistics:
    """Device-side accumulation + all-reduce of the ensemble statistics of one PlantEnsemble shard."""

    def __init__(self, ensemble, spec: Optional[StatsSpec] = None, suite=None, out: Optional[torch.Tensor] = None):
        """``suite``: the ensemble's SensorSuite; its per-sensor statistics then follow the plant statistics in the
        vector (valid count, sum, sum of squares, status and fault histograms per sensor).
        ``out``: where the vector is written (e.g. this sub-ensemble's row of a rank's stack); default: an own buffer."""
        self.ens = ensemble
        self.spec = spec or StatsSpec()
        self.suite = suite
        L = _lib.lib()
        n = ensemble.n_zones
        self.size_plants = L.wt_stats_size(n)
        self.size = self.size_plants + (L.wt_sensor_stats_size() if suite is not None else 0)
        assert self.size == stats_size(n, suite is not None)
        dev = ensemble.device
        self._spec_dev = torch.from_numpy(self.spec.as_row()).to(dev)
        self._out = torch.zeros(self.size, dtype=torch.float64, device=dev) if out is None else out
        assert self._out.numel() == self.size and self._out.is_contiguous() and self._out.dtype == torch.float64
        self._scratch = torch.empty(L.wt_stats_scratch_doubles(n), dtype=torch.float64, device=dev)
        if suite is not None:
            self._shift7 = torch.tensor(self.spec.sensor_shifts, dtype=torch.float64, device=dev)
            self._scratch_s = torch.empty(L.wt_sensor_stats_scratch_doubles(), dtype=torch.float64, device=dev)

    def local(self) -> torch.Tensor:
        """Statistics vector of this shard (device tensor, overwritten on every call)."""
        e = self.ens
        p = lambda t: C.c_void_p(t.data_ptr())
        with torch.cuda.device(e.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = _lib.lib().wt_stats(e.n_plants, e.n_zones, p(e._y), p(e._status), p(self._spec_dev), p(self._out),
                                     p(self._scratch), 0, C.c_void_p(stream))
            _lib.check(rc, "wt_stats")
            if self.suite is not None:
                su = self.suite
                rc = _lib.lib().wt_sensor_stats(e.n_plants, p(su._out[0]), p(su._out_status), p(su._out_fault), p(e._status),
                                                p(self._shift7), p(self._out[self.size_plants:]), p(self._scratch_s), 0,
                                                C.c_void_p(stream))
                _lib.check(rc, "wt_sensor_stats")
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


class OverlappedAllReduce:
    """Sum all-reduce of a small vector BESIDE the work that follows it ("issued every K steps on a side
    stream overlapping the next step").  ``submit(vec)`` copies the vector to one of two staging buffers and starts an
    asynchronous all-reduce on it; the caller's stream only waits for a reduction when its buffer comes round again,
    two submissions later (or in ``finish``).  Works on any backend (NCCL: the waits are stream-level; gloo in the CPU
    tests: they block the host)."""

    def __init__(self, like: torch.Tensor, group=None):
        self.buf = [torch.zeros_like(like) for _ in range(2)]
        self.work = [None, None]
        self.k = 0
        self.group = group

    def submit(self, vec: torch.Tensor) -> torch.Tensor:
        """Returns the staging buffer that will hold the reduced vector (valid after ``finish`` or two submissions)."""
        import torch.distributed as dist
        k = self.k % 2
        if self.work[k] is not None:
            self.work[k].wait()          # the reduction of two submissions ago
        self.buf[k].copy_(vec)           # after the producer of `vec` in stream order
        self.work[k] = dist.all_reduce(self.buf[k], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.k += 1
        return self.buf[k]

    def finish(self) -> None:
        for k, w in enumerate(self.work):
            if w is not None:
                w.wait()
                self.work[k] = None


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
        nstat = stats_size(self.n_zones, self.suites is not None)
        # every sub-ensemble writes its vector straight into its row of the stack; wt_sum_rows adds the rows in order
        self._stack = torch.zeros((len(self.engines), nstat), dtype=torch.float64, device=self.device)
        self.stats_parts = [EnsembleStatistics(e, spec, None if self.suites is None else self.suites[i], out=self._stack[i])
                            for i, e in enumerate(self.engines)]
        self._sum = torch.zeros(nstat, dtype=torch.float64, device=self.device)
        self._graph = None
        self._overlapped = None
        # side streams of the deferral (PlantEnsemble(catch_up_attempts=...)): catch-up launches of one block overlap the
        # ordinary steps of the next
        self.defer = bool(self.engines[0].catch_up_attempts > 0)
        if self.defer:
            with torch.cuda.device(self.device):
                self.side = [torch.cuda.Stream(device=self.device) for _ in self.engines]
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

    # ---- host-buffer step: the state lives in pinned HOST memory between steps (the reference keeps it in Python
    # objects and hands the readings to its Modbus layer every step, __main__.py:398-457) -------------------------------
    def alloc_host_io(self):
        """Pinned host buffers of ``step_host``, one dict per sub-ensemble: in  y [3, n, P], time [P], bnd [10, P];
        out y, time, flow [P], status [P] and, with sensor suites, sensor [5, 7, P] (value, raw_value, uncertainty,
        quality, timestamp), sensor_status [7, P], sensor_fault [7, P]."""
        io = []
        for i, e in enumerate(self.engines):
            P, n = e.n_plants, e.n_zones
            pin = lambda *shape, dtype=torch.float64: torch.zeros(shape, dtype=dtype).pin_memory()
            d = {"y": pin(3, n, P), "time": pin(P), "bnd": pin(self.bnd[i].shape[0], P), "flow": pin(P),
                 "status": pin(P, dtype=torch.int32)}
            d["y"].copy_(e._y)
            d["time"].copy_(e._time)
            d["bnd"].copy_(self.bnd[i])
            if self.suites is not None:
                d.update(sensor=pin(5, 7, P), sensor_status=pin(7, P, dtype=torch.int32), sensor_fault=pin(7, P, dtype=torch.int32))
            io.append(d)
        return io

    def step_host(self, io, dt: float, read_time: Optional[float] = None) -> None:
        """One ``step(dt)`` (+ one suite read at ``read_time``) of every plant with inputs FROM and results TO the pinned
        host buffers of ``alloc_host_io``.  Every sub-ensemble runs upload -> step kernels -> sensor read -> download on
        its own stream, so the copies of one overlap the kernels of the others (the DMA engines serve the uploads in
        issue order).  Asynchronous: ``synchronize()`` + a stream synchronize before the host reads the buffers."""
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            b = io[i]
            with torch.cuda.stream(s):
                e._y.copy_(b["y"], non_blocking=True)
                e._time.copy_(b["time"], non_blocking=True)
                self.bnd[i].copy_(b["bnd"], non_blocking=True)
                e.step(dt, self.bnd[i])
                if self.suites is not None and read_time is not None:
                    self.suites[i].read(e.state, read_time)
                b["y"].copy_(e._y, non_blocking=True)
                b["time"].copy_(e._time, non_blocking=True)
                b["flow"].copy_(e._flow, non_blocking=True)
                b["status"].copy_(e._status, non_blocking=True)
                if self.suites is not None and read_time is not None:
                    su = self.suites[i]
                    b["sensor"].copy_(su._out, non_blocking=True)
                    b["sensor_status"].copy_(su._out_status, non_blocking=True)
                    b["sensor_fault"].copy_(su._out_fault, non_blocking=True)

    def host_io_bytes(self, io, sensors: bool = True):
        """(H2D, D2H) bytes of one ``step_host`` call."""
        nb = lambda t: t.numel() * t.element_size()
        h2d = sum(nb(b["y"]) + nb(b["time"]) + nb(b["bnd"]) for b in io)
        d2h = sum(nb(b["y"]) + nb(b["time"]) + nb(b["flow"]) + nb(b["status"])
                  + ((nb(b["sensor"]) + nb(b["sensor_status"]) + nb(b["sensor_fault"])) if sensors and "sensor" in b else 0) for b in io)
        return h2d, d2h

    def _block(self, n_steps: int, dt: float, read) -> None:
        """One block of steps of every sub-ensemble on its stream; ``read(i)`` reads sub-ensemble i's sensors.  With
        deferral: the plants that ran out of budget during the PREVIOUS block are collected at the start, caught up on
        the side stream (larger budget, until the end time of THIS block) while this block's steps run, and rejoined at
        its end -- every plant is time-aligned again at each block boundary."""
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            with torch.cuda.stream(s):
                if self.defer:
                    e.collect_deferred(n_steps * dt)
                    self.side[i].wait_stream(s)
                    with torch.cuda.stream(self.side[i]):
                        e.catch_up(2 * n_steps, dt, self.bnd[i])
                for _ in range(n_steps):
                    e.step(dt, self.bnd[i])
                    if self.suites is not None:
                        read(i)
                if self.defer:
                    s.wait_stream(self.side[i])
                    e.rejoin_deferred()

    def start_deferral(self, t_now: float) -> None:
        """The stop time of the catch-up launches starts at the ensemble's current time (then + n_steps dt per block)."""
        for e in self.engines:
            e._t_stop.fill_(float(t_now))

    def block(self, n_steps: int, dt: float, t_first: float) -> None:
        """``n_steps`` eager steps (+ sensor reads at t_first, t_first + dt, ...) with the block-wise deferral."""
        k = [0] * len(self.engines)

        def read(i):
            self.suites[i].read(self.engines[i].state, t_first + k[i] * dt)
            k[i] += 1
        self._block(n_steps, dt, read)

    def advance(self, n_steps: int, dt: float) -> None:
        for i, (e, s) in enumerate(zip(self.engines, self.streams)):
            with torch.cuda.stream(s):
                e.advance(n_steps, dt, self.bnd[i])

    def local_stats(self) -> torch.Tensor:
        """Statistics vector of this rank's shard (device tensor): the parts' kernels on their own streams, joined,
        summed in a fixed order."""
        for i, (sp, s) in enumerate(zip(self.stats_parts, self.streams)):
            with torch.cuda.stream(s):
                sp.local()   # -> self._stack[i]
        self.synchronize()
        with torch.cuda.device(self.device):
            rc = _lib.lib().wt_sum_rows(len(self.engines), self._sum.numel(), C.c_void_p(self._stack.data_ptr()),
                                        C.c_void_p(self._sum.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "wt_sum_rows")
        self.fork()  # the statistics buffers are reused by the next call
        return self._sum

    def stats(self, group=None) -> torch.Tensor:
        """Statistics vector of the whole shard, summed over the ranks of ``group`` (one all-reduce)."""
        v = self.local_stats()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        return v

    # ---- one CUDA graph per rank for a block of steps ------------------------------------------
    def capture(self, n_steps: int, dt: float, t_next: float, with_stats: bool = True) -> None:
        """Capture ``n_steps`` x (step + sensor read [+ cost order]) of every sub-ensemble, followed by the local
        statistics, into ONE CUDA graph: a replay is a single launch from the host instead of ~10 per part and step.
        The time of the sensor reads lives in a device clock (SensorSuite.start_clock) that the captured kernels
        advance themselves; the next read happens at ``t_next``.  Warm up (run at least one eager step) first: the
        first launch of a kernel sets its attributes, which cannot be captured."""
        if self.suites is not None:
            for su in self.suites:
                su.start_clock(t_next, dt)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(device=self.device)
        cap.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cap):
            with torch.cuda.graph(g, stream=cap, capture_error_mode="thread_local"):
                self.fork()
                self._block(n_steps, dt, lambda i: self.suites[i].read_clocked())
                if with_stats:
                    self.local_stats()
                self.synchronize()
        self._graph, self._graph_steps, self._graph_stats = g, int(n_steps), bool(with_stats)

    def replay(self, group=None, overlap: bool = False) -> Optional[torch.Tensor]:
        """Replay the captured block of steps on the current stream; returns the (all-reduced) statistics vector when the
        block was captured with statistics.

        ``overlap``: the all-reduce runs BESIDE the next block instead of in front of it.  A blocking all-reduce makes
        every rank wait for the slowest one at every block boundary (the ranks' blocks differ by the stragglers they
        hold); here the local vector is copied to one of two staging buffers, reduced asynchronously on NCCL's stream,
        and the compute stream only waits for a reduction when its buffer comes round again two blocks later.  The
        returned tensor is valid after ``finish_stats()`` (or two blocks later)."""
        self._graph.replay()
        if self.suites is not None:
            for su in self.suites:
                su.account_reads(self._graph_steps)
        if not self._graph_stats:
            return None
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if not overlap:
                dist.all_reduce(self._sum, op=dist.ReduceOp.SUM, group=group)
                return self._sum
            if self._overlapped is None:
                self._overlapped = OverlappedAllReduce(self._sum, group)
            return self._overlapped.submit(self._sum)
        return self._sum

    def finish_stats(self) -> None:
        """The compute stream waits for the outstanding overlapped all-reduces (``replay(overlap=True)``)."""
        if self._overlapped is not None:
            self._overlapped.finish()

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

    def deferred(self) -> int:
        self.synchronize()
        return int(sum(int(((e.status & _lib.ST_DEFERRED) != 0).sum()) for e in self.engines))

    def degraded(self) -> int:
        """Plants whose LAST step was completed by a floor-mode catch-up with a forced acceptance (WT_ST_DEGRADED)."""
        self.synchronize()
        return int(sum(int(((e.status & _lib.ST_DEGRADED) != 0).sum()) for e in self.engines))
