#!/usr/bin/env python
"""Headline benchmark: plant-zone-steps/s of the batched plant step on B200.

    python bench.py --gpus N --steps K --warmup W            (this engine)
    python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port of the
                                                              reference path on the host cores)

Workload (BASELINE.json configs[4], the one the metric is quoted on): 1,048,576 plants x 10
zones, synthetic inputs of ensembles.config5 (seed 20260004), dt = 1 s, sharded over the ranks
(strong scaling).  A "step" is one IntegratedCSTR.step(dt) of every plant = one kernel launch.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "plant-zone-steps/sec"
UNIT = "plant-zone-steps/s"
N_ZONES = 10
TOTAL_PLANTS = 1048576
DT = 1.0


def flops_alg(cnt_sum: np.ndarray, n_plant_steps: float, n_zones: int) -> float:
    """algorithmic fp64 flops from the emitted solver path counters.
    cnt_sum: totals over plants of (nfev, njev, nlu, nsteps, nnewton, ...)."""
    nfev, njev, nlu, nsteps, nnewton = (float(cnt_sum[i]) for i in range(5))
    per_zone = 310.0 * (nfev + 9.0 * njev) + 950.0 * (nlu / 2.0) + 740.0 * nnewton + 240.0 * nsteps \
        + 900.0 * n_plant_steps
    return per_zone * n_zones


def bytes_alg(n_plant_steps: float, n_zones: int) -> float:
    """48 + 216/n bytes per plant-zone-step (+24 with derived state)."""
    return n_plant_steps * n_zones * (48.0 + 24.0 + 216.0 / n_zones)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md): one streaming
    nvidia-smi process (-lms 100), started before and killed after the region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def _pump(self):
        try:
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
        except Exception:
            pass

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
            time.sleep(0.35)  # let the first sample arrive before the region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()
            if self.th is not None:
                self.th.join(timeout=3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")):
            if any(r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(pw) if pw else None}


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm (oracle port, see oracle/wt_oracle.h) on the host cores."""
    if rank != 0:
        return
    from ics_wt_physicsengine_b200 import ensembles
    from oracle import wt_oracle as wo

    cores = os.cpu_count() or 1
    sample = args.cpu_plants
    e = ensembles.config5(TOTAL_PLANTS if sample > 65536 else 65536, N_ZONES).slice(slice(0, sample))
    par = wo.derive_params(e.cfg, N_ZONES)
    bnd = np.ascontiguousarray(e.bnd)
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1).copy()
    t = np.zeros(sample)
    wo.set_max_attempts(args.max_attempts)
    halted = np.zeros(sample, bool)

    def one_step():
        nonlocal y, t
        idx = np.nonzero(~halted)[0]
        ya, ta = y[idx].copy(), t[idx].copy()
        st, _, _ = wo.step_batch(par[idx].copy(), bnd[idx].copy(), N_ZONES, ta, ya, dt=DT, nthreads=cores)
        y[idx], t[idx] = ya, ta
        halted[idx[(st & wo.ST_HALT_MASK) != 0]] = True
        return idx.size

    for _ in range(args.warmup):
        one_step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += one_step()
    el = time.perf_counter() - t0
    val = done * N_ZONES / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config5 physics: first {sample} of {TOTAL_PLANTS} plants x {N_ZONES} zones, dt=1s "
                               f"(bounded CPU sample)", "max_attempts": args.max_attempts},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} plants x {args.steps} steps, oracle C port (pthreads), "
                                   "the Python reference itself does not exist on the GPU box"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--plants", type=int, default=TOTAL_PLANTS)
    ap.add_argument("--max-attempts", type=int, default=64)
    ap.add_argument("--cpu-plants", type=int, default=16384)
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fused", action="store_true", help="time one fused wt_advance(K) launch instead of K launches")
    ap.add_argument("--sort-every", type=int, default=2, help="re-order plants by last-step work every k launches (scheduling only)")
    ap.add_argument("--no-sensors", action="store_true", help="physics only (BASELINE configs[1]-style step)")
    ap.add_argument("--streams", type=int, default=4, help="independent sub-ensembles (CUDA streams) per GPU")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from ics_wt_physicsengine_b200 import PlantEnsemble, _lib, ensembles

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    P_total = args.plants
    from ics_wt_physicsengine_b200.partition import shard_bounds
    lo, hi = shard_bounds(P_total, rank, world)
    full = ensembles.config5(P_total, N_ZONES)
    e = full.slice(slice(lo, hi))
    P = e.n_plants
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    # the rank's shard as `--streams` independent sub-ensembles on their own CUDA streams (plants never
    # interact): the drain of one launch is filled by the next launch of another sub-ensemble
    shard = PipelinedShard(e, parts=args.streams, device=dev, plant0=lo,
                           sensor_seed=None if (args.no_sensors or args.fused) else 20260004,
                           max_attempts=args.max_attempts, sort_every=args.sort_every)
    eng = shard.engines[0]
    fp64_peak = _lib.measure_fp64_peak() if rank == 0 else 0.0
    if shard.suites is not None:
        shard.initialize_sensors(0.0)
    sim = {"k": 0}

    def do_steps(k, timed=False):
        if args.fused:
            shard.advance(k, DT)
            return
        for i in range(k):
            shard.step(DT, read_time=float(sim["k"]))   # step + sensor read (__main__.py:403-410), per sub-ensemble stream
            sim["k"] += 1
            if (i + 1) % 10 == 0:
                shard.stats()   # wt_stats kernels + one NCCL sum all-reduce of the statistics vector

    do_steps(args.warmup)
    shard.reset_counters()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_before = shard.time_sum().clone()
    shard.fork()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        torch.cuda.synchronize()
        ev0.record()
        shard.fork()
        do_steps(args.steps, timed=True)
        shard.synchronize()
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    # plant-steps actually completed inside the timed region: time advances by DT per completed
    # plant-step and halted plants stop advancing, so nothing skipped is ever credited
    done = (shard.time_sum() - t_before) / DT
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.cat([shard.counters_sum().to(torch.float64), done.reshape(1)])
    if world > 1:
        dist.barrier()
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg)
    ms = float(tms[0])
    cnt_sum = agg[:8].cpu().numpy()
    timed_plant_steps = float(agg[8])
    halted_after = shard.halted()
    value = timed_plant_steps * N_ZONES / (ms * 1e-3)

    # ---- the dominant kernel alone: the sub-ensembles' launches overlap on the device inside the timed region, so one
    # launch is timed here with CUDA events on the stream it is launched on, streams joined between launches
    # (3 more steps of every sub-ensemble, after the timed region; rank 0 only)
    def kernel_alone(rounds=3):
        shard.reset_counters()
        t0_sum = shard.time_sum().clone()
        step_t = sens_t = 0.0
        n_launch = 0
        for _ in range(rounds):
            for j, en in enumerate(shard.engines):
                torch.cuda.synchronize()
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                with torch.cuda.stream(shard.streams[j]):
                    e0.record()
                    en.step(DT, shard.bnd[j])
                    e1.record()
                    if shard.suites is not None:
                        shard.suites[j].read(en.state, float(sim["k"]))
                    e2.record()
                torch.cuda.synchronize()
                step_t += e0.elapsed_time(e1)
                sens_t += e1.elapsed_time(e2)
                n_launch += 1
            sim["k"] += 1
        done_cal = float((shard.time_sum() - t0_sum) / DT)
        return step_t, sens_t, n_launch, done_cal, shard.counters_sum().to(torch.float64).cpu().numpy()

    cal = kernel_alone() if (rank == 0 and not args.fused) else None

    # ---- e2e: the C-ABI host-buffer call (H2D + step + D2H inside the timed region) on EVERY rank's
    # shard at the same time; whole-job value = all plants / slowest rank
    if world > 1:
        dist.barrier()
    live_e2e, el_e2e, k_e2e, h2d, d2h = e2e_measure(e, shard, args)
    e2e_agg = torch.tensor([float(live_e2e), el_e2e, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    e2e_max = e2e_agg.clone()
    if world > 1:
        dist.all_reduce(e2e_agg)
        dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    e2e = {"value": float(e2e_agg[0]) * N_ZONES * k_e2e / float(e2e_max[1]), "unit": UNIT,
           "h2d_bytes_per_step": int(e2e_agg[2]), "d2h_bytes_per_step": int(e2e_agg[3]), "steps": k_e2e,
           "api": "wt_step_host (C ABI, pinned host buffers: H2D of state+boundary, step, D2H of state+time+flow+status per call, "
                  "pipelined over column slabs on three streams; per-plant constants resident after the first call)", "n_gpus": world}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # Headline roofline: the whole job's algorithmic flops over the whole timed region (conservative: the region also
    # holds the sensor and statistics kernels, and the sub-ensembles' launches overlap).  Beside it: the step kernel
    # alone, one launch at a time, from the calibration steps above with their own path counters.
    F = flops_alg(cnt_sum, timed_plant_steps, N_ZONES)
    parts = len(shard.engines)
    kernel_region_ms = ms
    achieved_tf = F / world / (kernel_region_ms * 1e-3) / 1e12  # per GPU
    alone = None
    if cal is not None:
        step_t, sens_t, n_launch, done_cal, cnt_cal = cal
        f_cal = flops_alg(cnt_cal, done_cal, N_ZONES)
        alone = {"ms_per_launch": step_t / n_launch, "plants_per_launch": P // parts, "launches_timed": n_launch,
                 "achieved": f_cal / (step_t * 1e-3) / 1e12, "frac": f_cal / (step_t * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                 "sensor_kernel_ms_per_launch": sens_t / n_launch, "step_share_of_gpu_time": step_t / (step_t + sens_t),
                 "basis": "CUDA events on the launching stream, launches serialised (3 extra steps after the timed region)"}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = bytes_alg(timed_plant_steps, N_ZONES) / world / (kernel_region_ms * 1e-3) / 1e9

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        cpu = cpu_baseline(args)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"BASELINE configs[4] physics: {P_total} plants x {N_ZONES} zones (ensembles.config5 seed 20260004), "
                        f"IntegratedCSTR.step(dt=1s) + 7-sensor suite read per plant per step, sharded over {world} GPU(s)",
            "launch_mode": "fused wt_advance(K)" if args.fused else "one wt_step launch per sub-ensemble per step",
            "l2": "state+params per GPU >> 126 MB L2 at N<=4; inputs larger than L2 (no flush needed)",
            "sensor_suite": shard.suites is not None, "sort_every": args.sort_every,
            "sub_ensembles_per_gpu": len(shard.engines),
            "max_attempts": args.max_attempts, "plants_halted_at_end_rank0": halted_after,
            "stats_allreduce_every": 10,
        },
        "roofline": {
            "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if fp64_peak else None,
            # dram__bytes_read+write of one wt_step launch: 81.1 B per plant-zone-step measured by ncu --set full on
            # the 262,144-plant launch (profiles/r1_step_kernel_v5_session2_final.txt), scaled to this launch's units
            "traffic": 81.1 * (timed_plant_steps / max(1, args.steps)) * N_ZONES / world,
            "traffic_source": "ncu capture of the 262144-plant launch scaled by units (profiles/r1_step_kernel_v5_session2_final.txt)",
            "peak_source": "measured in this run by wt_measure_fp64_peak (8 independent DFMA chains/thread); "
                           "MEASURED_PEAKS.json carries no FP64 figure",
            "flops_model": "SURVEY 8(d): 310(nfev+9njev)+950(nlu/2)+740 newton+240 steps+900 per zone, from emitted counters",
            "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": hbm_peak, "frac": hbm_ach / hbm_peak},
            "kernel": "wt_step_kernel", "kernel_ms_per_launch": alone["ms_per_launch"] if alone else None,
            "achieved_basis": "whole timed region (overlapping launches of the sub-ensembles; includes sensor/stats kernels)",
            "kernel_alone": alone,
            "counters_per_plant_step": {k: float(cnt_sum[i]) / timed_plant_steps for i, k in enumerate(_lib.CNT_NAMES)},
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": len(shard.engines) * (1 if args.fused else args.steps * (2 if shard.suites is not None else 1)
                                                 + 2 * (args.steps // 10)),
        "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_measure(e, shard, args):
    """Same metric through the C-ABI host-buffer entry point wt_step_host: state and boundary start in
    pinned HOST memory every step; the call copies them in, steps, and copies state, time, flow and
    status back.  Returns (live plants, seconds, steps, h2d bytes/step, d2h bytes/step) of this shard."""
    import ctypes as C

    import torch

    from ics_wt_physicsengine_b200 import _lib

    P, n = e.n_plants, e.n_zones
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    par = pin(np.concatenate([x.par_host for x in shard.engines]).T)
    bnd = pin(e.bnd.T)
    y = pin(np.stack([e.pH0.T, e.Cl0.T, e.T0.T]))
    t = torch.zeros(P, dtype=torch.float64).pin_memory()
    flow = torch.zeros(P, dtype=torch.float64).pin_memory()
    st = torch.zeros(P, dtype=torch.int32).pin_memory()
    p = lambda x: C.c_void_p(x.data_ptr())
    L = _lib.lib()
    k = max(3, min(args.steps, 10))
    for i in range(3):
        _lib.check(L.wt_step_host(P, n, DT, p(par), p(bnd), P, p(t), p(y), p(flow), p(st), args.max_attempts, 1 if i else 0),
                   "wt_step_host")
    t0 = time.perf_counter()
    for _ in range(k):
        _lib.check(L.wt_step_host(P, n, DT, p(par), p(bnd), P, p(t), p(y), p(flow), p(st), args.max_attempts, 1), "wt_step_host")
    el = time.perf_counter() - t0
    live = int(((st & _lib.ST_HALT_MASK) == 0).sum())
    h2d = (10 + 1 + 3 * n + 1) * P * 8 + P * 4
    d2h = (1 + 3 * n + 1) * P * 8 + P * 4
    return live, el, k, h2d, d2h


def cpu_baseline(args):
    from ics_wt_physicsengine_b200 import ensembles
    from oracle import wt_oracle as wo

    cores = os.cpu_count() or 1
    sample = args.cpu_plants
    e = ensembles.config5(65536, N_ZONES).slice(slice(0, sample))
    par = wo.derive_params(e.cfg, N_ZONES)
    bnd = np.ascontiguousarray(e.bnd)
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1).copy()
    t = np.zeros(sample)
    wo.set_max_attempts(args.max_attempts)
    wo.step_batch(par, bnd, N_ZONES, t, y, dt=DT, nsteps=1, nthreads=cores)
    t0 = time.perf_counter()
    wo.step_batch(par, bnd, N_ZONES, t, y, dt=DT, nsteps=args.cpu_steps, nthreads=cores)
    el = time.perf_counter() - t0
    return {"value": sample * N_ZONES * args.cpu_steps / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} plants of config5(65536) x {args.cpu_steps} steps on {cores} host threads "
                      "(oracle C port of reactor.py + scipy Radau; the Python reference runs ~1e3 zone-steps/s/core, BASELINE.md)"}


if __name__ == "__main__":
    main()
