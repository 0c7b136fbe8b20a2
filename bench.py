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
    e = ensembles.config5(args.plants, N_ZONES).slice(slice(0, sample))   # the first plants of the GPU arm's ensemble
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
        bad = (st & wo.ST_HALT_MASK) != 0
        halted[idx[bad]] = True
        return int((~bad).sum())   # completed plant-steps only

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
        "config": {"workload": f"config5 physics: first {sample} of {args.plants} plants x {N_ZONES} zones, dt=1s "
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
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying one CUDA graph per block of steps")
    ap.add_argument("--sort-every", type=int, default=1, help="re-order plants by last-step work every k launches (scheduling only)")
    ap.add_argument("--no-sensors", action="store_true", help="physics only (BASELINE configs[1]-style step)")
    ap.add_argument("--streams", type=int, default=0, help="independent sub-ensembles (CUDA streams) per GPU; 0 = by shard size")
    ap.add_argument("--stats-every", type=int, default=10)
    ap.add_argument("--blocking-allreduce", action="store_true",
                    help="all-reduce the statistics in front of the next block instead of beside it")
    ap.add_argument("--catch-up-attempts", type=int, default=0,
                    help="budget of the side-stream catch-up of plants that exhaust --max-attempts (0 = halt them for good, or "
                         "8 x --catch-up-floor-div when that is given)")
    ap.add_argument("--catch-up-floor-div", type=int, default=8,
                    help="0: no deferral unless --catch-up-attempts is given.  > 0: the catch-up runs in floor mode (step sizes >= dt / floor_div, forced acceptance at the floor, "
                         "WT_ST_DEGRADED): bounded-cost continuation of the plants on the 8 C density discontinuity, DESIGN.md section 7")
    args = ap.parse_args()
    if args.catch_up_floor_div > 0 and args.catch_up_attempts <= 0:
        args.catch_up_attempts = 8 * args.catch_up_floor_div

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
    from ics_wt_physicsengine_b200.partition import PipelinedShard, shard_bounds
    lo, hi = shard_bounds(P_total, rank, world)
    full = ensembles.config5(P_total, N_ZONES)
    e = full.slice(slice(lo, hi))
    P = e.n_plants
    # the rank's shard as independent sub-ensembles on their own CUDA streams (plants never interact): the drain of one
    # launch is filled by the next launch of another sub-ensemble.  No sub-ensemble below ~64k plants.
    parts = args.streams if args.streams > 0 else max(1, min(4, P // 65536))
    sensors_on = not args.no_sensors
    shard = PipelinedShard(e, parts=parts, device=dev, plant0=lo, sensor_seed=20260004 if sensors_on else None,
                           max_attempts=args.max_attempts, sort_every=args.sort_every, catch_up_attempts=args.catch_up_attempts,
                           catch_up_floor_div=args.catch_up_floor_div)
    fp64_peak = _lib.measure_fp64_peak() if rank == 0 else 0.0
    # Sensors: calibrated at t = -2000 s, then 100 reads of the initial state at t = -100 .. -1 s: when the timed region
    # starts every sensor is past its warm-up (10 / 30 / 60 / 300 / 1800 s) and the four 100-slot delay rings per plant
    # are full, i.e. the read kernel does its steady-state work (ring scans included).
    sim = {"k": 0}
    if sensors_on:
        shard.initialize_sensors(-2000.0)
        for j in range(100):
            for su, s in zip(shard.suites, shard.streams):
                with torch.cuda.stream(s):
                    su.read(None, float(j - 100))
    block = max(1, args.stats_every)

    def eager_steps(k):
        for i in range(k):
            shard.step(DT, read_time=float(sim["k"]) if sensors_on else None)   # step + sensor read (__main__.py:403-410)
            sim["k"] += 1
            if sim["k"] % block == 0:
                shard.stats()   # wt_stats + wt_sensor_stats kernels, one NCCL sum all-reduce

    eager_steps(args.warmup)
    use_graph = not args.no_graph and args.steps >= block
    n_blocks = args.steps // block if use_graph else 0
    rem = args.steps - n_blocks * block
    if use_graph:
        # align the block boundary with the statistics interval, then capture one block: `block` x (step, sensor read,
        # cost order) of every sub-ensemble + the local statistics kernels = ONE launch from the host per block
        while sim["k"] % block:
            eager_steps(1)
        shard.synchronize()
        if shard.defer:
            shard.start_deferral(float(sim["k"]))
        shard.capture(block, DT, t_next=float(sim["k"]), with_stats=True)
        # two untimed replays: graph upload, and both staging buffers of the overlapped all-reduce allocated and used once
        for _ in range(2):
            shard.replay(overlap=not args.blocking_allreduce)
            sim["k"] += block
        shard.finish_stats()
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
        stats_out = None
        for _ in range(n_blocks):
            # + the NCCL all-reduce of the statistics vector, overlapped with the next block (two staging buffers)
            stats_out = shard.replay(overlap=not args.blocking_allreduce)
            sim["k"] += block
        shard.finish_stats()      # every reduction complete inside the timed region
        if rem or not use_graph:
            shard.fork()
            eager_steps(rem if use_graph else args.steps)
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
    deferred_after = shard.deferred() if shard.defer else 0
    degraded_after = shard.degraded()
    # plants that have fallen behind for good: more than two blocks behind the front (a deferred plant is at most two
    # blocks behind: over budget in one block, caught up during the next)
    t_all = torch.cat([en.state.time for en in shard.engines])
    stalled_after = int((t_all < t_all.max() - 2 * block * DT - 0.5 * DT).sum())
    value = timed_plant_steps * N_ZONES / (ms * 1e-3)
    stats_vec = (stats_out if (use_graph and n_blocks and stats_out is not None) else shard._sum).cpu().numpy().copy()

    # ---- the kernels alone: the sub-ensembles' launches overlap on the device inside the timed region, so single launches
    # are timed here with CUDA events on the stream they are launched on, streams joined between launches
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

    cal = kernel_alone() if rank == 0 else None
    k2 = calc_ph_roofline(dev, fp64_peak) if rank == 0 else None

    # ---- e2e: the C-ABI host-buffer call (H2D + step + D2H inside the timed region) on EVERY rank's
    # shard at the same time; whole-job value = all plants / slowest rank
    if world > 1:
        dist.barrier()
    live_e2e, el_e2e, k_e2e, h2d, d2h, floor_e2e = e2e_measure(e, shard, args)
    e2e_agg = torch.tensor([float(live_e2e), el_e2e, float(h2d), float(d2h), floor_e2e], dtype=torch.float64, device=dev)
    e2e_max = e2e_agg.clone()
    if world > 1:
        dist.all_reduce(e2e_agg)
        dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    e2e = {"value": float(e2e_agg[0]) * N_ZONES * k_e2e / float(e2e_max[1]), "unit": UNIT,
           "h2d_bytes_per_step": int(e2e_agg[2]), "d2h_bytes_per_step": int(e2e_agg[3]), "steps": k_e2e,
           "work": "physics step only (IntegratedCSTR.step of every plant); the sensor suite is not part of this call",
           "copy_only_floor_value": float(e2e_agg[0]) * N_ZONES * k_e2e / float(e2e_max[4]),
           "copy_only_floor": "the same H2D and D2H bytes per step moved concurrently on two streams with no kernel, all ranks at "
                              "once (max over ranks): what the host links of this box allow for this call",
           "api": "wt_step_host (C ABI, pinned host buffers: H2D of state+boundary, step, D2H of state+time+flow+status per call, "
                  "pipelined over column slabs that ramp up and down (16k .. 128k .. 16k plants) on a copy-in, three compute and a "
                  "copy-out stream; per-plant constants resident after the first call)", "n_gpus": world}

    # ---- the same work as `value` (step + suite read) from and to host buffers, through the Python API
    e2e_s = None
    if sensors_on:
        if world > 1:
            dist.barrier()
        live_s, el_s, k_s, h2d_s, d2h_s, parts_s = e2e_sensors_measure(e, lo, dev, args)
        ag = torch.tensor([float(live_s), float(h2d_s), float(d2h_s)], dtype=torch.float64, device=dev)
        mx = torch.tensor([el_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ag)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        e2e_s = {"value": float(ag[0]) * N_ZONES * k_s / float(mx[0]), "unit": UNIT, "h2d_bytes_per_step": int(ag[1]),
                 "d2h_bytes_per_step": int(ag[2]), "steps": k_s, "sub_ensembles_per_gpu": parts_s,
                 "work": "IntegratedCSTR.step of every plant + one read of the 7-sensor suite per plant (the work `value` times)",
                 "api": "PipelinedShard.step_host (Python API over wt_advance + wt_sensors_read): state, time and boundary rows from "
                        "pinned host buffers, state, time, flow, status and the 7 x 5 sensor outputs + status / fault words back to them "
                        "every step; every sub-ensemble uploads, steps, reads and downloads on its own stream"}
    e2e["with_sensors"] = e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # Headline roofline: the whole job's algorithmic flops over the whole timed region (conservative: the region also
    # holds the sensor and statistics kernels, and the sub-ensembles' launches overlap).  Beside it: the step kernels
    # alone, one step at a time, from the calibration steps above with their own path counters.
    F = flops_alg(cnt_sum, timed_plant_steps, N_ZONES)
    nparts = len(shard.engines)
    achieved_tf = F / world / (ms * 1e-3) / 1e12  # per GPU
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alone = sensors_rf = None
    if cal is not None:
        step_t, sens_t, n_launch, done_cal, cnt_cal = cal
        f_cal = flops_alg(cnt_cal, done_cal, N_ZONES)
        ppl = P // nparts
        alone = {"ms_per_step_launch_pair": step_t / n_launch, "plants_per_launch": ppl, "launches_timed": n_launch,
                 "achieved": f_cal / (step_t * 1e-3) / 1e12, "frac": f_cal / (step_t * 1e-3) / 1e12 / fp64_peak if fp64_peak else None,
                 "step_share_of_gpu_time": step_t / (step_t + sens_t),
                 "basis": "CUDA events on the launching stream around wt_step_begin_kernel + wt_step_run_kernel, launches "
                          "serialised (3 extra steps after the timed region)"}
        if shard.suites is not None:
            # K3: algorithmic bytes of one suite read per plant: 7 sensors x (10 state doubles r/w + 4 ints r/w
            # + 5 outputs + 2 ints) + 4 delay-line searches x (8 timestamps + 1 value) x 8 B + 4 ring appends x 16 B + 6 state doubles
            b_read = 7 * (10 * 8 * 2 + 4 * 4 * 2 + 5 * 8 + 2 * 4) + 4 * 9 * 8 + 4 * 16 + 6 * 8
            ach = b_read * ppl / (sens_t / n_launch * 1e-3) / 1e9
            sensors_rf = {"kernel": "wt_sensors_read_kernel", "bound": "hbm", "ms_per_launch": sens_t / n_launch,
                          "algorithmic_bytes_per_plant_read": b_read, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                          "frac": ach / hbm_peak, "peak_source": hbm_src,
                          "state": "all sensors warm, delay rings full (one window of 8 timestamps read per search)",
                          "traffic": None, "traffic_source": "not captured for this kernel revision (profiles/r2_sensor_kernel.txt is the "
                                                             "full-scan revision: 5,030 B per plant-read)"}
    hbm_ach = bytes_alg(timed_plant_steps, N_ZONES) / world / (ms * 1e-3) / 1e9

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # rank 0 at N=1 only
        cpu = cpu_baseline(args)

    from ics_wt_physicsengine_b200.partition import StatsSpec, finalize_stats
    fin = finalize_stats(stats_vec, N_ZONES, StatsSpec())
    # kernels of THIS repo launched inside the timed region (per rank; graph replays launch them as graph nodes):
    per_part_step = 2 + (2 if sensors_on else 0)                                  # begin, run, sensor read, clock tick
    per_part_sort = 3 * (args.steps // max(1, args.sort_every)) if args.sort_every > 0 else 0   # hist, scan, scatter (+ a memset node)
    blocks_timed = args.steps // block
    per_part_stats = (2 + (2 if sensors_on else 0)) * blocks_timed                # wt_stats (2), wt_sensor_stats (2)
    per_part_defer = 3 * blocks_timed if (shard.defer and use_graph) else 0       # collect, the fused catch-up, rejoin
    per_rank_stats = blocks_timed                                                 # wt_sum_rows
    n_launches = nparts * (per_part_step * args.steps + per_part_sort + per_part_stats + per_part_defer) + per_rank_stats
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"BASELINE configs[4] physics: {P_total} plants x {N_ZONES} zones (ensembles.config5 seed 20260004), "
                        f"IntegratedCSTR.step(dt=1s) + 7-sensor suite read per plant per step, sharded over {world} GPU(s)",
            "launch_mode": (f"one CUDA graph per rank per {block} steps (step begin + run kernels, sensor read, clock tick, cost "
                            f"order of every sub-ensemble, local statistics), then one NCCL all-reduce "
                            + ("in front of the next block" if args.blocking_allreduce else "overlapped with the next block")) if use_graph
                           else "every kernel launched from the host",
            "l2": "state+params per GPU >> 126 MB L2 at N<=4; inputs larger than L2 (no flush needed)",
            "sensor_suite": shard.suites is not None,
            "sensor_state": "warm (calibrated at t=-2000 s), delay rings full" if sensors_on else None,
            "sort_every": args.sort_every, "sub_ensembles_per_gpu": nparts,
            "max_attempts": args.max_attempts, "catch_up_attempts": args.catch_up_attempts,
            "catch_up_floor_div": args.catch_up_floor_div,
            "deferral": ("plants that exhaust max_attempts are collected at the next block boundary, continued with catch_up_attempts "
                         + ("in floor mode (step sizes >= dt / catch_up_floor_div, forced acceptance at the floor, WT_ST_DEGRADED) "
                            if args.catch_up_floor_div > 0 else "")
                         + "on a side stream during that block and rejoined at its end (inside the graph)") if shard.defer and use_graph
                        else "off (eager launches)" if shard.defer else "off",
            "plants_halted_at_end_rank0": halted_after, "plants_deferred_at_end_rank0": deferred_after,
            "halted_fraction_rank0": halted_after / P,
            "plants_degraded_in_last_step_rank0": degraded_after,
            "plants_stalled_rank0": stalled_after, "stalled_fraction_rank0": stalled_after / P,
            "stalled": "plants more than two statistics blocks behind the front at the end of the timed region (halted for good; "
                       "'halted' above also counts plants that ran over the budget in the last blocks and are waiting for their catch-up)",
            "stats_allreduce_every": block, "stats_vector_doubles": int(stats_vec.size),
        },
        "roofline": {
            "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved_tf / fp64_peak if fp64_peak else None,
            # dram__bytes_read+write of one step (begin + run kernels): per plant-zone-step from the ncu --set full capture of
            # the 262,144-plant launches, scaled to this run's units
            "traffic": TRAFFIC_B_PER_UNIT * (timed_plant_steps / max(1, args.steps)) * N_ZONES / world,
            "traffic_source": f"modelled: {TRAFFIC_B_PER_UNIT} B per plant-zone-step measured by ncu on the 262144-plant launches "
                              "(profiles/r2_step_kernels.txt) x this run's units; it includes the begin -> run hand-off rows "
                              "(2 x 614 B per plant-zone-step written and read back; begin kernel 235 B, run kernel 316 B per plant-zone-step of DRAM traffic)",
            "peak_source": "measured in this run by wt_measure_fp64_peak (8 independent DFMA chains/thread); "
                           "MEASURED_PEAKS.json carries no FP64 figure (nominal 37.2 TFLOP/s at 1965 MHz)",
            "flops_model": "SURVEY 8(d): 310(nfev+9njev)+950(nlu/2)+740 newton+240 steps+900 per zone, from emitted counters",
            "hbm": {"achieved_gbs": hbm_ach, "peak_gbs": hbm_peak, "frac": hbm_ach / hbm_peak},
            "kernel": "wt_step_begin_kernel + wt_step_run_kernel",
            "kernel_ms_per_launch": alone["ms_per_step_launch_pair"] if alone else None,
            "achieved_basis": "whole timed region (overlapping launches of the sub-ensembles; includes sensor/stats kernels)",
            "kernel_alone": alone,
            "counters_per_plant_step": {k: float(cnt_sum[i]) / timed_plant_steps for i, k in enumerate(_lib.CNT_NAMES)},
            "sensors": sensors_rf,
            "calc_ph": k2,
        },
        "ensemble_statistics": {
            "live": float(fin["live"]), "halted": float(fin["halted"]),
            "mean_outlet_pH": float(fin["mean_pH"][-1]), "mean_outlet_chlorine": float(fin["mean_chlorine"][-1]),
            "sensor_valid_fraction": [float(x) for x in fin.get("sensor_valid_fraction", [])],
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": n_launches,
        "host_launches": (n_blocks + (nparts * (per_part_step * rem)) if use_graph else None),
        "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read + dram__bytes_write of one step per plant-zone-step (ncu --set full, 262,144 x 10 launches)
TRAFFIC_B_PER_UNIT = 551.0


def calc_ph_roofline(dev, fp64_peak):
    """K2 alone: 262,144 calculate_pH solves of BASELINE configs[3] (ensembles.config4), CUDA events, best of 5."""
    import torch

    from ics_wt_physicsengine_b200 import ensembles
    from ics_wt_physicsengine_b200.chemistry import calculate_pH_batch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    alk, ct, temp, guess = (t(a) for a in ensembles.config4(262144))
    best, iters = None, None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ph, it, st = calculate_pH_batch(alk, ct, temp, guess)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
        iters = float(it.sum())
    # per Newton iteration (chemistry.py:193-269): exp10 (35) + 9 divisions (20 each) + ~40 mul/add = ~255 flops
    flops = 255.0 * iters
    ach = flops / (best * 1e-3) / 1e12
    return {"kernel": "wt_calc_ph_kernel", "bound": "fp64", "iterations_total": iters, "ms_per_launch": best,
            "solves": int(alk.numel()), "flops_per_iteration": 255.0, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach / fp64_peak if fp64_peak else None,
            "note": "persistent warps; a lane that finishes a solve draws the next system from a global queue (iteration counts 1..100)"}


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
    # the floor of this call on this box: the same bytes copied both ways at once (two streams), no kernel
    dev = shard.device
    up_h, dn_h = torch.empty(h2d, dtype=torch.uint8).pin_memory(), torch.empty(d2h, dtype=torch.uint8).pin_memory()
    up_d, dn_d = torch.empty(h2d, dtype=torch.uint8, device=dev), torch.empty(d2h, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    floor = None
    for rep in range(2):
        t0 = time.perf_counter()
        for _ in range(k):
            with torch.cuda.stream(s1):
                up_d.copy_(up_h, non_blocking=True)
            with torch.cuda.stream(s2):
                dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.synchronize()
        floor = time.perf_counter() - t0
    return live, el, k, h2d, d2h, floor


def e2e_sensors_measure(e, lo, dev, args):
    """Step + suite read with the state in pinned HOST memory between steps (PipelinedShard.step_host).  Returns
    (live plants, seconds, steps, h2d bytes/step, d2h bytes/step, sub-ensembles) of this rank's shard."""
    import torch

    from ics_wt_physicsengine_b200 import _lib
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    P = e.n_plants
    parts = max(4, min(8, P // 32768))
    sh = PipelinedShard(e, parts=parts, device=dev, plant0=lo, sensor_seed=20260004, max_attempts=args.max_attempts,
                        sort_every=args.sort_every)
    sh.initialize_sensors(-2000.0)
    for j in range(100):
        for su, s in zip(sh.suites, sh.streams):
            with torch.cuda.stream(s):
                su.read(None, float(j - 100))
    sh.synchronize()
    torch.cuda.synchronize()
    io = sh.alloc_host_io()
    k = max(3, min(args.steps, 10))
    tsim = 0.0

    def one():
        nonlocal tsim
        sh.fork()
        sh.step_host(io, DT, read_time=tsim)
        sh.synchronize()
        torch.cuda.current_stream().synchronize()
        tsim += DT
    for _ in range(3):
        one()
    t0 = time.perf_counter()
    for _ in range(k):
        one()
    el = time.perf_counter() - t0
    live = int(sum(int(((b["status"] & _lib.ST_SKIP_MASK) == 0).sum()) for b in io))
    h2d, d2h = sh.host_io_bytes(io)
    return live, el, k, h2d, d2h, parts


def cpu_baseline(args):
    """The oracle port on the host cores, on the FIRST `cpu_plants` plants of the GPU arm's own ensemble; only
    plant-steps that completed (time advanced) are credited."""
    from ics_wt_physicsengine_b200 import ensembles
    from oracle import wt_oracle as wo

    cores = os.cpu_count() or 1
    sample = args.cpu_plants
    e = ensembles.config5(args.plants, N_ZONES).slice(slice(0, sample))
    par = wo.derive_params(e.cfg, N_ZONES)
    bnd = np.ascontiguousarray(e.bnd)
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1).copy()
    t = np.zeros(sample)
    wo.set_max_attempts(args.max_attempts)
    wo.step_batch(par, bnd, N_ZONES, t, y, dt=DT, nsteps=1, nthreads=cores)
    t_start = t.copy()
    t0 = time.perf_counter()
    wo.step_batch(par, bnd, N_ZONES, t, y, dt=DT, nsteps=args.cpu_steps, nthreads=cores)
    el = time.perf_counter() - t0
    done = float(((t - t_start) / DT).sum())
    return {"value": done * N_ZONES / el, "unit": UNIT, "cores": cores, "kind": "port", "same_plants_as_gpu_arm": True,
            "plant_steps_completed": done, "plant_steps_attempted": float(sample * args.cpu_steps),
            "sample": f"first {sample} plants of the GPU arm's ensemble (config5({args.plants})) x {args.cpu_steps} steps on {cores} "
                      "host threads (oracle C port of reactor.py + scipy Radau; the Python reference itself runs ~1e3 "
                      "zone-steps/s/core, BASELINE.md)"}


if __name__ == "__main__":
    main()
