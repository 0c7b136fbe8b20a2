"""Multi-process path on CPU: world_size-2 gloo all-reduce of the ensemble statistics vector and the
plant partitioner.  The per-shard vectors come from a numpy restatement of the wt_stats kernel
layout (include/wt_b200.h) -- the GPU tests check the kernel against the same function."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ics_wt_physicsengine_b200 import ensembles as ens
from ics_wt_physicsengine_b200.partition import (StatsSpec, finalize_stats, shard_bounds, shard_ensemble, stats_size)


def local_stats_numpy(y_pn, status, n, spec: StatsSpec) -> np.ndarray:
    """y_pn: [P, 3n] species-major.  Same layout as the wt_stats kernel."""
    live = (status & (2 | 128 | 256)) == 0   # not halted, not deferred
    v = np.zeros(stats_size(n))
    v[0], v[1] = live.sum(), (~live).sum()
    yl = y_pn[live]
    ph, cl, T = yl[:, n - 1], yl[:, 2 * n - 1], yl[:, 3 * n - 1]
    v[2] = (cl < spec.chlorine_min).sum()
    v[3] = ((ph < spec.pH_low) | (ph > spec.pH_high)).sum()
    v[4] = (T > spec.temperature_max).sum()
    v[5] = ((status & 512) != 0)[live].sum()            # floor-mode continuations among the live plants
    v[6] = ((status & (128 | 256)) != 0)[~live].sum()   # over budget: being caught up or waiting for it
    sh = np.repeat([spec.shift_pH, spec.shift_chlorine, spec.shift_temperature], n)
    d = yl - sh[None, :]
    v[8::2] = d.sum(axis=0)
    v[9::2] = (d * d).sum(axis=0)
    return v


def sensor_stats_numpy(value_7p, status_7p, fault_7p, plant_status, spec: StatsSpec) -> np.ndarray:
    """Numpy restatement of the wt_sensor_stats layout (include/wt_b200.h): per sensor valid count, shifted sum,
    shifted sum of squares over the finite readings, SensorStatus (12) and SensorFault (7) histograms; live plants."""
    live = (plant_status & (2 | 128 | 256)) == 0
    out = np.zeros((7, 22))
    for s in range(7):
        v = value_7p[s][live]
        ok = np.isfinite(v)
        d = v[ok] - spec.sensor_shifts[s]
        out[s, 0], out[s, 1], out[s, 2] = ok.sum(), d.sum(), (d * d).sum()
        out[s, 3:15] = np.bincount(status_7p[s][live], minlength=12)[:12]
        out[s, 15:22] = np.bincount(fault_7p[s][live], minlength=7)[:7]
    return out.reshape(-1)


def test_shard_bounds_cover_the_ensemble_exactly():
    for total in (1, 7, 64, 1000, 1048576):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_bounds(total, r, world)
                assert 0 <= lo <= hi <= total
                seen.extend(range(lo, hi)) if total <= 1000 else seen.append((lo, hi))
            if total <= 1000:
                assert seen == list(range(total))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == total
                assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _fake_readings(P):
    """Deterministic stand-in for the last suite read of P plants (values with NaNs, status and fault codes)."""
    rng = np.random.default_rng(5)
    val = rng.normal([[7.2], [7.4], [1.9], [1.5], [11.0], [21.0], [22.0]], 0.3, size=(7, P))
    val[rng.random((7, P)) < 0.1] = np.nan
    return val, rng.integers(0, 12, size=(7, P)), rng.integers(0, 7, size=(7, P))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = ens.config2(257, 10)
    e = shard_ensemble(full, rank, world)
    spec = StatsSpec()
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1)
    status = np.zeros(e.n_plants, dtype=np.uint32)
    if rank == 0:
        status[3] = 2    # a halted plant is excluded from the moments and counted as halted
    else:
        status[5] = 512  # a floor-mode continuation: live, counted as degraded
        status[7] = 256  # a plant being caught up: excluded, counted as halted and as pending
    val, st, ft = _fake_readings(full.n_plants)
    lo, hi = shard_bounds(full.n_plants, rank, world)
    v = torch.from_numpy(np.concatenate([local_stats_numpy(y, status, 10, spec),
                                         sensor_stats_numpy(val[:, lo:hi], st[:, lo:hi], ft[:, lo:hi], status, spec)]))
    dist.all_reduce(v, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put(v.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_allreduce_of_statistics_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = ens.config2(257, 10)
    y = np.concatenate([full.pH0, full.Cl0, full.T0], axis=1)
    status = np.zeros(257, dtype=np.uint32)
    status[3] = 2
    lo1 = shard_bounds(257, 1, world)[0]
    status[lo1 + 5], status[lo1 + 7] = 512, 256
    spec = StatsSpec()
    val, st, ft = _fake_readings(257)
    want = np.concatenate([local_stats_numpy(y, status, 10, spec), sensor_stats_numpy(val, st, ft, status, spec)])
    assert got.size == stats_size(10, sensors=True)
    assert np.allclose(got, want, rtol=1e-13, atol=1e-9)
    r = finalize_stats(got, 10, spec)
    lv = (status & (2 | 128 | 256)) == 0
    assert np.array_equal(r["sensor_valid_count"], np.isfinite(val[:, lv]).sum(axis=1))
    assert np.allclose(r["sensor_mean"], np.nanmean(val[:, lv], axis=1), rtol=1e-12)
    assert np.allclose(r["sensor_var"], np.nanvar(val[:, lv], axis=1), rtol=1e-9)
    assert np.allclose(r["sensor_status_hist"].sum(axis=1), 1.0) and np.allclose(r["sensor_fault_hist"].sum(axis=1), 1.0)
    live = lv
    assert r["live"] == 255 and r["halted"] == 2 and r["degraded"] == 1 and r["pending_catch_up"] == 1
    assert np.allclose(r["mean_pH"], full.pH0[live].mean(axis=0), rtol=1e-12)
    assert np.allclose(r["var_temperature"], full.T0[live].var(axis=0), rtol=1e-9)
    assert 0 <= r["frac_outlet_chlorine_low"] <= 1


def _worker_overlap(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ics_wt_physicsengine_b200.partition import OverlappedAllReduce
    local = torch.zeros(222, dtype=torch.float64)
    ar = OverlappedAllReduce(local)
    outs = []
    for blk in range(5):                       # five "blocks": the local vector changes while reductions are in flight
        local.copy_(torch.arange(222, dtype=torch.float64) * (rank + 1) + 1000.0 * blk)
        outs.append((blk, ar.submit(local)))
        if blk >= 1:                           # the buffer of the previous block is not touched by this submission
            assert outs[-1][1].data_ptr() != outs[-2][1].data_ptr()
        local.fill_(-1.0)                      # the producer overwrites its vector right away (the next block's statistics)
    ar.finish()
    got = {blk: t.clone() for blk, t in outs[-2:]}   # the two staging buffers hold the last two blocks
    if rank == 0:
        q.put({k: v.numpy() for k, v in got.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_allreduce_world2():
    """The statistics all-reduce that runs beside the next block (PipelinedShard.replay(overlap=True)): two staging
    buffers, a wait only when a buffer comes round again; every block's reduced vector is the sum over the ranks of the
    vector that was submitted, although the producer overwrites it immediately."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    base = np.arange(222, dtype=np.float64)
    for blk in (3, 4):
        want = base * 1 + 1000.0 * blk + base * 2 + 1000.0 * blk
        assert np.array_equal(got[blk], want), blk
