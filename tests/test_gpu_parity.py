"""Parity tests proper: the CUDA engine, called through the C ABI (include/wt_b200.h) via the
Python facade, against the CPU oracle on the same seeded inputs, against the committed golden
fixtures produced by the reference, and -- at BASELINE.json's full sizes -- through
size-independent properties.  Tolerance: 1e-9 relative per zone variable per step
(BASELINE.json north_star), fp64 throughout."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from ics_wt_physicsengine_b200 import (BoundaryConditions, IntegratedCSTR, PlantEnsemble,  # noqa: E402
                                       ReactorConfiguration, calculate_pH_batch, ensembles as ens)
from tests._util import HALT, check_step_parity, relerr, species_major  # noqa: E402

TOL = 1e-9
CAP = 64


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    from ics_wt_physicsengine_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "CUDA extension not built"
    assert _lib.lib().wt_device_count() > 0


def _stepwise_parity(oracle, e, steps, dt=1.0, cap=CAP):
    """Step the engine and the oracle side by side; before each step the engine is re-seeded
    with the oracle state so the comparison is per step (the north-star statistic)."""
    n, P = e.n_zones, e.n_plants
    eng = PlantEnsemble(e, max_attempts=cap)
    par = np.ascontiguousarray(eng.par_host)
    bnd = np.ascontiguousarray(e.bnd)
    oracle.set_max_attempts(cap)
    yo, to = species_major(e), np.zeros(P)
    halted = np.zeros(P, bool)
    n_excused, worst_ok, path_same, path_tot = 0, 0.0, 0, 0
    worst_excused = [0.0]
    for s in range(steps):
        y_before, t_before = yo.copy(), to.copy()
        eng.set_state(yo[:, :n], yo[:, n:2 * n], yo[:, 2 * n:], time=to)
        eng.reset_status()
        eng.reset_counters()
        eng.step(dt, bnd)
        so, co, fo = oracle.step_batch(par, bnd, n, to, yo, dt=dt, nthreads=os.cpu_count() or 8)
        got = eng.state_numpy()
        sg = eng.status.cpu().numpy().astype(np.uint32)
        cg = eng.counters.cpu().numpy().T
        live = ~halted & ((so & HALT) == 0) & ((sg & HALT) == 0)
        assert abs(int(((so & HALT) != 0).sum()) - int(((sg & HALT) != 0).sum())) <= max(2, P // 2000)
        same = (co[live][:, :7] == cg[live][:, :7]).all(axis=1)
        r, excused = check_step_parity(oracle, got[live], yo[live], par[live], bnd[live], n, t_before[live],
                                       y_before[live], dt, cap, tol=TOL, what=f"step {s}", path_same=same)
        n_excused += len(excused)
        worst_excused[0] = max([worst_excused[0]] + [x[1] for x in excused])
        ok = r <= TOL
        worst_ok = max(worst_ok, float(r[ok].max()) if ok.any() else 0.0)
        path_same += int(same.sum())
        path_tot += int(live.sum())
        assert np.array_equal(eng.state.time.cpu().numpy()[live], to[live])
        assert np.array_equal(eng.state.flow_rate.cpu().numpy()[live], fo[live])
        assert np.array_equal((sg & ~np.uint32(HALT))[live], (so & ~np.uint32(HALT))[live])
        gh = (sg & HALT) != 0
        assert np.array_equal(got[gh], y_before[gh]), "halted plants must be left untouched"
        halted |= gh | ((so & HALT) != 0)
        yo[halted] = y_before[halted]
        to[halted] = t_before[halted]
    assert path_same / path_tot > 0.995, (path_same, path_tot)
    print(f"parity {P} x {n} x {steps}: {path_tot} plant-steps, worst error of a well-conditioned one {worst_ok:.2e}, "
          f"excused (ill-conditioned in the oracle itself) {n_excused} = {n_excused / path_tot:.2e}, largest excused error "
          f"{worst_excused[0]:.2e}, identical solver path {path_same / path_tot:.4f}")
    return n_excused, worst_ok


def test_default_plant_one_hour_against_golden(golden_dir):
    """BASELINE configs[0]: the drop-in IntegratedCSTR against the reference's own trajectory."""
    g = np.load(os.path.join(golden_dir, "config1_default_3600.npz"))
    r = IntegratedCSTR(ReactorConfiguration())
    b = BoundaryConditions()
    worst = 0.0
    for k in range(3):  # 300 of the 3600 steps through the one-plant facade (host round trip per step)
        for _ in range(100):
            s = r.step(1.0, b)
        y = np.concatenate([s.pH, s.chlorine, s.temperature])
        worst = max(worst, float(relerr(y, g["Y"][k][0]).max()))
    assert worst < 1e-12
    assert s.time == 300.0 and s.flow_rate == 5.0


def test_default_plant_full_hour_fused(golden_dir):
    g = np.load(os.path.join(golden_dir, "config1_default_3600.npz"))
    eng = PlantEnsemble(ens.config1(), max_attempts=0)
    for k in range(36):
        eng.advance(100, 1.0, BoundaryConditions())
        assert relerr(eng.state_numpy()[0], g["Y"][k][0]).max() < 1e-12
    c = eng.counters.cpu().numpy()[:, 0]
    assert tuple(c[:4]) == (16 * 3600, 3600, 4 * 3600, 2 * 3600)  # path (16,1,4,2) on every step


@pytest.mark.parametrize("name", ["config2_64x10_25", "config3_48x20_12", "config2_16x10_dt10", "config2_16x5_dt01"])
def test_golden_trajectories_per_step(oracle, golden_dir, name):
    """One engine step from each recorded reference state vs the next recorded reference state."""
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n, dt, P = int(g["n_zones"]), float(g["dt"]), g["cfg"].shape[0]
    e = ens.Ensemble(n, g["cfg"], g["bnd"], g["pH0"], g["Cl0"], g["T0"])
    eng = PlantEnsemble(e, max_attempts=0)
    par = np.ascontiguousarray(eng.par_host)
    y = species_major(e)
    n_excused = 0
    for k in range(int(g["nsteps"])):
        if k > 0:
            y = g["Y"][k - 1].copy()
        eng.set_state(y[:, :n], y[:, n:2 * n], y[:, 2 * n:], time=np.full(P, k * dt))
        eng.reset_status()
        eng.reset_counters()
        eng.step(dt, g["bnd"])
        got = eng.state_numpy()
        assert np.all(eng.status.cpu().numpy() == 0)
        _, excused = check_step_parity(oracle, got, g["Y"][k], par, g["bnd"], n, np.full(P, k * dt), y, dt, 0,
                                       tol=TOL, what=f"{name} step {k}")
        n_excused += len(excused)
        cg = eng.counters.cpu().numpy().T[:, :4]
        assert (cg == g["counters"][k]).all(axis=1).mean() > 0.97
    assert n_excused <= 3


def test_config2_parity_4096x10(oracle):
    """BASELINE configs[1] at full size, 6 steps side by side with the oracle."""
    n_excused, worst = _stepwise_parity(oracle, ens.config2(4096, 10), 6)
    assert worst <= TOL and n_excused <= 8


def test_config3_parity_slice(oracle):
    """BASELINE configs[2] inputs (T sweep 0-100 C, stratified / unstable profiles), 2048 plants."""
    n_excused, worst = _stepwise_parity(oracle, ens.config3(2048, 20), 4)
    assert worst <= TOL and n_excused <= 16


@pytest.mark.parametrize("n", [2, 5, 7, 16, 32])
def test_other_zone_counts(oracle, n):
    n_excused, worst = _stepwise_parity(oracle, ens.config2(300, n, seed=300 + n), 3)
    assert worst <= TOL and n_excused <= 2


def test_derivatives_operator(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "rhs_config3.npz"))
    n, P = int(g["n_zones"]), g["cfg"].shape[0]
    e = ens.Ensemble(n, g["cfg"], g["bnd"], g["Y"][:, :n], g["Y"][:, n:2 * n], g["Y"][:, 2 * n:])
    eng = PlantEnsemble(e, validate=False)
    dy, bad = eng.derivatives(g["bnd"])
    dy = dy.permute(2, 0, 1).reshape(P, 3 * n).cpu().numpy()
    assert not bad.any()
    for v in range(3):
        blk = slice(v * n, (v + 1) * n)
        scale = np.abs(g["F"][:, blk]).max(axis=1, keepdims=True)
        assert (np.abs(dy[:, blk] - g["F"][:, blk]) <= 1e-11 * scale + 1e-300).all()


# ---- size-independent properties at full BASELINE sizes --------------------------------------
def test_fused_advance_equals_repeated_steps_65536x20():
    """configs[2] size: advance(k) must be bit-identical to k step() calls (state stays in
    registers between fused steps; nothing else may change)."""
    e = ens.config3(65536, 20)
    a = PlantEnsemble(e)
    b = PlantEnsemble(e)
    for _ in range(3):
        a.step(1.0, e.bnd)
    b.advance(3, 1.0, e.bnd)
    torch.cuda.synchronize()
    assert torch.equal(a.state.pH, b.state.pH) and torch.equal(a.state.chlorine, b.state.chlorine)
    assert torch.equal(a.state.temperature, b.state.temperature) and torch.equal(a.state.time, b.state.time)
    assert torch.equal(a.status, b.status) and torch.equal(a.counters, b.counters)


def test_plant_order_invariance_262144x10():
    """Plants are independent: permuting the ensemble permutes the result bit for bit (no lane-,
    warp- or block-placement dependence), checked at 262,144 plants."""
    P = 262144
    e = ens.config5(P, 10)
    perm = np.random.default_rng(1).permutation(P)
    a = PlantEnsemble(e)
    b = PlantEnsemble(e.slice(perm))
    a.advance(2, 1.0, e.bnd)
    b.advance(2, 1.0, e.bnd[perm])
    pt = torch.from_numpy(perm).to(a.device)
    assert torch.equal(a.state.pH[pt], b.state.pH) and torch.equal(a.state.chlorine[pt], b.state.chlorine)
    assert torch.equal(a.state.temperature[pt], b.state.temperature)
    assert torch.equal(a.status[pt], b.status)
    st = a.status.cpu().numpy()
    assert ((st & HALT) != 0).mean() < 1e-3
    live = (st & HALT) == 0
    assert np.all(a.state.time.cpu().numpy()[live] == 2.0)


def test_physical_invariants_1m_plants():
    """configs[4] size (1,048,576 x 10), one step: bounds respected, no NaN, batch plants keep
    chlorine monotone (closed, no dosing: total chlorine can only decay)."""
    P = 1048576
    e = ens.config5(P, 10)
    eng = PlantEnsemble(e)
    cl0 = torch.from_numpy(e.Cl0.sum(axis=1)).to(eng.device)
    eng.step(1.0, e.bnd)
    s = eng.state
    assert torch.isfinite(s.pH).all() and torch.isfinite(s.chlorine).all() and torch.isfinite(s.temperature).all()
    assert (s.pH >= 0).all() and (s.pH <= 14).all() and (s.chlorine >= 0).all()
    assert (s.temperature >= 0).all() and (s.temperature <= 100).all()
    no_dose = torch.from_numpy((e.bnd[:, ens.BND_FIELDS.index("chlorine_flow_rate")] == 0)
                               & (e.bnd[:, ens.BND_FIELDS.index("inlet_chlorine")] <= e.cfg[:, ens.CFG_FIELDS.index("initial_chlorine")])).to(eng.device)
    assert (s.chlorine.sum(dim=1)[no_dose] <= cl0[no_dose] * (1 + 1e-12)).all()
    c = eng.counters.cpu().numpy()
    assert c[3].min() >= 0 and c[0].sum() > 16 * P * 0.9


def test_halting_and_status_semantics(oracle):
    e = ens.config1(5)
    e.T0[0] = 100.0
    eng = PlantEnsemble(e, validate=False)
    eng.step(1.0, e.bnd)
    assert int(eng.status[0]) & 2
    assert float(eng.state.time[0]) == 0.0
    assert np.array_equal(eng.state_numpy(), species_major(e))
    with pytest.raises(ValueError):
        r = IntegratedCSTR(ReactorConfiguration())
        r.state.temperature = np.full(5, 100.0)
        r.step(1.0, BoundaryConditions())


# ---- calculate_pH (BASELINE configs[3]) --------------------------------------------------------
def _ph_stable_mask(oracle, alk, ct, temp, guess):
    """Solves whose (iterations, status) survive a 1-2 ulp perturbation of H = 10**(-pH) at every
    iteration IN THE ORACLE ITSELF (wt_oracle_set_ph_h_eps).  About 4 % of the stress inputs
    bounce between the pH clips for 34-100 iterations and are chaotic: no two exp10
    implementations (glibc pow, SVML, CUDA) agree on them."""
    oracle.set_ph_h_eps(0.0)
    _, it, st = oracle.calc_ph_batch(alk, ct, temp, guess, nthreads=8)
    stable = np.ones(alk.size, bool)
    try:
        for eps in (2.3e-16, -2.3e-16, 4.5e-16, -4.5e-16):
            oracle.set_ph_h_eps(eps)
            _, it2, st2 = oracle.calc_ph_batch(alk, ct, temp, guess, nthreads=8)
            stable &= (it2 == it) & (st2 == st)
    finally:
        oracle.set_ph_h_eps(0.0)
    return stable


def test_calculate_ph_golden(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "calc_ph_4096.npz"))
    ph, it, st = calculate_pH_batch(g["alk"], g["ct"], g["temp"], g["guess"])
    ph, it, st = ph.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    stable = _ph_stable_mask(oracle, g["alk"], g["ct"], g["temp"], g["guess"])
    assert stable.mean() > 0.93
    agree = (st == g["status"]) & (it == g["iters"])
    assert agree[g["iters"] <= 30].all(), "short solves are never chaotic"
    assert agree[stable].mean() > 0.995, "well-conditioned solves: same status and iteration count as the reference"
    ok = agree & (st == 0)
    assert relerr(ph[ok], g["ph"][ok]).max() < TOL
    assert set(np.unique(st)) <= {0, 1, 2}


def test_calculate_ph_262144_histogram(oracle):
    alk, ct, temp, guess = ens.config4(262144)
    ph, it, st = calculate_pH_batch(alk, ct, temp, guess)
    ph, it, st = ph.cpu().numpy(), it.cpu().numpy(), st.cpu().numpy()
    pho, ito, sto = oracle.calc_ph_batch(alk, ct, temp, guess, nthreads=8)
    stable = _ph_stable_mask(oracle, alk, ct, temp, guess)
    n_unstable = int((~stable).sum())
    assert n_unstable < 0.06 * alk.size
    agree = (st == sto) & (it == ito)
    assert agree[ito <= 30].all(), "short solves are never chaotic"
    assert agree[stable].mean() > 0.995
    # iteration-count histogram and status counts vs the oracle: differences only from chaotic solves
    hg, ho = np.bincount(it, minlength=101), np.bincount(ito, minlength=101)
    assert np.abs(hg - ho).sum() <= 2 * n_unstable
    assert np.abs(np.bincount(st, minlength=4) - np.bincount(sto, minlength=4)).max() <= n_unstable
    ok = agree & (st == 0)
    assert ok.sum() > 0.8 * alk.size
    assert relerr(ph[ok], pho[ok]).max() < TOL
    # default buffer: 8.39839641036611 in 6 iterations from the grid guess 7.0 (chemistry.py:546-550)
    k = alk.size - 29 + 14
    assert it[k] == 6 and abs(ph[k] - 8.39839641036611) < 1e-12


# ---- ensemble statistics kernel (payload of the NCCL all-reduce) -------------------------------
def test_stats_kernel_matches_numpy_and_is_deterministic():
    from ics_wt_physicsengine_b200.partition import EnsembleStatistics, StatsSpec, finalize_stats
    from tests.test_partition_gloo import local_stats_numpy
    e = ens.config5(200003, 10)  # odd size: ragged last block
    eng = PlantEnsemble(e)
    eng.advance(3, 1.0, e.bnd)
    st = EnsembleStatistics(eng, StatsSpec())
    a = st.local().cpu().numpy().copy()
    b = st.local().cpu().numpy().copy()
    assert np.array_equal(a, b), "fixed summation order: bitwise reproducible"
    want = local_stats_numpy(eng.state_numpy(), eng.status.cpu().numpy().astype(np.uint32), 10, StatsSpec())
    assert np.array_equal(a[:8], want[:8])
    assert np.allclose(a[8:], want[8:], rtol=1e-11, atol=1e-7)
    r = finalize_stats(a, 10, StatsSpec())
    assert r["live"] + r["halted"] == 200003
    assert np.all(r["var_pH"] >= 0) and np.all((r["mean_temperature"] > 0) & (r["mean_temperature"] < 45))


def test_sensor_statistics_kernel_matches_numpy():
    """The sensor half of the all-reduce payload (valid count, sum, sum of squares, status and fault histograms per
    sensor) against its numpy restatement, on a stepped ensemble with warm and warming-up sensors."""
    from ics_wt_physicsengine_b200.partition import EnsembleStatistics, StatsSpec, finalize_stats, stats_size
    from ics_wt_physicsengine_b200.sensors import SENSOR_NAMES, create_realistic_sensor_suite
    from tests.test_partition_gloo import local_stats_numpy, sensor_stats_numpy
    P = 100003
    e = ens.config5(P, 10)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=3)
    suite.initialize(-400.0)   # warm: chlorine, flow, temperature; the pH pair still warms up
    eng._status[5:500:7] = 2   # some halted plants: excluded
    for k in range(40):
        eng.step(1.0, e.bnd)
        r = suite.read(eng.state, float(k))
    st = EnsembleStatistics(eng, StatsSpec(), suite)
    a = st.local().cpu().numpy().copy()
    assert a.size == stats_size(10, sensors=True) and np.array_equal(a, st.local().cpu().numpy())
    ps = eng.status.cpu().numpy().astype(np.uint32)
    val = np.stack([r[k].value.cpu().numpy() for k in SENSOR_NAMES])
    sst = np.stack([r[k].status.cpu().numpy() for k in SENSOR_NAMES])
    sft = np.stack([r[k].fault.cpu().numpy() for k in SENSOR_NAMES])
    want = np.concatenate([local_stats_numpy(eng.state_numpy(), ps, 10, StatsSpec()), sensor_stats_numpy(val, sst, sft, ps, StatsSpec())])
    sb, wb = a[68:].reshape(7, 22), want[68:].reshape(7, 22)
    assert np.array_equal(sb[:, 0], wb[:, 0]) and np.array_equal(sb[:, 3:], wb[:, 3:])   # counts and histograms: exact
    assert np.allclose(sb[:, 1:3], wb[:, 1:3], rtol=1e-11, atol=1e-7)
    f = finalize_stats(a, 10, StatsSpec())
    assert f["sensor_valid_fraction"][0] == 0.0 and f["sensor_valid_fraction"][2] > 0.95   # pH warming up, chlorine reads
    assert f["sensor_status_hist"][0, 2] > 0.99                                            # WARMING_UP (a few power faults)


def test_sorted_scheduling_does_not_change_results():
    """sort_every only permutes which plants share a warp / start first: bitwise identical results."""
    e = ens.config5(30011, 10)
    a = PlantEnsemble(e, sort_every=0)
    b = PlantEnsemble(e, sort_every=1)
    for _ in range(5):
        a.step(1.0, e.bnd)
        b.step(1.0, e.bnd)
    assert torch.equal(a.state.pH, b.state.pH) and torch.equal(a.state.chlorine, b.state.chlorine)
    assert torch.equal(a.state.temperature, b.state.temperature) and torch.equal(a.status, b.status)
    assert torch.equal(a.counters, b.counters) and torch.equal(a.state.time, b.state.time)
    assert b._order is not None and sorted(b._order.cpu().tolist()) == list(range(30011))
    # the counting sort itself: a permutation, most expensive plants first
    import ctypes as C
    from ics_wt_physicsengine_b200 import _lib
    order = torch.empty(30011, dtype=torch.int32, device=b.device)
    pp = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().wt_cost_order(30011, pp(b._cost), pp(order), pp(b._bins), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "wt_cost_order")
    o = order.cpu().numpy().astype(np.int64)
    assert sorted(o.tolist()) == list(range(30011))
    assert np.all(np.diff(np.minimum(b._cost.cpu().numpy()[o], 1023)) <= 0)


@pytest.mark.parametrize("P,bcast", [(33, False), (70001, False), (70001, True), (131072 + 77, False)])
def test_host_buffer_entry_point_equals_the_device_path(P, bcast):
    """wt_step_host (host SoA buffers in, pipelined over column slabs on copy-in / compute / copy-out streams) runs the same kernel
    as the device-resident path: bit-identical state, time, flow and status, for ragged sizes, one and several
    slabs, per-plant and broadcast boundaries, repeated calls with the constants kept resident."""
    import ctypes as C

    from ics_wt_physicsengine_b200 import _lib
    n = 10
    e = ens.config2(P, n, seed=4242)
    eng = PlantEnsemble(e, max_attempts=CAP)
    bnd_rows = np.ascontiguousarray(e.bnd[0]) if bcast else np.ascontiguousarray(e.bnd)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    par = pin(eng.par_host.T)
    bnd = pin(bnd_rows if bcast else bnd_rows.T)
    y = pin(np.stack([e.pH0.T, e.Cl0.T, e.T0.T]))
    t = torch.zeros(P, dtype=torch.float64).pin_memory()
    flow = torch.zeros(P, dtype=torch.float64).pin_memory()
    st = torch.zeros(P, dtype=torch.int32).pin_memory()
    p = lambda x: C.c_void_p(x.data_ptr())
    for k in range(3):
        rc = _lib.lib().wt_step_host(P, n, 1.0, p(par), p(bnd), 0 if bcast else P, p(t), p(y), p(flow), p(st), CAP,
                                     1 if k else 0)
        _lib.check(rc, "wt_step_host")
        eng.step(1.0, bnd_rows)
        torch.cuda.synchronize()
        assert np.array_equal(y.numpy().reshape(3 * n, P).T, eng.state_numpy())
        assert np.array_equal(t.numpy(), eng.state.time.cpu().numpy())
        assert np.array_equal(st.numpy(), eng.status.cpu().numpy())
        live = (st.numpy() & HALT) == 0
        assert np.array_equal(flow.numpy()[live], eng.state.flow_rate.cpu().numpy()[live])


def test_pipelined_shard_equals_one_ensemble():
    """A shard run as independent sub-ensembles on their own CUDA streams (partition.PipelinedShard) gives
    bit-identical plant states, sensor readings (noise is keyed by the global plant id) and statistics."""
    from ics_wt_physicsengine_b200.partition import EnsembleStatistics, PipelinedShard
    from ics_wt_physicsengine_b200.sensors import create_realistic_sensor_suite
    P, n = 3001, 10
    e = ens.config2(P, n, seed=99)
    one = PlantEnsemble(e, max_attempts=CAP)
    suite = create_realistic_sensor_suite(one, seed=7, plant0=1000)
    suite.initialize(0.0)
    st = EnsembleStatistics(one)
    sh = PipelinedShard(e, parts=3, plant0=1000, sensor_seed=7, max_attempts=CAP)
    sh.initialize_sensors(0.0)
    for k in range(5):
        one.step(1.0, e.bnd)
        suite.read(one.state, float(k))
        sh.step(1.0, read_time=float(k))
    v = sh.stats().clone()
    torch.cuda.synchronize()
    st = EnsembleStatistics(one, None, suite)
    got = np.concatenate([x.state_numpy() for x in sh.engines])
    assert np.array_equal(got, one.state_numpy())
    assert np.array_equal(np.concatenate([x.status.cpu().numpy() for x in sh.engines]), one.status.cpu().numpy())
    assert np.array_equal(torch.cat([s._out for s in sh.suites], dim=2).cpu().numpy(), suite._out.cpu().numpy(),
                          equal_nan=True), "sensor readings differ"  # warming-up sensors read NaN
    assert torch.equal(torch.cat([s._out_status for s in sh.suites], dim=1), suite._out_status)
    w = st.local()
    assert torch.allclose(v, w, rtol=1e-13, atol=1e-9)  # additive vector; only the summation order differs


def test_captured_graph_replay_equals_eager_steps():
    """One CUDA graph per rank for a block of steps (step + sensor read + cost order of every sub-ensemble + local
    statistics): replays give bit-identical plant states, sensor readings and statistics to the same steps
    launched eagerly."""
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    P, n, G = 6007, 10, 4
    e = ens.config5(P, n)
    mk = lambda: PipelinedShard(e, parts=2, plant0=500, sensor_seed=11, max_attempts=CAP, sort_every=2)
    a, b = mk(), mk()
    for sh in (a, b):
        sh.initialize_sensors(-100.0)
    k = 0
    for _ in range(2):          # warm-up, eager on both (a captured launch cannot set kernel attributes)
        a.step(1.0, read_time=float(k)); b.step(1.0, read_time=float(k)); k += 1
    b.capture(G, 1.0, t_next=float(k), with_stats=True)
    for rep in range(3):
        for _ in range(G):
            a.step(1.0, read_time=float(k)); k += 1
        va = a.stats().clone()
        vb = b.replay().clone()
        torch.cuda.synchronize()
        for ea, eb in zip(a.engines, b.engines):
            assert torch.equal(ea._y, eb._y) and torch.equal(ea._status, eb._status) and torch.equal(ea._time, eb._time)
            assert torch.equal(ea._counters, eb._counters)
        for sa, sb in zip(a.suites, b.suites):
            assert torch.equal(torch.nan_to_num(sa._out, nan=-1.0), torch.nan_to_num(sb._out, nan=-1.0))
            assert torch.equal(sa._out_status, sb._out_status) and sa.read_index == sb.read_index and sa.last_time == sb.last_time
        assert torch.equal(va, vb), rep


def test_derived_state_matches_oracle_and_reference(oracle, golden_dir):
    """a14 _update_derived_state (reactor.py:511-524): the kernel-written H_concentration / density /
    chlorine_decay_rate (i) against the reference's own ReactorState fields after each of 4 steps of 64 config-3
    plants (tests/golden/derived_config3.npz) and (ii) against the oracle on 4,096 config-3 plants incl. the
    <= 8 C density branch."""
    g = np.load(os.path.join(golden_dir, "derived_config3.npz"))
    n, P = int(g["n_zones"]), g["cfg"].shape[0]
    e = ens.Ensemble(n, g["cfg"], g["bnd"], g["pH0"], g["Cl0"], g["T0"])
    eng = PlantEnsemble(e, max_attempts=0)
    for s in range(int(g["nsteps"])):
        eng.step(1.0, g["bnd"])
        d = eng._derived.permute(2, 0, 1).reshape(P, 3 * n).cpu().numpy()
        assert relerr(eng.state_numpy(), g["Y"][s]).max() < TOL
        assert relerr(d, g["D"][s]).max() < TOL, s
    # (ii) kernel vs oracle, one step from identical states
    e = ens.config3(4096, 20, seed=77)
    eng = PlantEnsemble(e, max_attempts=CAP)
    par = np.ascontiguousarray(eng.par_host)
    oracle.set_max_attempts(CAP)
    eng.step(1.0, e.bnd)
    got_d = eng._derived.permute(2, 0, 1).reshape(4096, 60).cpu().numpy()
    st = eng.status.cpu().numpy()
    y0 = species_major(e)
    n_cold, worst = 0, 0.0
    for p in range(0, 4096, 8):
        if st[p] & (HALT | 64):
            continue
        y, t, fl, d, cnt = y0[p].copy(), np.zeros(1), np.zeros(1), np.zeros(60), np.zeros(8, np.int32)
        so = oracle.lib().wt_oracle_step(oracle._dp(par[p]), oracle._dp(np.ascontiguousarray(e.bnd[p])), 20, 1.0, oracle._dp(t),
                                         oracle._dp(y), oracle._dp(fl), oracle._dp(d), oracle._ip(cnt))
        if so & (HALT | 64):
            continue
        worst = max(worst, float(relerr(got_d[p], d).max()))
        n_cold += int((y[40:] <= 8.0).any())
    assert worst < TOL and n_cold >= 10


def test_config3_full_size_20_steps(oracle):
    """BASELINE configs[2] at FULL size: 65,536 plants x 20 zones (temperature sweep, stratified and unstable
    profiles), 20 steps side by side with the multithreaded oracle, per-step parity at 1e-9."""
    n_excused, worst = _stepwise_parity(oracle, ens.config3(65536, 20), 20)
    print(f"config3 65536x20x20: worst non-excused error {worst:.2e}, excused plant-steps {n_excused} of {65536 * 20}")
    assert worst <= TOL and n_excused <= 65536 * 20 // 2000


def test_config2_600_steps_on_a_slice(oracle):
    """BASELINE configs[1] for its full LENGTH (600 steps, dt = 1 s) on a 1,024-plant slice, per-step parity."""
    n_excused, worst = _stepwise_parity(oracle, ens.config2(4096, 10).slice(slice(0, 1024)), 600)
    print(f"config2 1024x10x600: worst non-excused error {worst:.2e}, excused plant-steps {n_excused} of {1024 * 600}")
    assert worst <= TOL and n_excused <= 1024 * 600 // 2000


def test_two_devices_in_one_process():
    """The step kernels need > 48 KB of dynamic shared memory; that attribute belongs to the (kernel, device) pair.
    Round 1 set it once per process, so the first launch on a second GPU of the same process failed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    e = ens.config2(999, 10, seed=5)
    a = PlantEnsemble(e, device="cuda:0", max_attempts=CAP)
    b = PlantEnsemble(e, device="cuda:1", max_attempts=CAP)
    for _ in range(3):
        a.step(1.0, e.bnd)
        b.step(1.0, e.bnd)
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert np.array_equal(a.state_numpy(), b.state_numpy()) and np.array_equal(a.status.cpu().numpy(), b.status.cpu().numpy())


def test_deferred_plants_are_continued_and_equal_a_run_with_the_larger_budget(oracle):
    """Budget-exhausted plants are not dropped: with a deliberately tiny main budget (8 collocation solves per
    plant-step, so that thousands of plant-steps overrun it) and a catch-up budget of 2,048 on the side stream, every
    plant that rejoins carries exactly the state of a run in which ALL plants had the larger budget; a plant that needs
    65..2,000 collocation solves for one step -- halted for good in round 1 -- equals the oracle without any budget."""
    from ics_wt_physicsengine_b200 import _lib
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    P, n, B, blocks, BIG = 40009, 10, 5, 4, 2048
    e = ens.config5(P, n)
    ref = PipelinedShard(e, parts=2, max_attempts=BIG)
    dfr = PipelinedShard(e, parts=2, max_attempts=8, catch_up_attempts=BIG)
    dfr.start_deferral(0.0)
    n_deferred_seen = 0
    for b in range(blocks):
        for _ in range(B):
            ref.step(1.0)
        dfr.block(B, 1.0, t_first=float(b * B))
        n_deferred_seen += dfr.halted()   # over budget in this block: collected and caught up during the next one
    dfr.block(B, 1.0, t_first=float(blocks * B))       # one more block: the plants deferred in the last one rejoin
    for _ in range(B):
        ref.step(1.0)
    torch.cuda.synchronize()
    yr = np.concatenate([x.state_numpy() for x in ref.engines]); yd = np.concatenate([x.state_numpy() for x in dfr.engines])
    tr = np.concatenate([x.state.time.cpu().numpy() for x in ref.engines]); td = np.concatenate([x.state.time.cpu().numpy() for x in dfr.engines])
    sr = np.concatenate([x.status.cpu().numpy() for x in ref.engines]); sd = np.concatenate([x.status.cpu().numpy() for x in dfr.engines])
    assert n_deferred_seen > 100, "the tiny budget must actually defer plants"
    settled = ((sd & _lib.ST_SKIP_MASK) == 0) & ((sr & _lib.ST_HALT_MASK) == 0)   # not deferred in the very last block, not halted
    assert settled.mean() > 0.9   # the rest ran over the tiny budget in the very last block (collected by the next one)
    assert np.array_equal(td[settled], tr[settled]) and np.all(tr[settled] == (blocks + 1) * B)
    assert np.array_equal(yd[settled], yr[settled]), "a continued plant must equal the run with the larger budget bit for bit"
    # plants halted in the reference run (monsters beyond 2,048 solves) are halted here too, nothing else is
    print(f"deferral: {n_deferred_seen} plant-steps over the budget of 8 were continued; halted for good in both runs: "
          f"{int(((sr & _lib.ST_WORK_LIMIT) != 0).sum())} / {int(((sd & _lib.ST_WORK_LIMIT) != 0).sum())}")

    # a step that needs 65 .. 2,000 collocation solves, against the oracle WITHOUT a budget
    e2 = ens.config5(131072, n)
    eng = PlantEnsemble(e2, max_attempts=64)
    found = None
    for k in range(40):
        y_before, t_before = eng.state_numpy(), eng.state.time.cpu().numpy().copy()
        eng.step(1.0, e2.bnd)
        hit = np.nonzero(eng.status.cpu().numpy() & _lib.ST_WORK_LIMIT)[0]
        for p in hit:
            one = PlantEnsemble(e2.slice(slice(int(p), int(p) + 1)), max_attempts=2000)
            one.set_state(y_before[p:p + 1, :n], y_before[p:p + 1, n:2 * n], y_before[p:p + 1, 2 * n:], time=t_before[p:p + 1])
            one.step(1.0, e2.bnd[p:p + 1])
            c = one.counters.cpu().numpy()[:, 0]
            if int(one.status[0]) & _lib.ST_WORK_LIMIT == 0 and c[3] + c[5] + c[6] > 64:
                found = (int(p), y_before[p].copy(), float(t_before[p]), one.state_numpy()[0], int(c[3] + c[5] + c[6]))
                break
        if found:
            break
    if found is None:
        pytest.skip("no plant-step with 65..2,000 collocation solves in this sample")
    p, y0, t0, got, attempts = found
    oracle.set_max_attempts(0)
    par = np.ascontiguousarray(eng.par_host[p:p + 1])
    yo, to = y0[None, :].copy(), np.array([t0])
    oracle.step_batch(par, np.ascontiguousarray(e2.bnd[p:p + 1]), n, to, yo, dt=1.0)
    # Such a step is chaotic in the reference itself (tens of rejected attempts on the 8 C density discontinuity): a few
    # ulp on its input move the oracle's own result by s, typically 1e-3 .. 1e-1.  The continued plant must agree with
    # the unbudgeted oracle to 1e-9, or within 10 s where the oracle is not reproducible.
    from tests._util import oracle_sensitivity
    r = float(relerr(got[None, :], yo).max())
    sens = oracle_sensitivity(oracle, par[0], np.ascontiguousarray(e2.bnd[p]), n, t0, y0, 1.0, 0, seed=p, trials=16, max_ulps=8)
    print(f"plant {p}: {attempts} collocation solves for one step; error vs the unbudgeted oracle {r:.2e}; the oracle's own "
          f"sensitivity to 1-8 ulp of input there {sens:.2e}")
    assert r <= TOL or r <= 10.0 * sens


def test_floor_mode_catch_up_equals_the_oracle_and_nobody_stays_halted(oracle, golden_dir):
    """Floor mode (engine policy).  (1) The plant-steps of tests/golden/overrun_plants.npz -- config-5
    plants on the 8 C density discontinuity that exhaust the budget of 64 -- go through collect / catch-up / rejoin with
    catch_up_floor_div = 16 and come back with the state of the oracle's mirror of the policy, WT_ST_DEGRADED set, time
    advanced, nobody halted.  (2) A config-5 ensemble run block-wise with floor-mode deferral loses no plant to the
    budget (the run without it does)."""
    from ics_wt_physicsengine_b200 import _lib
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    g = np.load(os.path.join(golden_dir, "overrun_plants.npz"))
    n, K = int(g["n_zones"]), len(g["plant"])
    e = ens.config5(65536, n).slice(g["plant"])
    y0 = g["y0"]
    eng = PlantEnsemble(e, max_attempts=64, catch_up_attempts=128, catch_up_floor_div=16)
    eng.set_state(y0[:, :n], y0[:, n:2 * n], y0[:, 2 * n:], time=np.zeros(K))
    eng.step(1.0, e.bnd)
    torch.cuda.synchronize()
    assert np.all(eng.status.cpu().numpy() & _lib.ST_WORK_LIMIT), "these plant-steps overrun the budget of 64"
    assert np.array_equal(eng.state_numpy(), y0) and np.all(eng.state.time.cpu().numpy() == 0.0)
    eng.reset_counters()   # (the path counters of the abandoned attempt are kept by the engine; the oracle's start here)
    eng._t_stop.fill_(1.0)
    eng.collect_deferred()
    assert np.all(eng.status.cpu().numpy() & _lib.ST_DEFERRED)
    eng.catch_up(2, 1.0, e.bnd)   # the second launch finds every plant at the stop time
    eng.rejoin_deferred()
    torch.cuda.synchronize()
    yo, to = y0.copy(), np.zeros(K)
    oracle.set_max_attempts(128)
    oracle.set_floor_div(16)
    try:
        so, co, _ = oracle.step_batch(np.ascontiguousarray(g["par"]), np.ascontiguousarray(g["bnd"]), n, to, yo, dt=1.0)
    finally:
        oracle.set_floor_div(0)
        oracle.set_max_attempts(0)
    st = eng.status.cpu().numpy()
    rel = np.abs(eng.state_numpy() - yo) / np.maximum(np.abs(yo), 1e-300)
    # These plants sit ON a discontinuity of the RHS: a last-bit difference (FMA contraction, exp) flips a Richardson
    # switch in some stage evaluation and moves the result by O(jump x step).  The kernel SOURCE equals the oracle's
    # mirror of the policy bit for bit (tests/test_floor_mode.py, lane-emulation build without contraction); the
    # compiled kernel must equal it on the plant-steps whose switches do not flip, and stay inside the spread of the
    # policy itself (floor dt/4 .. dt/256 agree to 5e-5; bound 2e-3) on the others.
    worst = rel.max(axis=1)
    cnt = eng.counters.cpu().numpy().T[:, :7]
    same_path = (cnt == co[:, :7]).all(axis=1)
    print(f"floor mode: {K} overrun plant-steps vs the oracle's floor mode: {int((worst < 1e-9).sum())} within 1e-9, "
          f"{int(same_path.sum())} on the same solver path, worst {worst.max():.2e}, median {np.median(worst):.2e}")
    assert (worst < 1e-9).sum() >= K - 3 and worst.max() < 2e-3   # measured on B200: 21 of 21 within 1e-9 (worst 3e-14)
    assert np.all(worst[same_path] < 1e-6)
    assert np.all(st & _lib.ST_DEGRADED) and not np.any(st & _lib.ST_SKIP_MASK), (st, so, cnt[:, 3] + cnt[:, 5] + cnt[:, 6])
    assert np.array_equal(st[worst < 1e-9], so[worst < 1e-9])
    assert np.all(eng.state.time.cpu().numpy() == 1.0)
    # the next ordinary step takes them again (they are neither halted nor deferred)
    eng.step(1.0, e.bnd)
    torch.cuda.synchronize()
    assert np.all((eng.state.time.cpu().numpy() == 2.0) | ((eng.status.cpu().numpy() & _lib.ST_WORK_LIMIT) != 0))

    # (2) block-wise, with sensors off: budget 64, floor-mode catch-up vs plain halting
    P, B, blocks = 65536, 5, 8
    e2 = ens.config5(P, n)
    plain = PipelinedShard(e2, parts=2, max_attempts=64)
    floor = PipelinedShard(e2, parts=2, max_attempts=64, catch_up_attempts=128, catch_up_floor_div=16)
    floor.start_deferral(0.0)
    for b in range(blocks):
        for _ in range(B):
            plain.step(1.0)
        floor.block(B, 1.0, t_first=float(b * B))
    floor.block(B, 1.0, t_first=float(blocks * B))   # the plants deferred in the last block rejoin
    torch.cuda.synchronize()
    sd = np.concatenate([x.status.cpu().numpy() for x in floor.engines])
    td = np.concatenate([x.state.time.cpu().numpy() for x in floor.engines])
    lost_plain = plain.halted()
    print(f"floor mode: {lost_plain} of {P} plants halted after {blocks * B} steps without deferral; with floor-mode deferral "
          f"{int(((sd & _lib.ST_T_RANGE) != 0).sum())} (temperature range) + {int(((sd & _lib.ST_WORK_LIMIT) != 0).sum())} (over budget "
          f"in the very last block, collected by the next one); degraded plant-steps in the last block: {int(((sd & _lib.ST_DEGRADED) != 0).sum())}")
    assert lost_plain >= 10
    settled = (sd & _lib.ST_SKIP_MASK) == 0
    assert settled.mean() > 0.999 and np.all(td[settled] == (blocks + 1) * B)


def test_step_host_of_a_shard_equals_the_device_resident_steps():
    """PipelinedShard.step_host (state, time and boundary rows from pinned host buffers, state / time / flow / status and
    the suite's readings back to them, every step) gives bit for bit what the device-resident steps give."""
    from ics_wt_physicsengine_b200.partition import PipelinedShard
    P, n = 5003, 10
    e = ens.config5(P, n)
    mk = lambda: PipelinedShard(e, parts=3, plant0=77, sensor_seed=13, max_attempts=CAP, sort_every=1)
    a, b = mk(), mk()
    for sh in (a, b):
        sh.initialize_sensors(-50.0)
    io = b.alloc_host_io()
    h2d, d2h = b.host_io_bytes(io)
    assert h2d == P * (3 * n + 1 + 10) * 8 and d2h == P * ((3 * n + 2) * 8 + 4 + 7 * 5 * 8 + 2 * 7 * 4)
    for k in range(4):
        a.step(1.0, read_time=float(k))
        b.fork()
        b.step_host(io, 1.0, read_time=float(k))
        b.synchronize()
        torch.cuda.synchronize()
        for i, (ea, sa) in enumerate(zip(a.engines, a.suites)):
            assert torch.equal(io[i]["y"], ea._y.cpu()) and torch.equal(io[i]["time"], ea._time.cpu())
            assert torch.equal(io[i]["status"], ea._status.cpu()) and torch.equal(io[i]["flow"], ea._flow.cpu())
            assert torch.equal(torch.nan_to_num(io[i]["sensor"], nan=-1.0), torch.nan_to_num(sa._out.cpu(), nan=-1.0))
            assert torch.equal(io[i]["sensor_status"], sa._out_status.cpu()) and torch.equal(io[i]["sensor_fault"], sa._out_fault.cpu())
    # the host owns the state: what it writes into the buffers is what the next step starts from
    io[0]["y"][2] += 1.0   # + 1 K in every zone of the first sub-ensemble
    b.fork(); b.step_host(io, 1.0, read_time=4.0); b.synchronize(); torch.cuda.synchronize()
    a.step(1.0, read_time=4.0); torch.cuda.synchronize()
    assert not torch.equal(io[0]["y"], a.engines[0]._y.cpu()) and torch.equal(io[1]["y"], a.engines[1]._y.cpu())
