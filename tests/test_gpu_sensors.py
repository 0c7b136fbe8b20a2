"""Sensor-suite kernel (K3) through the facade / C ABI: value-for-value against the CPU port (both
draw from the same counter-based Philox stream), in distribution against the committed samples of
the unmodified reference, plus the API behaviour of the reference (calibrate, monotonic time)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles as ens  # noqa: E402
from ics_wt_physicsengine_b200.sensors import SENSOR_NAMES, SensorStatus, create_realistic_sensor_suite  # noqa: E402
from tests.test_sensors_oracle import compare_distributions  # noqa: E402


def _gpu_read(suite, t):
    r = suite.read(None, t)
    val = torch.stack([r[k].value for k in SENSOR_NAMES]).cpu().numpy()
    raw = torch.stack([r[k].raw_value for k in SENSOR_NAMES]).cpu().numpy()
    noise = torch.stack([r[k].noise for k in SENSOR_NAMES]).cpu().numpy()
    drift = torch.stack([r[k].drift for k in SENSOR_NAMES]).cpu().numpy()
    unc = torch.stack([r[k].uncertainty for k in SENSOR_NAMES]).cpu().numpy()
    st = torch.stack([r[k].status for k in SENSOR_NAMES]).cpu().numpy()
    ft = torch.stack([r[k].fault for k in SENSOR_NAMES]).cpu().numpy()
    return np.stack([val, raw, noise, drift, unc], axis=2), st, ft   # [7, P, 5]


def _close(a, b, tol=1e-9):
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-6)
    d[both_nan] = 0.0
    return np.nan_to_num(d, nan=np.inf).max() <= tol


def test_default_plant_1900_reads_value_for_value(oracle, golden_dir):
    """All warm-ups (10/30/60/300/1800 s), the shared 30 s delay lines and the absorbing faults, on the
    golden default-plant trajectory, 2,048 suites: every field of every reading equals the CPU port."""
    g = np.load(os.path.join(golden_dir, "sensors_default_plant.npz"))
    P, n, t0 = 2048, 5, float(g["t0"])
    e = ens.config1(5)
    e = ens.Ensemble(5, np.repeat(e.cfg, P, 0), np.repeat(e.bnd, P, 0), np.repeat(e.pH0, P, 0), np.repeat(e.Cl0, P, 0), np.repeat(e.T0, P, 0))
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=777)
    suite.initialize(t0)
    osu = oracle.SensorSuiteOracle(e.cfg[:, 3], e.cfg[:, 12], e.cfg[:, 13], t0, seed=777, nthreads=8)
    mism = 0
    for k in range(1901):
        y = np.concatenate([g["traj_pH"][k], g["traj_Cl"][k], g["traj_T"][k]])
        Y = np.broadcast_to(y, (P, 15))
        eng.set_state(Y[:, :5], Y[:, 5:10], Y[:, 10:])
        eng._flow.fill_(float(g["traj_flow"][k]))
        out_g, st_g, ft_g = _gpu_read(suite, t0 + k)
        out_o, st_o, ft_o = osu.read(Y, np.full(P, float(g["traj_flow"][k])), t0 + k, n)
        same = (st_g.T == st_o) & (ft_g.T == ft_o)
        mism += int((~same).sum())
        if k % 50 == 0 or k in (1800, 1801, 1805, 1830, 1840):
            assert _close(np.transpose(out_g, (1, 0, 2)), out_o), k
    assert mism == 0


def test_random_plants_with_stepping_value_for_value(oracle):
    """Sensors reading an ensemble that is actually being stepped (config-2 plants, n = 10)."""
    P, n, t0 = 1536, 10, 50.0
    e = ens.config2(P, n, seed=99)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=5, plant0=1000)
    suite.initialize(t0)
    osu = oracle.SensorSuiteOracle(e.cfg[:, 3], e.cfg[:, 12], e.cfg[:, 13], t0, seed=5, plant0=1000, nthreads=8)
    for k in range(340):
        eng.step(1.0, e.bnd)
        out_g, st_g, ft_g = _gpu_read(suite, t0 + k)
        out_o, st_o, ft_o = osu.read(eng.state_numpy(), eng.state.flow_rate.cpu().numpy(), t0 + k, n)
        assert np.array_equal(st_g.T, st_o) and np.array_equal(ft_g.T, ft_o), k
        if k % 20 == 0 or k > 300:
            assert _close(np.transpose(out_g, (1, 0, 2)), out_o), k


def test_distribution_against_reference_samples(golden_dir):
    """10,240 engine suites vs 10,240 reference suites at the recorded check times (moments + KS)."""
    g = np.load(os.path.join(golden_dir, "sensors_default_plant.npz"))
    checks = list(g["checks"])
    P, t0 = 10240, float(g["t0"])
    e = ens.config1(5)
    e = ens.Ensemble(5, np.repeat(e.cfg, P, 0), np.repeat(e.bnd, P, 0), np.repeat(e.pH0, P, 0), np.repeat(e.Cl0, P, 0), np.repeat(e.T0, P, 0))
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=20260004)
    suite.initialize(t0)
    n_tested, ci = 0, 0
    for k in range(max(checks) + 1):
        y = np.concatenate([g["traj_pH"][k], g["traj_Cl"][k], g["traj_T"][k]])
        Y = np.broadcast_to(y, (P, 15))
        eng.set_state(Y[:, :5], Y[:, 5:10], Y[:, 10:])
        eng._flow.fill_(float(g["traj_flow"][k]))
        r = suite.read(None, t0 + k)
        if ci < len(checks) and k == checks[ci]:
            for s, name in enumerate(SENSOR_NAMES):
                b = r[name].value.cpu().numpy()
                a = g["values"][ci, s]
                pa, pb = np.isnan(a).mean(), np.isnan(b).mean()
                assert abs(pa - pb) < 5 * np.sqrt(max(pa * (1 - pa), 1e-4) * 2 / P) + 1e-3, (k, name, pa, pb)
                ha = np.bincount(g["status"][ci, s].astype(int), minlength=12) / a.size
                hb = np.bincount(r[name].status.cpu().numpy(), minlength=12) / P
                assert np.abs(ha - hb).max() < 0.02, (k, name)
                a, b = a[np.isfinite(a)], b[np.isfinite(b)]
                if a.size >= 500:
                    n_tested += bool(compare_distributions(a, b, (k, name)))
            ci += 1
    assert n_tested >= 20


def test_results_do_not_depend_on_sharding():
    """Counter-based RNG keyed by the GLOBAL plant id: a shard reproduces its slice of the full run."""
    P, n, t0 = 512, 10, 0.0
    e = ens.config2(P, n, seed=3)
    full = PlantEnsemble(e)
    sf = create_realistic_sensor_suite(full, seed=42)
    sf.initialize(t0)
    lo, hi = 200, 456
    part = PlantEnsemble(e.slice(slice(lo, hi)))
    sp = create_realistic_sensor_suite(part, seed=42, plant0=lo)
    sp.initialize(t0)
    for k in range(70):
        full.step(1.0, e.bnd)
        part.step(1.0, e.bnd[lo:hi])
        a, sa, fa = _gpu_read(sf, t0 + k)
        b, sb, fb = _gpu_read(sp, t0 + k)
        assert np.array_equal(np.nan_to_num(a[:, lo:hi], nan=-1e300), np.nan_to_num(b, nan=-1e300))
        assert np.array_equal(sa[:, lo:hi], sb) and np.array_equal(fa[:, lo:hi], fb)


def test_calibrate_and_time_semantics():
    e = ens.config2(64, 10, seed=4)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng)
    with pytest.raises(RuntimeError):
        suite.read(None, 0.0)
    suite.initialize(0.0)
    for k in range(15):
        eng.step(1.0, e.bnd)
        r = suite.read(eng.state, float(k))
    assert r["flow_main"].status_of(0) in (SensorStatus.DRIFT_WARNING, SensorStatus.POWER_FAULT, SensorStatus.SATURATED)
    with pytest.raises(ValueError, match="Non-monotonic"):
        suite.read(eng.state, 3.0)                           # base_sensor.py:543-549
    # calibrate(reference, t): offset = reference - current_value and the warm-up timer restarts
    cur = suite._sens[0, 4].clone()
    suite.calibrate("flow_main", 5.0, 15.0)
    assert torch.allclose(suite._sens[2, 4], 5.0 - cur)
    r = suite.read(eng.state, 16.0)
    assert (r["flow_main"].status.cpu().numpy() == 2).mean() > 0.9    # WARMING_UP again for 10 s
    assert set(suite.keys()) == set(SENSOR_NAMES)


def test_orchestrator_closed_loop_on_device():
    """step -> sensors -> controller -> clamps -> boundary, no host round trip of the state."""
    from ics_wt_physicsengine_b200.orchestrator import EnsembleOrchestrator
    P, n = 4096, 10
    e = ens.config2(P, n, seed=21)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=1)
    bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).to(eng.device)
    orch = EnsembleOrchestrator(eng, suite, bnd, t0=0.0)

    def controller(readings, state, k):
        # dose chlorine where the outlet reading (once the sensor is warm) is below 1 mg/L; garbage on purpose elsewhere
        cl = readings["chlorine_outlet"].value
        want = torch.where(torch.isnan(cl), torch.full_like(cl, float("nan")), (1.0 - cl) * 5.0)
        return torch.zeros_like(cl), want, torch.full_like(cl, 25.0)   # inlet command above the 20 L/min clamp

    orch.run(80, 1.0, controller)
    assert float(bnd[0].max()) == 20.0 and float(bnd[0].min()) == 20.0          # inlet clamped to 20
    assert float(bnd[6].min()) >= 0.0 and float(bnd[6].max()) <= 1.0            # chlorine dosing in [0, 1]
    assert not torch.isnan(bnd).any()
    assert np.all(eng.state.time.cpu().numpy()[(eng.status.cpu().numpy() & 130) == 0] == 80.0)


def test_register_image_matches_the_wire_oracle():
    """K6 (wt_register_image): the Modbus input-register image of selected plants, bit for bit against the
    restatement of update_modbus_inputs + ModbusEncoder (oracle/wt_wire_oracle.py, pinned against the
    reference's own encoder), including NaN readings of warming-up sensors, faults and the +-1e9 rejection."""
    from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles as ens
    from ics_wt_physicsengine_b200.sensors import create_realistic_sensor_suite
    from oracle import wt_wire_oracle as ww
    P, n = 5000, 10
    e = ens.config2(P, n, seed=12)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=5)
    suite.initialize(0.0)
    for k in range(45):
        eng.step(1.0, e.bnd)
        suite.read(eng.state, float(k))
    # poke special cases into the last read's outputs
    suite._out[0, 2, 17] = float("inf")
    suite._out[0, 5, 18] = 2.5e9        # the reference's update raises on this row
    suite._out_fault[3, 19] = 3
    sel = np.concatenate([np.arange(0, 64), [17, 18, 19, P - 1], np.random.default_rng(0).integers(0, P, 300)]).astype(np.int32)
    ir, di, ok = suite.register_image(sel, 44.0)
    torch.cuda.synchronize()
    vals = suite._out[0].cpu().numpy().T[sel]
    flt = suite._out_fault.cpu().numpy().T[sel]
    want_ir, want_di, want_ok = ww.register_image(vals, flt, 44.0)
    assert np.array_equal(ir.cpu().numpy().view(np.uint16), want_ir)
    assert np.array_equal(di.cpu().numpy(), want_di)
    assert np.array_equal(ok.cpu().numpy().astype(bool), want_ok)
    assert not want_ok[65] and want_ok[64] and np.isnan(vals).any()   # the special cases are really in the sample


def test_maintenance_operations_value_for_value(oracle):
    """calibrate_two_point / clean_electrode / replace_membrane / replace_reagent applied between reads: every
    later reading of the kernel equals the CPU port (itself pinned against the reference's own methods in
    tests/test_sensors_oracle.py), and the reference's ValueError cases raise."""
    P, n, t0 = 1024, 10, 0.0
    e = ens.config2(P, n, seed=7)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=11, plant0=5)
    suite.initialize(t0)
    osu = oracle.SensorSuiteOracle(e.cfg[:, 3], e.cfg[:, 12], e.cfg[:, 13], t0, seed=11, plant0=5, nthreads=8)
    ops = {400: [("clean_electrode", ("pH_inlet", "acid_clean"), (0, 1, (1.0,))),
                 ("replace_membrane", ("chlorine_inlet",), (2, 2, ()))],
           1900: [("calibrate_two_point", ("pH_outlet", 7.0, 4.0, 7.02, 4.05), (1, 0, (7.0, 4.0, 7.02, 4.05))),
                  ("replace_reagent", ("chlorine_outlet",), (3, 3, ())),
                  ("clean_electrode", ("pH_outlet", "water_rinse"), (1, 1, (0.0,)))]}
    for k in range(0, 4000, 5):       # 5 s between reads: covers the 1800 s pH warm-up after each op
        t = t0 + k
        for name, gargs, (si, op, oargs) in ops.get(k, []):
            getattr(suite, name)(*gargs, t)
            assert osu.maintain(si, op, t, tuple(oargs) + (0.0,) * (4 - len(oargs))) == 0
        eng.step(1.0, e.bnd)
        out_g, st_g, ft_g = _gpu_read(suite, t)
        out_o, st_o, ft_o = osu.read(eng.state_numpy(), eng.state.flow_rate.cpu().numpy(), t, n)
        assert np.array_equal(st_g.T, st_o) and np.array_equal(ft_g.T, ft_o), k
        if k % 100 == 0 or k in (405, 700, 705, 1905, 2200, 2205, 3700, 3705):
            assert _close(np.transpose(out_g, (1, 0, 2)), out_o), k
    with pytest.raises(ValueError):
        suite.replace_membrane("chlorine_outlet", 5000.0)      # DPD has no membrane
    with pytest.raises(ValueError):
        suite.replace_reagent("chlorine_inlet", 5000.0)
    with pytest.raises(ValueError):
        suite.clean_electrode("pH_inlet", "sandblast", 5000.0)
    with pytest.raises(ValueError):
        suite.clean_electrode("flow_main", "water_rinse", 5000.0)


def test_get_statistics_matches_the_oracle(golden_dir):
    """wt_sensor_window_stats: BaseSensor.get_statistics over the device history ring, against the numpy oracle
    (pinned against the reference in tests/test_sensor_stats_oracle.py), including the reference's golden inputs
    pushed through the ring, a wrapped ring, an empty history and all-NaN sensors (warming up)."""
    from oracle import wt_sensor_stats_oracle as ws
    g = np.load(os.path.join(golden_dir, "sensor_statistics.npz"))
    K, P = g["values"].shape
    e = ens.config2(P, 10, seed=3)
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=1, history=16)       # ring shorter than the 40 readings: it wraps
    suite.initialize(0.0)
    st0 = suite.get_statistics("pH_inlet", 60.0)
    assert all(float(v.abs().max()) == 0.0 for v in st0.values())          # no readings yet: zeros
    vals = torch.from_numpy(g["values"]).to(eng.device)
    for k in range(K):
        # drive the ring directly with the golden readings (sensor 0), as read() does with its outputs
        suite._out[0, 0].copy_(vals[k])
        suite._hist[suite.read_index % suite.history].copy_(suite._out[0])
        suite._hist_times.append(float(g["timestamps"][k]))
        if len(suite._hist_times) > suite.history:
            suite._hist_times.pop(0)
        suite.read_index += 1
    for win in (0.5, 10.0, 20.0, 1e6):
        got = suite.get_statistics("pH_inlet", win)
        want = ws.statistics(g["values"][-16:], g["timestamps"][-16:], win)   # what is still in the ring
        for i, k in enumerate(ws.FIELDS):
            a, b = got[k].cpu().numpy(), want[i]
            assert np.array_equal(np.isnan(a), np.isnan(b)), (k, win)
            ok = ~np.isnan(b)
            assert np.allclose(a[ok], b[ok], rtol=1e-13, atol=1e-15), (k, win)
    # through the real read(): 40 reads of a warming-up pH sensor are all NaN -> fault_rate 1, NaN moments
    suite2 = create_realistic_sensor_suite(eng, seed=1, history=8)
    suite2.initialize(0.0)
    for k in range(12):
        suite2.read(eng.state, float(k))
    s2 = suite2.get_statistics("pH_outlet", 5.0)
    assert float(s2["count"][0]) == 6 and float(s2["fault_rate"].min()) == 1.0 and bool(torch.isnan(s2["mean"]).all())
    f2 = suite2.get_statistics("flow_main", 100.0)                          # flow warms up in 10 s: finite values by now
    assert float(f2["count"][0]) == 8 and bool(torch.isfinite(f2["mean"]).any())


def test_variants_and_reset_value_for_value_on_random_plants(oracle, golden_dir):
    """Thermocouple temperature sensors, turbine flow meter, reset() + calibrate() of the flow meter, on 4,096
    random config-5 plants over the golden schedule: every field equals the CPU port, which is pinned in
    distribution against the unmodified reference on the same plants (tests/test_sensors_oracle.py)."""
    from ics_wt_physicsengine_b200.sensors import FlowSensorType, TemperatureSensorType
    g = np.load(os.path.join(golden_dir, "sensors_random_plants.npz"))
    P, n, t0, t_first = 4096, 10, float(g["t0"]), float(g["t_first"])
    e = ens.config5(int(g["n"]), n, seed=int(g["seed"])).slice(slice(0, P))
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=777, temperature_sensor_type=TemperatureSensorType.THERMOCOUPLE_K,
                                          flow_sensor_type=FlowSensorType.TURBINE)
    suite.initialize(t0)
    osu = oracle.SensorSuiteOracle(e.cfg[:, 3], e.cfg[:, 12], e.cfg[:, 13], t0, seed=777, nthreads=8, temp_kind=2, flow_kind=1)
    y = eng.state_numpy()
    flow = eng.state.flow_rate.cpu().numpy()
    for k in range(int(g["checks"].max()) + 1):
        t = t_first + k
        if k == int(g["k_reset"]):
            suite.reset("flow_main", t)
            osu.reset(4, t)
        if k == int(g["k_recal"]):
            suite.calibrate("flow_main", e.cfg[:, 3], t)
            osu.calibrate(4, e.cfg[:, 3], t)
        out_g, st_g, ft_g = _gpu_read(suite, t)
        out_o, st_o, ft_o = osu.read(y, flow, t, n)
        assert np.array_equal(st_g.T, st_o) and np.array_equal(ft_g.T, ft_o), k
        assert _close(np.transpose(out_g, (1, 0, 2)), out_o), k
    # ... and directly against the reference's samples of the same plants, in distribution
    ci = list(g["checks"]).index(130)
    for s in (4, 5, 6):
        a = g["var_values"][ci, s].astype(np.float64)
        b = out_g[s, :, 0]
        assert compare_distributions(a[np.isfinite(a)], b[np.isfinite(b)], ("var", s))


def test_standard_suite_on_random_plants_against_reference_samples(golden_dir):
    """The factory's suite on 10,240 random config-5 plants against the unmodified reference on the same plants."""
    g = np.load(os.path.join(golden_dir, "sensors_random_plants.npz"))
    P, n = int(g["n"]), 10
    e = ens.config5(P, n, seed=int(g["seed"]))
    eng = PlantEnsemble(e)
    suite = create_realistic_sensor_suite(eng, seed=31)
    suite.initialize(float(g["t0"]))
    checks, ci, n_tested = list(g["checks"]), 0, 0
    for k in range(max(checks) + 1):
        r = suite.read(None, float(g["t_first"]) + k)
        if ci < len(checks) and k == checks[ci]:
            for s, name in enumerate(SENSOR_NAMES):
                a, ar = g["std_values"][ci, s].astype(np.float64), g["std_raw"][ci, s].astype(np.float64)
                b, br = r[name].value.cpu().numpy(), r[name].raw_value.cpu().numpy()
                pa, pb = np.isnan(a).mean(), np.isnan(b).mean()
                assert abs(pa - pb) < 5 * np.sqrt(max(pa * (1 - pa), 1e-4) * 2 / P) + 1e-3, (k, name, pa, pb)
                ha = np.bincount(g["std_status"][ci, s].astype(int), minlength=12) / P
                hb = np.bincount(r[name].status.cpu().numpy(), minlength=12) / P
                assert np.abs(ha - hb).max() < 0.02, (k, name)
                ma, mb = np.isfinite(a), np.isfinite(b)
                if ma.sum() >= 500:
                    n_tested += bool(compare_distributions(a[ma], b[mb], (k, name)))
                    compare_distributions((a - ar)[ma], (b - br)[mb], (k, name, "value - raw"))
            ci += 1
    assert n_tested >= 20


def test_clocked_reads_equal_host_timed_reads():
    """read_clocked() (time, previous time and read index from the device clock; what a CUDA graph replays) gives
    the same readings as read(state, t) with host scalars."""
    P, n, t0, dt = 768, 10, 10.0, 0.5
    e = ens.config2(P, n, seed=8)
    a, b = PlantEnsemble(e), PlantEnsemble(e)
    sa, sb = create_realistic_sensor_suite(a, seed=9), create_realistic_sensor_suite(b, seed=9)
    sa.initialize(t0); sb.initialize(t0)
    for k in range(5):
        sa.read(None, t0 + k * dt); sb.read(None, t0 + k * dt)
    sb.start_clock(t0 + 5 * dt, dt)
    for k in range(5, 80):
        sa.read(None, t0 + k * dt)
        sb.read_clocked()
    sb.account_reads(75)
    assert sb.read_index == sa.read_index and sb.last_time == sa.last_time
    assert torch.equal(torch.nan_to_num(sa._out, nan=-1.0), torch.nan_to_num(sb._out, nan=-1.0))
    assert torch.equal(sa._out_status, sb._out_status) and torch.equal(sa._sens_i, sb._sens_i)
    sa.read(None, t0 + 80 * dt); sb.read(None, t0 + 80 * dt)   # and host-timed reads continue from there
    assert torch.equal(torch.nan_to_num(sa._out, nan=-1.0), torch.nan_to_num(sb._out, nan=-1.0))
