"""The CPU port of the sensor suite (oracle/wt_sensors_oracle.c) against 10,240 instances of the
UNMODIFIED reference suite (tests/golden/sensors_default_plant.npz, oracle/gen_golden_sensors.py).
RNG streams differ by construction (the reference seeds from secrets), so the pin is distributional:
dead-sensor (NaN) fractions, status / fault histograms, first four moments and a two-sample KS test
per sensor at the recorded check times."""
import os

import numpy as np
import pytest
from scipy import stats

N_ORACLE = 10240


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "sensors_default_plant.npz"))


@pytest.fixture(scope="module")
def oracle_samples(oracle, golden):
    """Run N_ORACLE oracle suites over the golden default-plant trajectory; record the check reads."""
    g = golden
    checks = list(g["checks"])
    t0 = float(g["t0"])
    n = g["traj_pH"].shape[1]
    P = N_ORACLE
    suite = oracle.SensorSuiteOracle(np.full(P, 5.0), np.full(P, 2.0), np.full(P, 20.0), t0, seed=12345, nthreads=8)
    vals = np.zeros((len(checks), 7, P))
    stat = np.zeros((len(checks), 7, P), dtype=np.int32)
    flt = np.zeros((len(checks), 7, P), dtype=np.int32)
    ci = 0
    for k in range(max(checks) + 1):
        y = np.concatenate([g["traj_pH"][k], g["traj_Cl"][k], g["traj_T"][k]])
        out, st, ft = suite.read(np.broadcast_to(y, (P, 3 * n)), np.full(P, float(g["traj_flow"][k])), t0 + k, n)
        if ci < len(checks) and k == checks[ci]:
            vals[ci], stat[ci], flt[ci] = out[:, :, 0].T, st.T, ft.T
            ci += 1
    return vals, stat, flt


def test_dead_sensor_fractions_and_status_histograms(golden, oracle_samples):
    vals, stat, flt = oracle_samples
    gv, gs, gf = golden["values"], golden["status"], golden["fault"]
    N = gv.shape[2]
    for ci, k in enumerate(golden["checks"]):
        for s in range(7):
            pa, pb = np.isnan(gv[ci, s]).mean(), np.isnan(vals[ci, s]).mean()
            se = np.sqrt(max(pa * (1 - pa), 1e-4) * (1 / N + 1 / N_ORACLE))
            assert abs(pa - pb) < 5 * se + 1e-3, (int(k), s, pa, pb)
            ha = np.bincount(gs[ci, s].astype(int), minlength=12) / N
            hb = np.bincount(stat[ci, s], minlength=12) / N_ORACLE
            assert np.abs(ha - hb).max() < 0.02, (int(k), s, ha, hb)
            fa = np.bincount(gf[ci, s].astype(int), minlength=7) / N
            fb = np.bincount(flt[ci, s], minlength=7) / N_ORACLE
            assert np.abs(fa - fb).max() < 0.02, (int(k), s, fa, fb)


def _main_mode(x):
    """Values within 6 robust sigmas of the median (the dominant mode of a possibly mixed sample)."""
    med = np.median(x)
    mad = 1.4826 * np.median(np.abs(x - med)) + 1e-12
    return x[np.abs(x - med) < 6 * mad]


def compare_distributions(a, b, what):
    """a: reference sample, b: candidate sample (finite values).  Point masses at the clip bounds are
    compared as fractions, the whole sample with a two-sample KS test, and the dominant mode with its
    first four moments (mixtures such as 'sensor whose line partner died' make raw moments meaningless)."""
    lo, hi = min(a.min(), b.min()), max(a.max(), b.max())
    for bound in (lo, hi):
        fa, fb = np.mean(a == bound), np.mean(b == bound)
        assert abs(fa - fb) < 0.02, (what, "mass at bound", bound, fa, fb)
    ks = stats.ks_2samp(a, b)
    assert ks.statistic < 0.03 or ks.pvalue > 1e-4, (what, ks)
    ai, bi = a[(a > lo) & (a < hi)], b[(b > lo) & (b < hi)]
    if ai.size < 500 or bi.size < 500:
        return False
    assert abs(ai.size / a.size - bi.size / b.size) < 0.03, (what, "interior fraction")
    am, bm = _main_mode(ai), _main_mode(bi)
    assert abs(am.size / ai.size - bm.size / bi.size) < 0.03, (what, "main-mode fraction")
    sa, sb = am.std(), bm.std()
    assert abs(am.mean() - bm.mean()) < 6 * np.sqrt(sa ** 2 / am.size + sb ** 2 / bm.size) + 1e-9, (what, "mean", am.mean(), bm.mean())
    assert abs(sa - sb) < 0.06 * max(sa, sb) + 1e-9, (what, "std", sa, sb)
    assert abs(stats.skew(am) - stats.skew(bm)) < 0.3, (what, "skew")
    assert abs(stats.kurtosis(am) - stats.kurtosis(bm)) < 1.0, (what, "kurtosis")  # one-sided clipped modes: large sampling error
    return True


def test_value_distributions_match_reference(golden, oracle_samples):
    vals, _, _ = oracle_samples
    gv = golden["values"]
    n_tested = 0
    for ci, k in enumerate(golden["checks"]):
        for s in range(7):
            a = gv[ci, s][np.isfinite(gv[ci, s])]
            b = vals[ci, s][np.isfinite(vals[ci, s])]
            if a.size < 500:
                assert b.size < 0.1 * N_ORACLE, (int(k), s)
                continue
            n_tested += bool(compare_distributions(a, b, (int(k), s)))
    assert n_tested >= 20


def test_emergent_quirks_are_reproduced(golden, oracle_samples):
    """the behaviours that dominate the reference's output distribution."""
    vals, stat, _ = oracle_samples
    checks = list(golden["checks"])
    m = lambda ci, s: np.nanmedian(vals[ci, s])
    i1805, i1840, i400, i40 = checks.index(1805), checks.index(1840), checks.index(400), checks.index(40)
    assert np.isnan(vals[i400, 0]).all()                      # pH sensors still warming up (1800 s)
    assert m(i1805, 0) > 12.5                                  # pH wakes up on a delayed TEMPERATURE sample
    assert 6.9 < m(i1840, 0) < 7.1                             # ... then reads pH again
    assert 11.0 < m(i1840, 5) < 12.6                           # temp sensor now reads the delayed pH (shared line)
    assert abs(m(i1840, 5) - np.nanmedian(golden["values"][i1840, 5])) < 0.05
    assert 24.5 < m(i40, 5) < 25.5                             # RTD lead-wire offset fed back through the lag filter
    assert 3.0 < m(i400, 2) < 4.2                              # chlorine reads true + calibration offset (= reference)
    assert m(i400, 4) > 9.9                                    # flow saturates at full scale (offset = flow_rate)
    assert (stat[i400, 4] == 5).mean() > 0.5                   # ... and sits in DRIFT_WARNING
    assert np.isnan(vals[checks.index(1900), 4]).mean() > 0.15  # absorbing power / open-circuit faults


def test_maintenance_operations_match_the_reference(oracle, golden_dir):
    """calibrate_two_point / clean_electrode / replace_membrane / replace_reagent of the CPU port against the
    attributes of the unmodified reference's sensors before and after the same call
    (tests/golden/sensor_maintenance.npz, oracle/gen_golden_maint.py), including the ValueError cases."""
    import os
    g = np.load(os.path.join(golden_dir, "sensor_maintenance.npz"))
    fields = oracle.SensorSuiteOracle.FIELDS
    assert tuple(g["attrs"]) == fields[:11]
    n_raise = 0
    for row in g["cases"]:
        si, op, t, args = int(row[0]), int(row[1]), row[2], row[3:7]
        before, after, raised = row[7:20], row[20:33], int(row[33])
        su = oracle.SensorSuiteOracle(np.array([5.0]), np.array([2.0]), np.array([20.0]), 0.0)
        for i, f in enumerate(fields[:11]):
            su.poke(si, f, before[i])
        rc = su.maintain(si, op, t, args)
        assert (rc != 0) == bool(raised)
        n_raise += raised
        if raised:
            continue
        got = su.peek(si)
        for i, f in enumerate(fields):
            kind_has = not (after[i] == 0.0 and before[i] == 0.0)   # attributes the sensor kind does not have are 0 in the dump
            if kind_has or f in ("status", "fault"):
                assert got[f][0] == after[i], (si, op, f, got[f][0], after[i])
    assert n_raise == 120  # replace_membrane on the DPD sensor and replace_reagent on the amperometric one


# ---- random config-5 plants, both variant sets, reset() (tests/golden/sensors_random_plants.npz) -----------------
@pytest.fixture(scope="module")
def golden_random(golden_dir):
    return np.load(os.path.join(golden_dir, "sensors_random_plants.npz"))


def _oracle_random_run(oracle, g, variants):
    from ics_wt_physicsengine_b200 import ensembles
    from ics_wt_physicsengine_b200.ensembles import CFG_FIELDS
    N, checks = int(g["n"]), list(g["checks"])
    e = ensembles.config5(N, 10, seed=int(g["seed"]))
    col = lambda k: np.ascontiguousarray(e.cfg[:, CFG_FIELDS.index(k)])
    suite = oracle.SensorSuiteOracle(col("flow_rate"), col("initial_chlorine"), col("temperature"), float(g["t0"]), seed=777,
                                     nthreads=8, temp_kind=2 if variants else 0, flow_kind=1 if variants else 0)
    y = np.concatenate([e.pH0, e.Cl0, e.T0], axis=1)
    vals = np.zeros((len(checks), 7, N)); raw = np.zeros_like(vals)
    stat = np.zeros((len(checks), 7, N), dtype=np.int32); flt = np.zeros_like(stat)
    ci = 0
    for k in range(max(checks) + 1):
        t = float(g["t_first"]) + k
        if variants and k == int(g["k_reset"]):
            suite.reset(4, t)
        if variants and k == int(g["k_recal"]):
            suite.calibrate(4, col("flow_rate"), t)
        out, st, ft = suite.read(y, col("flow_rate"), t, 10)
        if ci < len(checks) and k == checks[ci]:
            vals[ci], raw[ci], stat[ci], flt[ci] = out[:, :, 0].T, out[:, :, 1].T, st.T, ft.T
            ci += 1
    return vals, raw, stat, flt


@pytest.mark.parametrize("tag", ["std", "var"])
def test_random_plants_match_the_reference_in_distribution(oracle, golden_random, tag):
    """10,240 RANDOM config-5 plants (own full scale, calibration references, temperatures, zone profiles): the port
    against the unmodified reference, for the factory's suite ("std") and for thermocouple + turbine sensors with a
    reset() / calibrate() of the flow meter on the way ("var")."""
    g = golden_random
    vals, raw, stat, flt = _oracle_random_run(oracle, g, tag == "var")
    gv, gr, gs, gf = g[f"{tag}_values"].astype(np.float64), g[f"{tag}_raw"].astype(np.float64), g[f"{tag}_status"], g[f"{tag}_fault"]
    N = gv.shape[2]
    n_tested = 0
    for ci, k in enumerate(g["checks"]):
        for s in range(7):
            pa, pb = np.isnan(gv[ci, s]).mean(), np.isnan(vals[ci, s]).mean()
            se = np.sqrt(max(pa * (1 - pa), 1e-4) * (2.0 / N))
            assert abs(pa - pb) < 5 * se + 1e-3, (tag, int(k), s, pa, pb)
            ha = np.bincount(gs[ci, s].astype(int), minlength=12) / N
            hb = np.bincount(stat[ci, s], minlength=12) / N
            assert np.abs(ha - hb).max() < 0.02, (tag, int(k), s, ha, hb)
            fa = np.bincount(gf[ci, s].astype(int), minlength=7) / N
            fb = np.bincount(flt[ci, s], minlength=7) / N
            assert np.abs(fa - fb).max() < 0.02, (tag, int(k), s, fa, fb)
            ma, mb = np.isfinite(gv[ci, s]), np.isfinite(vals[ci, s])
            if ma.sum() < 500:
                continue
            # same plant population on both sides: the value and the residual value - raw_value are comparable pooled
            n_tested += bool(compare_distributions(gv[ci, s][ma], vals[ci, s][mb], (tag, int(k), s, "value")))
            compare_distributions((gv[ci, s] - gr[ci, s])[ma], (vals[ci, s] - raw[ci, s])[mb], (tag, int(k), s, "value - raw"))
    assert n_tested >= 20
    if tag == "var":   # reset(): warming up, then CALIBRATION_EXPIRED (6) while the calibration history is empty
        c = list(g["checks"])
        assert (stat[c.index(45), 4] == 2).mean() > 0.97 and (stat[c.index(62), 4] == 2).mean() > 0.97
        assert (stat[c.index(55), 4] == 6).mean() > 0.9 and (gs[c.index(55), 4] == 6).mean() > 0.9
