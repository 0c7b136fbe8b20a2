"""The kernel core (csrc/wt_step_core.h) compiled with -DWT_EMU -- the same source the sm_100a
kernel is built from, lanes widened to arrays -- against the dense CPU oracle.  Exercises what
is B200-specific in the design: zones-on-lanes, several plants per warp running in lockstep
under masks, the structured finite-difference Jacobian and the PCR tridiagonal solves."""
import numpy as np
import pytest

from ics_wt_physicsengine_b200 import ensembles as ens
from tests._util import HALT, check_step_parity, relerr, species_major

CAP = 64


def _run(oracle, emu, e, steps, dt=1.0, cap=CAP):
    n = e.n_zones
    par = oracle.derive_params(e.cfg, n)
    bnd = np.ascontiguousarray(e.bnd)
    oracle.set_max_attempts(cap)
    yo, to = species_major(e), np.zeros(e.n_plants)
    halted = np.zeros(e.n_plants, bool)
    excused_total = 0
    for s in range(steps):
        ye, te = yo.copy(), to.copy()
        y_before, t_before = yo.copy(), to.copy()
        so, co, fo = oracle.step_batch(par, bnd, n, to, yo, dt=dt, nthreads=4)
        se, ce, fe = emu.step(par, bnd, n, te, ye, dt=dt, max_attempts=cap)
        live = ~halted & ((so & HALT) == 0) & ((se & HALT) == 0)
        # halting decisions agree except on chaotic (budget-exhausting) plants
        assert ((so & HALT) != 0).sum() == pytest.approx(((se & HALT) != 0).sum(), abs=2)
        r, excused = check_step_parity(oracle, ye[live], yo[live], par[live], bnd[live], n, t_before[live],
                                       y_before[live], dt, cap, what=f"step {s}")
        excused_total += len(excused)
        well = np.ones(live.sum(), bool)
        well[[i for i, (p, _, _) in enumerate(excused)]] = True
        same_path = (co[live][:, :7] == ce[live][:, :7]).all(axis=1)
        assert same_path.mean() > 0.995
        assert np.array_equal(to[live], te[live]) and np.array_equal(fo[live], fe[live])
        # halted plants: state and time untouched
        stuck = (so & HALT) != 0
        assert np.array_equal(yo[stuck], y_before[stuck]) and np.array_equal(to[stuck], t_before[stuck])
        halted |= stuck | ((se & HALT) != 0)
        yo[halted] = y_before[halted]
    return excused_total


@pytest.mark.parametrize("n", [2, 3, 5, 7, 10, 16, 20, 32])
def test_zone_counts_and_plants_per_warp(oracle, emu, n):
    e = ens.config2(40, n, seed=100 + n)
    assert _run(oracle, emu, e, 4) <= 1


def test_default_plant_bitwise_path(oracle, emu):
    e = ens.config1()
    par = oracle.derive_params(e.cfg, 5)
    y, t = species_major(e), np.zeros(1)
    for _ in range(30):
        st, cnt, _ = emu.step(par, e.bnd, 5, t, y)
        assert st[0] == 0 and tuple(cnt[0, :4]) == (16, 1, 4, 2)
    yo, to = species_major(e), np.zeros(1)
    oracle.set_max_attempts(0)
    oracle.step_batch(par, e.bnd, 5, to, yo, nsteps=30)
    assert relerr(y, yo).max() < 1e-13 and t[0] == to[0] == 30.0


def test_config2_slice(oracle, emu):
    assert _run(oracle, emu, ens.config2(192), 8) <= 2


def test_config3_stratified_temperature_sweep(oracle, emu):
    assert _run(oracle, emu, ens.config3(96), 6) <= 3


@pytest.mark.parametrize("dt", [0.1, 10.0])
def test_other_dt(oracle, emu, dt):
    assert _run(oracle, emu, ens.config2(30, 10, seed=5), 3, dt=dt) <= 2


def test_rhs_lane_mapping(oracle, emu, golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, "rhs_config3.npz"))
    n = int(g["n_zones"])
    par = oracle.derive_params(g["cfg"], n)
    for p in range(16):
        dy, bad = emu.rhs(par[p], g["bnd"][p], n, g["Y"][p])
        ref, rc = oracle.rhs(par[p], g["bnd"][p], n, g["Y"][p])
        assert bad == rc == 0
        for v in range(3):
            blk = slice(v * n, (v + 1) * n)
            assert np.abs(dy[blk] - ref[blk]).max() <= 1e-13 * np.abs(ref[blk]).max() + 1e-300


def test_plants_in_one_warp_are_independent(oracle, emu):
    """A plant's result must not depend on which lanes it occupies or on its warp neighbours."""
    e = ens.config2(9, 10, seed=77)
    par = oracle.derive_params(e.cfg, 10)
    y1, t1 = species_major(e), np.zeros(9)
    emu.step(par, e.bnd, 10, t1, y1, nsteps=3)
    perm = np.array([4, 8, 0, 2, 6, 1, 7, 3, 5])
    y2, t2 = species_major(e)[perm].copy(), np.zeros(9)
    emu.step(np.ascontiguousarray(par[perm]), np.ascontiguousarray(e.bnd[perm]), 10, t2, y2, nsteps=3)
    assert np.array_equal(y2, y1[perm])


def test_work_limit_and_t_range_halt(oracle, emu):
    """A plant on the 8 C density discontinuity exhausts the budget: flagged, untouched, skipped."""
    e = ens.config1(10)
    e.T0[0] = np.linspace(8.4, 7.6, 10)  # straddles the 999.842 / 1000.715 kg/m3 jump (spatial.py:177-189)
    par = oracle.derive_params(e.cfg, 10)
    y, t = species_major(e), np.zeros(1)
    y0 = y.copy()
    st = np.zeros(1, np.uint32)
    _, cnt, _ = emu.step(par, e.bnd, 10, t, y, max_attempts=64, status=st)
    if st[0] & 128:
        assert np.array_equal(y, y0) and t[0] == 0.0
        _, cnt2, _ = emu.step(par, e.bnd, 10, t, y, max_attempts=64, status=st)
        assert cnt2.sum() == 0 and np.array_equal(y, y0)  # halted plants are skipped
    # T = 100 C: the finite-difference perturbation leaves [0, 100] -> ValueError in the reference
    e = ens.config1(5)
    e.T0[0] = 100.0
    par = oracle.derive_params(e.cfg, 5)
    y, t = species_major(e), np.zeros(1)
    st, _, _ = emu.step(par, e.bnd, 5, t, y)
    assert st[0] & 2 and t[0] == 0.0
