"""Shared helpers for the parity tests."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HALT = 2 | 128
EXCUSE_CEILING = 1e-6  # largest error an ill-conditioned (excused) plant-step may show


def relerr(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def species_major(e):
    """Ensemble initial state -> [P, 3n] (the reference's ODE vector layout)."""
    return np.concatenate([e.pH0, e.Cl0, e.T0], axis=1).copy()


class EmuLib:
    """CPU lane-emulation build of the kernel core (tests/emu/wt_emu.cpp) -- test infrastructure."""

    def __init__(self):
        src = os.path.join(ROOT, "tests", "emu", "wt_emu.cpp")
        so = os.path.join(ROOT, "tests", "emu", "libwt_emu.so")
        inc = os.path.join(ROOT, "ics_wt_physicsengine_b200", "csrc")
        deps = [src, os.path.join(inc, "wt_step_core.h"), os.path.join(inc, "wt_simt.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DWT_EMU", "-ffp-contract=off",
                                   "-I" + inc, "-o", so, src])
        self.L = C.CDLL(so)

    def step(self, par, bnd, n, t, y, dt=1.0, nsteps=1, max_attempts=0, status=None):
        dp, ip, up = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
        P = y.shape[0]
        st = np.zeros(P, np.uint32) if status is None else status
        cnt = np.zeros((P, 8), np.int32)
        fl = np.zeros(P)
        bnd = np.ascontiguousarray(bnd, dtype=np.float64)
        assert bnd.shape == (P, 10)
        self.L.wt_emu_step_batch(P, n, nsteps, C.c_double(dt), par.ctypes.data_as(dp), bnd.ctypes.data_as(dp), 10,
                                 t.ctypes.data_as(dp), y.ctypes.data_as(dp), fl.ctypes.data_as(dp),
                                 st.ctypes.data_as(up), cnt.ctypes.data_as(ip), None, int(max_attempts))
        return st, cnt, fl

    def rhs(self, par, bnd, n, y):
        dp = C.POINTER(C.c_double)
        dy = np.zeros(3 * n)
        bad = C.c_int(0)
        par = np.ascontiguousarray(par)
        bnd = np.ascontiguousarray(bnd)
        y = np.ascontiguousarray(y)
        self.L.wt_emu_rhs(par.ctypes.data_as(dp), bnd.ctypes.data_as(dp), n, y.ctypes.data_as(dp),
                          dy.ctypes.data_as(dp), C.byref(bad))
        return dy, bad.value


def oracle_sensitivity(wo, par, bnd, n, t0, y0, dt, max_attempts, seed=0, trials=3):
    """Largest relative change of the oracle's own one-step result under 1-ulp input perturbations.

    A plant-step whose reference result moves by more than the parity tolerance when its input
    moves by one ulp cannot be matched to that tolerance by ANY other implementation (the
    reference itself is not reproducible there, e.g. across BLAS builds): see DESIGN.md.
    """
    rng = np.random.default_rng(seed)
    wo.set_max_attempts(max_attempts)
    base = y0.copy()[None, :]
    tb = np.array([t0])
    wo.step_batch(par[None, :].copy(), bnd[None, :].copy(), n, tb, base, dt=dt)
    worst = 0.0
    for _ in range(trials):
        yp = (y0 * (1.0 + rng.choice([-1.0, 1.0], size=y0.size) * 1.1e-16))[None, :].copy()
        tp = np.array([t0])
        wo.step_batch(par[None, :].copy(), bnd[None, :].copy(), n, tp, yp, dt=dt)
        worst = max(worst, float(relerr(yp, base).max()))
    return worst


def check_step_parity(wo, got, want, par, bnd, n, t_before, y_before, dt, max_attempts, tol=1e-9, what=""):
    """Per-plant one-step parity: |got - want| <= tol * |want| for every zone variable, except
    plant-steps the oracle itself cannot reproduce under a 1-ulp input change."""
    r = relerr(got, want).max(axis=1)
    bad = np.nonzero(r > tol)[0]
    excused = []
    for p in bad:
        b = bnd[p] if bnd.ndim == 2 else bnd
        s = oracle_sensitivity(wo, par[p], b, n, float(t_before[p]), y_before[p], dt, max_attempts, seed=int(p))
        # excused only while the error stays within 10x what one ulp of input noise does to the oracle itself,
        # and never beyond an absolute ceiling
        if r[p] <= 10.0 * s and r[p] <= EXCUSE_CEILING:
            excused.append((int(p), float(r[p]), s))
        else:
            raise AssertionError(f"{what}: plant {p} off by {r[p]:.3e} (tol {tol:.1e}); the oracle's own 1-ulp "
                                 f"sensitivity there is only {s:.3e}")
    return r, excused
