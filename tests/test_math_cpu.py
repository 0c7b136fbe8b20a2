"""The kernel's branch-free exp / exp10 (csrc/wt_simt.h) against libm, on the CPU.

They are written with fma() and integer operations only, so this build produces the same bits as the
sm_100a build; the GPU parity tests then cover them inside the kernel."""
import math
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run():
    src = os.path.join(ROOT, "tests", "cpu_math", "wt_math_check.cpp")
    exe = os.path.join(ROOT, "tests", "cpu_math", "wt_math_check")
    inc = os.path.join(ROOT, "ics_wt_physicsengine_b200", "csrc")
    deps = [src, os.path.join(inc, "wt_simt.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-I" + inc, "-o", exe, src, "-lm"])
    return subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()


def test_exp_and_exp10_within_two_ulp_of_libm():
    out = dict(l.split() for l in _run()[:3])
    assert float(out["max_ulp_exp"]) <= 2.0
    assert float(out["max_ulp_exp10"]) <= 2.0
    assert float(out["max_denormal_units"]) <= 1.0  # gradual underflow comes out of the arithmetic, no branch


def test_special_cases_match_libm():
    for line in _run()[3:]:
        name, x, got, want = line.split()
        g, w = float.fromhex(got) if "nan" not in got else math.nan, float.fromhex(want) if "nan" not in want else math.nan
        if math.isnan(w):
            assert math.isnan(g), line
        elif math.isinf(w) or w == 0.0:
            assert g == w, line
        else:
            assert abs(g - w) <= 4.0 * abs(w) * 2.0 ** -52 + 5e-324, line
