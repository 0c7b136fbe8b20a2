"""The delay-line search of the sensor kernel (csrc/wt_sensors.cuh, wt_transport_sample): the guess-and-verify fast path
must return the reference's answer -- the FIRST minimum of |timestamp - target| in deque order, SampleLine.transport_sample
(base_sensor.py:177-216) -- whenever it accepts.  This is a Python restatement of the acceptance rule, checked against a
brute-force first-minimum on random non-decreasing timestamp sequences with plateaus (the deque holds equal timestamps
in pairs: a pH and a temperature sensor share each line); the kernel itself is compared value for value with the CPU
port in tests/test_gpu_sensors.py."""
import numpy as np

INF = float("inf")


def fast_path(ts, target, jg):
    """ts: timestamps newest first; jg: guessed entries-back-from-newest.  Returns the accepted j or None (fall through)."""
    count = len(ts)
    jg = min(jg, count - 1)
    ja = max(jg - 3, 0)
    jb = min(ja + 7, count - 1)
    d = [abs(ts[ja + k] - target) if ja + k <= jb else INF for k in range(8)]
    dmin, m = d[0], 0
    for k in range(1, 8):
        if d[k] <= dmin:
            dmin, m = d[k], k
    jm = ja + m
    if (ja == 0 or d[0] > dmin) and (jm < jb or jb == count - 1) and dmin < INF:
        return jm
    return None


def first_minimum(ts, target):
    arr = ts[::-1]   # deque order: oldest first
    return len(arr) - 1 - int(np.argmin(np.abs(arr - target)))


def test_fast_path_never_accepts_a_wrong_entry():
    rng = np.random.default_rng(20260005)
    accepted = 0
    for _ in range(60000):
        count = int(rng.integers(1, 101))
        ts = np.cumsum(rng.choice([0, 0, 1, 1, 1, 2, 0.5], size=count))[::-1].copy()
        target = ts[0] - rng.choice([0, 3, 5, 7.25, 29.5, 30, 30.5, 200])
        r = fast_path(ts, target, int(rng.integers(0, 120)))
        if r is not None:
            accepted += 1
            assert r == first_minimum(ts, target)
    assert accepted > 10000


def test_steady_reads_take_the_fast_path():
    """Two entries per read at a steady interval (the shared lines of the factory's suite), delay 30 s: the guess
    2 * delay / dt lands on the wanted entry for both sensors of a pair."""
    for extra in (0, 1):   # the second sensor of a pair has one more entry of the current time in front
        t = np.repeat(np.arange(100.0, 50.0, -1.0), 2)[: 100 - extra]
        ts = np.concatenate([[100.0] * extra, t])[:100]
        r = fast_path(ts, 100.0 - 30.0, int(2.0 * 30.0 / 1.0 + 0.5))
        assert r is not None and r == first_minimum(ts, 70.0)
