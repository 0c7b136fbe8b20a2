"""CPU oracle of BaseSensor.get_statistics over a window of recent readings --
TEST INFRASTRUCTURE.  Restates base_sensor.py:757-775 (get_recent_readings), :777-807 (calculate_drift_rate)
and :809-856 (get_statistics), vectorised over plants.  Pinned against the unmodified reference by
tests/golden/sensor_statistics.npz (oracle/gen_golden_sensor_stats.py).  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

FIELDS = ("mean", "std", "min", "max", "count", "drift_rate", "fault_rate")


def window_rows(timestamps, window_seconds: float):
    """Indices of the readings inside the window, newest first (base_sensor.py:757-775)."""
    ts = np.asarray(timestamps, dtype=np.float64)
    if ts.size == 0:
        return np.zeros(0, dtype=np.int64)
    cutoff = ts[-1] - window_seconds
    return np.array([i for i in range(ts.size - 1, -1, -1) if ts[i] >= cutoff], dtype=np.int64)


def statistics(values: np.ndarray, timestamps, window_seconds: float = 60.0) -> np.ndarray:
    """values [K, P] reading values of one sensor (oldest first, may hold NaN / inf), timestamps [K].
    Returns [7, P] in FIELDS order."""
    values = np.asarray(values, dtype=np.float64)
    P = values.shape[1] if values.ndim == 2 else 0
    rows = window_rows(timestamps, window_seconds)
    out = np.zeros((7, P))
    if rows.size == 0:
        return out                                            # base_sensor.py:821-830
    v = values[rows]                                          # newest first
    fin = np.isfinite(v)
    nfin = fin.sum(axis=0)
    with np.errstate(all="ignore"):
        cnt = np.maximum(nfin, 1)
        mean = np.where(fin, v, 0.0).sum(axis=0) / cnt
        var = (np.where(fin, v - mean, 0.0) ** 2).sum(axis=0) / cnt
        out[0] = np.where(nfin > 0, mean, np.nan)
        out[1] = np.where(nfin > 0, np.sqrt(var), np.nan)
        out[2] = np.where(nfin > 0, np.where(fin, v, np.inf).min(axis=0), np.nan)
        out[3] = np.where(nfin > 0, np.where(fin, v, -np.inf).max(axis=0), np.nan)
    out[4] = rows.size
    # calculate_drift_rate (base_sensor.py:777-807): the window is newest-first, so times[-1] - times[0] is never
    # positive and the method returns 0.0 for every history with increasing timestamps (kept as is)
    out[5] = 0.0
    out[6] = np.where(nfin > 0, (rows.size - nfin) / rows.size, 1.0)
    return out
