"""The get_statistics oracle (oracle/wt_sensor_stats_oracle.py) against the unmodified reference
(tests/golden/sensor_statistics.npz, oracle/gen_golden_sensor_stats.py)."""
import os

import numpy as np

from oracle import wt_sensor_stats_oracle as ws


def test_matches_the_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "sensor_statistics.npz"))
    assert tuple(g["fields"]) == ws.FIELDS
    for w, win in enumerate(g["windows"]):
        got = ws.statistics(g["values"], g["timestamps"], float(win))
        want = g["out"][w]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.allclose(got[ok], want[ok], rtol=1e-13, atol=1e-15), win
    assert np.array_equal(ws.statistics(np.zeros((0, 5)), [], 60.0), np.zeros((7, 5)))
    assert np.array_equal(g["empty"], np.zeros(7))          # no readings yet: all zeros (base_sensor.py:821-830)
    assert not g["out"][:, 5].any()                          # drift_rate is always 0.0 in the reference (newest-first window)
