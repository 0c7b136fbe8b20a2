"""Command -> boundary path of the reference's main loop (__main__.py:57-63, 227-271).
CPU: the numpy oracle against the outputs of the reference's own functions (tests/golden/commands.npz).
GPU: the wt_apply_commands / wt_scenario_commands kernels, bit-exact against the oracle and the golden outputs."""
import os

import numpy as np
import pytest

from oracle import wt_commands_oracle as co


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "commands.npz"))


def test_oracle_matches_the_reference_functions(golden):
    g = golden
    for j, mx in enumerate(g["vfr_max"]):
        assert np.array_equal(co.validate_flow_rate(g["vfr_in"], mx), g["vfr_out"][j])
    a, c, i = co.apply_commands(g["commands"][:, 0], g["commands"][:, 1], g["commands"][:, 2], g["inlet_before"])
    assert np.array_equal(np.stack([a, c, i], axis=1), g["boundary_after"])
    assert np.isnan(g["commands"]).any() and np.isinf(g["commands"]).any()   # the edge cases are in the vectors


def test_scenario_segment_lookup():
    times = [0.0, 10.0, 25.0]
    cmd = np.arange(2 * 3 * 3, dtype=np.float64).reshape(2, 3, 3)
    assert co.scenario_commands(times, cmd, [0, 1], -1.0) is None
    assert np.array_equal(co.scenario_commands(times, cmd, [0, 1], 0.0), cmd[[0, 1], 0])
    assert np.array_equal(co.scenario_commands(times, cmd, [0, 1], 24.999), cmd[[0, 1], 1])
    assert np.array_equal(co.scenario_commands(times, cmd, [1, 0, 7], 1e9), cmd[[1, 0, 1], 2])


@pytest.mark.gpu
def test_apply_commands_kernel_is_bit_exact(golden):
    torch = pytest.importorskip("torch")
    from ics_wt_physicsengine_b200.ensembles import BND_FIELDS, default_bnd_row
    from ics_wt_physicsengine_b200.orchestrator import apply_boundary_conditions
    g = golden
    P = g["commands"].shape[0]
    bnd = np.repeat(default_bnd_row()[:, None], P, axis=1).copy()
    bnd[BND_FIELDS.index("inlet_flow_rate")] = g["inlet_before"]
    b = torch.from_numpy(bnd).cuda()
    cmd = torch.from_numpy(g["commands"]).cuda()
    apply_boundary_conditions(b, cmd[:, 0], cmd[:, 1], cmd[:, 2])
    got = b.cpu().numpy()
    rows = [BND_FIELDS.index(k) for k in ("acid_flow_rate", "chlorine_flow_rate", "inlet_flow_rate")]
    assert np.array_equal(got[rows].T, g["boundary_after"])
    others = [r for r in range(10) if r not in rows]
    assert np.array_equal(got[others], bnd[others])


@pytest.mark.gpu
def test_scenario_table_drives_a_stepped_ensemble():
    """Scripted scenario (two scripts, three segments) on the device clock: boundary rows after every step equal the
    oracle's, with no host-to-device copy inside the loop; the plants respond (acid dosing lowers the pH)."""
    torch = pytest.importorskip("torch")
    from ics_wt_physicsengine_b200 import PlantEnsemble, ensembles as ens
    from ics_wt_physicsengine_b200.ensembles import BND_FIELDS
    from ics_wt_physicsengine_b200.orchestrator import EnsembleOrchestrator, ScenarioTable
    from ics_wt_physicsengine_b200.sensors import create_realistic_sensor_suite
    P, n = 1024, 10
    e = ens.config2(P, n, seed=21)
    eng = PlantEnsemble(e)
    bnd = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).cuda()
    times = [0.0, 5.0, 12.0]
    cmd = np.array([[[0.0, 0.0, 0.0], [1.5, 0.2, 8.0], [float("nan"), 5.0, 0.05]],
                    [[3.0, 0.5, 25.0], [0.0, 0.0, 0.0], [0.3, 0.1, 12.0]]])
    sid = (np.arange(P) % 2).astype(np.int32)
    table = ScenarioTable(times, cmd, sid)
    suite = create_realistic_sensor_suite(eng, seed=1)
    suite.initialize(0.0)
    orch = EnsembleOrchestrator(eng, suite, bnd, t0=0.0)
    want = e.bnd.copy()
    rows = [BND_FIELDS.index(k) for k in ("acid_flow_rate", "chlorine_flow_rate", "inlet_flow_rate")]
    pH0 = eng.state.pH.clone()
    for k in range(16):
        orch.run(1, 1.0, scenario=table)
        c = co.scenario_commands(times, cmd, sid, float(k + 1))
        if c is not None:
            a, cl, i = co.apply_commands(c[:, 0], c[:, 1], c[:, 2], want[:, rows[2]])
            want[:, rows[0]], want[:, rows[1]], want[:, rows[2]] = a, cl, i
        assert np.array_equal(bnd.cpu().numpy().T, want), k
    # the device-clock form (what a captured graph replays) gives the same rows
    clock = torch.tensor([12.5, 11.5, 12.0, 0.5, 1.0], dtype=torch.float64).cuda()
    b2 = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).cuda()
    table.apply(b2, clock=clock)
    b3 = torch.from_numpy(np.ascontiguousarray(e.bnd.T)).cuda()
    table.apply(b3, t=12.5)
    assert torch.equal(b2, b3)
    assert not torch.equal(eng.state.pH, pH0)   # and the plants were stepped with those boundaries
