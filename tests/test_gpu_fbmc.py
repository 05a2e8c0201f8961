"""Force-bias Monte Carlo beyond the golden trajectory: the statistical property the reference tests
(tests/mc/test_fbmc.py:12-26: mean |displacement| = delta / 3 at weak bias), the driver plumbing, restart."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fb(R=32, **kw):
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import ForceBias
    from tests.helpers import cu_replicas

    pos, cell, n = cu_replicas(3, R, seed=17, stdev=0.02)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    return ForceBias(atoms, delta=0.01, temperature=5000.0, seed=42, **kw)


def test_mean_displacement_is_a_third_of_delta(cuda_lib, tmp_path):
    fb = _fb(logfile=tmp_path / "fb.log", trajectory=tmp_path / "fb.xyz", logging_interval=25, logging_mode="w")
    fb.default_trajectory.write_kwargs["replicas"] = [0]
    steps, moved = 50, []
    last = fb.download().positions.copy()
    for step in fb.irun(steps):
        for _ in step:
            pass
        now = fb.download().positions.copy()
        moved.append(np.abs(now - last))
        last = now
    fb.close()
    mean_abs = np.mean(moved, axis=0)  # (R, n, 3): weak bias -> triangular zeta -> <|zeta|> = 1 / 3
    assert abs(mean_abs.mean() - 0.01 / 3) < 0.01 / 60
    assert np.max(moved) < 0.01 and fb.step_count == steps
    lines = (tmp_path / "fb.log").read_text().splitlines()
    assert lines[0].split() == ["Class", "Step", "Epot[eV]", "Gamma/GammaMax"] and [ln.split()[1] for ln in lines[1:]] == ["0", "25", "50"]
    assert float(lines[-1].split()[3]) < 0.05 and 0.0 < fb.gamma.max() < 1.0  # weak bias: gamma far below the clipping value
    assert 3 <= fb.rejection_rounds.min() and fb.rejection_rounds.max() < 40


def test_force_bias_restart_and_dictionary(cuda_lib, tmp_path):
    from quansino_b200 import EMT
    from quansino_b200.io import RestartObserver, restore_simulation

    straight = _fb(R=4)
    d = straight.to_dict()  # tests/mc/test_fbmc.py:68-93
    assert d["name"] == "ForceBias" and d["kwargs"]["seed"] == 42 and d["kwargs"]["delta"] == 0.01 and d["attributes"]["step_count"] == 0
    assert np.all(d["kwargs"]["temperature"] == 5000.0) and d["attributes"]["masses_scaling_power"] == 0.25
    straight.run(6)
    first = _fb(R=4)
    first.attach(RestartObserver(first, tmp_path / "fb.json", interval=3, mode="w"), name="restart")
    first.run(3)
    first.close()
    second = restore_simulation(tmp_path / "fb.json", EMT())
    assert type(second).__name__ == "ForceBias" and second.step_count == 3
    second.run(3)
    assert np.array_equal(second.download().positions, straight.download().positions)
    np.testing.assert_allclose(second.context.last_potential_energy, straight.context.last_potential_energy, rtol=1e-13)
