"""Statistical and conservation properties the reference's own tests assert, on the CUDA path: force-bias mean displacement
(tests/mc/test_fbmc.py:12-26), Maxwell-Boltzmann temperature (tests/utils/test_dynamics.py:23-32), Verlet energy
conservation (tests/integrators/test_verlet_integrator.py:22-31); plus the ForceBias driver plumbing and restart."""

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


def test_hamiltonian_known_properties(cuda_lib):
    """The reference's property tests for the Hamiltonian path: Maxwell-Boltzmann momenta give <T> = 300 K
    (tests/utils/test_dynamics.py:23-32) and Verlet with dt = 0.01 fs x 20 conserves the total energy to four decimals
    (tests/integrators/test_verlet_integrator.py:22-31)."""
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.integrators import Verlet
    from quansino_b200.mc import HamiltonianCanonical
    from quansino_b200.moves import HamiltonianDisplacementMove
    from tests.helpers import cu_replicas

    R = 256
    pos, cell, n = cu_replicas(3, R, seed=5, stdev=0.03)
    atoms = BatchedAtoms(pos.copy(), cell, np.full(R, n, dtype=np.int32), "Cu", (True, True, True), EMT())
    mc = HamiltonianCanonical(atoms, temperature=300.0, max_cycles=1, seed=8, trace=True, resync_interval=0,
                              default_displacement_move=HamiltonianDisplacementMove(operation=Verlet(dt=0.01, max_steps=20)))  # fmt: skip
    mc.run(1)
    kB = 8.617330337217213e-05
    t_kin = 2.0 * mc.context.last_kinetic_energy / (3 * n * kB)  # after a (practically always accepted) trajectory
    assert abs(t_kin.mean() - 300.0) < 1.5 and abs(t_kin.std() - 300.0 * np.sqrt(2.0 / (3 * n))) < 3.0
    assert np.max(np.abs(mc.last_trace["delta"])) < 1e-4  # H is conserved along the 0.2 fs trajectory
    assert mc.last_trace["accepted"].mean() > 0.99
