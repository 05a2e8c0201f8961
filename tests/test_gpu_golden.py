"""The CUDA path against the golden traces of the GENUINE quansino sources (tests/golden/, oracle/gen_golden.py)."""

import contextlib

import numpy as np
import pytest

from tests.golden_cases import CASES, gpu_replay, load

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_variant")]


@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_genuine_quansino(cuda_lib, name):
    g = load(name)
    mc = gpu_replay(g)
    tr = mc.last_trace
    assert np.array_equal(tr["accepted"][0], g["accepted"]), "accept / reject / failed sequence differs from quansino"
    assert np.array_equal(tr["move"][0], g["move"])
    np.testing.assert_allclose(tr["energy"][0], g["energy"], rtol=1e-10, atol=1e-10)
    out = mc.download()
    n = int(out.n_atoms[0])
    assert n == g["n_atoms"][-1]
    np.testing.assert_allclose(out.positions[0, :n], g["positions"], rtol=0, atol=1e-10)
    if g["kind"] in ("isobaric", "hybrid", "forced"):
        np.testing.assert_allclose(out.cell[0], g["cell"], rtol=1e-13)
    if g["kind"] == "general_cell":  # expm: scaling-and-squaring Taylor polynomial on the device, Pade approximants in scipy
        np.testing.assert_allclose(out.cell[0], g["cell"], rtol=0, atol=1e-12)
    if g["kind"] in ("gc_lj", "gc_emt", "gc_molecule", "gc_composite"):
        assert mc.number_of_exchange_particles[0] == g["n_exchange"]
        assert np.array_equal(mc.labels("default_displacement_move")[0, :n], g["labels_displacement"])
        assert np.array_equal(mc.labels("default_exchange_move")[0, :n], g["labels_exchange"])
    if g["kind"] == "hmc":
        np.testing.assert_allclose(out.momenta[0], g["momenta"], rtol=0, atol=1e-9)


def test_cuda_reproduces_genuine_force_bias(cuda_lib):
    """ForceBias over the batch API against the trajectory of the genuine sources, and 6 more replicas against the oracle."""
    from oracle import capi
    from oracle.potentials import emt_params
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import ForceBias
    from tests.helpers import cu_replicas

    g = load("force_bias")
    steps, n = len(g["energy"]), len(g["positions0"])
    atoms = BatchedAtoms(g["positions0"][None].copy(), g["cell"], np.array([n], np.int32), "Cu", (True,) * 3, EMT())
    fb = ForceBias(atoms, delta=g["delta"], temperature=g["temperature"], seed=int(g["seed"]), replica_offset=int(g["replica"]))
    for s in range(steps):
        fb.run(1)
        np.testing.assert_allclose(fb.download().positions[0], g["trajectory"][s], rtol=0, atol=1e-12)
        assert abs(fb.context.last_potential_energy[0] - g["energy"][s]) <= 1e-10 * abs(g["energy"][s])
        assert abs(fb.gamma[0] - g["gamma_max"][s]) <= 1e-9 * g["gamma_max"][s]
    assert fb.step_count == steps and np.all(fb.rejection_rounds >= 5)
    with pytest.raises(NotImplementedError):
        fb.add_move(None)
    # a batch at several temperatures, several steps, against the oracle replica by replica
    R, steps = 6, 3
    pos, cell, n = cu_replicas(3, R, seed=91)
    temps = np.linspace(200.0, 1500.0, R)
    fb = ForceBias(BatchedAtoms(pos.copy(), cell, np.full(R, n, np.int32), "Cu", (True,) * 3, EMT()), delta=0.08, temperature=temps, seed=77)
    fb.run(steps)
    out, e_gpu, par = fb.download(), fb.context.last_potential_energy, emt_params("Cu")
    for r in range(R):
        p = pos[r].copy()
        for s in range(steps):
            e, _, _ = capi.fbmc_step(par, p, cell, (1, 1, 1), 63.546, float(temps[r]), 0.08, 77, r, s)
        np.testing.assert_allclose(out.positions[r], p, rtol=0, atol=1e-11)
        assert abs(e_gpu[r] - e) <= 1e-10 * abs(e)
    f = fb.forces  # the forces left by the last step are the forces at the final positions
    ref = capi.energy(par, out.positions[0], cell, (1, 1, 1), want_forces=True)
    np.testing.assert_allclose(f[0], ref["forces"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", ["adaptive_force_bias", "adaptive_force_bias_exp"])
def test_cuda_reproduces_genuine_adaptive_force_bias(cuda_lib, name):
    """AdaptiveForceBias (mc/fbmc.py:285-488) of the genuine sources over EMT: the "forces" scheme falls back to the reference
    variance every step; the "energy" scheme falls back on the first step and then follows std(E_atom) / N."""
    from quansino_b200 import EMT, BatchedAtoms
    from quansino_b200.mc import AdaptiveForceBias

    g = load(name)
    steps, n = len(g["energy"]), len(g["positions0"])
    scheme, fn, lo, hi, ref = (("forces", "tanh", 0.04, 0.13, 0.1) if name == "adaptive_force_bias" else ("energy", "exp", 0.02, 0.2, 0.3))
    pos = np.repeat(g["positions0"][None], 2, axis=0)
    pos[1] *= 0.999  # a second replica with other atomic energies: its delta must differ in the energy scheme
    atoms = BatchedAtoms(pos, g["cell"], np.array([n, n], np.int32), "Cu", (True,) * 3, EMT())
    fb = AdaptiveForceBias(atoms, lo, hi, temperature=g["temperature"], scheme=scheme, update_function=fn, reference_variance=ref,
                           seed=int(g["seed"]), replica_offset=int(g["replica"]))  # fmt: skip
    for s in range(steps):
        with pytest.warns(UserWarning) if (scheme == "forces" or s == 0) else contextlib.nullcontext():
            fb.run(1)
        d = np.broadcast_to(np.asarray(fb.delta, float), (2,))
        assert abs(d[0] - g["delta_steps"][s]) <= 1e-13 * g["delta_steps"][s]
        assert (d[0] == d[1]) == (scheme == "forces" or s == 0)
        np.testing.assert_allclose(fb.download().positions[0], g["trajectory"][s], rtol=0, atol=1e-12)
        assert abs(fb.context.last_potential_energy[0] - g["energy"][s]) <= 1e-10 * abs(g["energy"][s])
    d = fb.to_dict()
    assert d["name"] == "AdaptiveForceBias" and d["kwargs"]["min_delta"] == lo and "delta" not in d["kwargs"]


def test_atom_energies_match_the_numpy_restatement(cuda_lib):
    """qmc_atom_energies = ``calc.results["energies"]``: EMT (periodic Cu) and Lennard-Jones (Pt cluster), sums = total energy."""
    from oracle.potentials import LJParams, emt_energy_forces, emt_params, lj_energy_forces
    from quansino_b200 import EMT, BatchedAtoms, LennardJones
    from quansino_b200.batch import ReplicaBatch
    from quansino_b200.systems import octahedron
    from tests.helpers import cu_replicas

    pos, cell, n = cu_replicas(3, 3, seed=5)
    b = ReplicaBatch(BatchedAtoms(pos.copy(), cell, np.full(3, n, np.int32), "Cu", (True,) * 3, EMT()))
    e = b.atom_energies().cpu().numpy()
    for r in range(3):
        ref = emt_energy_forces(pos[r], cell, (True,) * 3, emt_params("Cu"), want_forces=False)
        np.testing.assert_allclose(e[r, :n], ref["energies"], rtol=1e-11, atol=1e-12)
        assert abs(e[r, :n].sum() - float(b.energy[r])) <= 1e-11 * abs(ref["energy"])
    p, c = octahedron("Pt", 4, 10.0)
    padded = np.zeros((1, 64, 3))
    padded[0, : len(p)] = p
    b = ReplicaBatch(BatchedAtoms(padded, c, np.array([len(p)], np.int32), "Pt", (False,) * 3, LennardJones(sigma=2.5, epsilon=0.1)))
    ref = lj_energy_forces(p, c, (False,) * 3, LJParams(sigma=2.5, epsilon=0.1), want_forces=False)
    np.testing.assert_allclose(b.atom_energies().cpu().numpy()[0, : len(p)], ref["energies"], rtol=1e-11, atol=1e-12)
